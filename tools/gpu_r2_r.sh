#!/bin/bash
# Round-2 GPU session R (1 GPU): ncu --set full of the backward convolution kernels of the final build (multi-tile wgrad).
mkdir -p gpurun_out
O=gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel" -s 44 -c 40 -o $O/r_prof_bwd \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-roofline --also '' > $O/r_ncu_bwd.log 2>&1
echo "ncu rc=$?"; tail -2 $O/r_ncu_bwd.log | cut -c1-200; ls -la $O/r_prof_bwd.ncu-rep
