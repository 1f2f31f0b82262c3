#!/bin/bash
# Round-2 GPU session P (1 GPU): profiles of the final build -- launch list of the bench command, ncu --set full of the new
# kernels (multi-tile wgrad, tcgen05 stem wgrad, halo dgrad with the TMA addend, the HBM-bound stem / BatchNorm passes).
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/p_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/p_t0) )) s] $*"; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/p_launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --also '' > $O/p_ncu_list.log 2>&1
stamp "1 ncu launch list rc=$?"; python tools/ncu_launch_table.py $O/p_launches.csv "launch list" "" 2>/dev/null | head -30
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-roofline --also ''"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stem_wgrad_tc_kernel|bn_relu_maxpool|stem_fprop64q" -c 8 -o $O/p_prof_stem \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-roofline --also '' > $O/p_ncu_stem.log 2>&1
stamp "2a ncu --set full: stem kernels rc=$?"; tail -1 $O/p_ncu_stem.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel<\(int\)1" -c 28 -o $O/p_prof_wgrad \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-roofline --also '' > $O/p_ncu_wgrad.log 2>&1
stamp "2b ncu --set full: wgrad kernels rc=$?"; tail -1 $O/p_ncu_wgrad.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_fused_kernel|bn_apply_kernel" -c 10 -o $O/p_prof_bn \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-roofline --also '' > $O/p_ncu_bn.log 2>&1
stamp "2c ncu --set full: BatchNorm passes rc=$?"; tail -1 $O/p_ncu_bn.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel" -s 6 -c 6 -o $O/p_prof_halo \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-roofline --also '' > $O/p_ncu_halo.log 2>&1
stamp "3 ncu --set full: halo dgrad rc=$?"; tail -1 $O/p_ncu_halo.log | cut -c1-200
ls -la $O/p_*.ncu-rep
