#!/bin/bash
# Runs the GPU test-suite on the B200 box file by file (own timeout each, full logs kept).
# usage: gpurun -- bash tools/gpu_tests.sh [extra pytest args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
rc=0
for f in tests/test_gpu_kernels.py tests/test_gpu_parity_iid.py tests/test_gpu_parity_nets.py tests/test_gpu_parity_seg.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --tb=short -p no:cacheprovider "$@" > gpurun_out/$b.log 2>&1
  r=$?; [ $r -ne 0 ] && rc=$r
  echo "== $f rc=$r"; tail -25 gpurun_out/$b.log
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; r=$?; [ $r -ne 0 ] && rc=$r
echo "== smoke rc=$r"; tail -8 gpurun_out/smoke.log
exit $rc
