#!/bin/bash
# experiment round 3: uniform-register MMA issue path (all tcgen05 conv kernels) + halo variant
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "conv or exact or adjoint" > gpurun_out/exp3_tests.log 2>&1; echo "conv tests rc=$?"; tail -5 gpurun_out/exp3_tests.log
IIC_CONV_HALO=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "conv or exact or adjoint" > gpurun_out/exp3_halo_tests.log 2>&1; echo "halo(forced) conv tests rc=$?"; tail -5 gpurun_out/exp3_halo_tests.log
echo "--- sweep, im2col kernels"; IIC_CONV_HALO=0 timeout 300 python tools/conv_sweep.py 1408 2>&1 | tee gpurun_out/conv_sweep_exp3.txt
echo "--- sweep, halo"; IIC_CONV_HALO=1 timeout 300 python tools/conv_sweep.py 1408 2>&1 | head -1
summ() {
python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d.get("roofline", {})
print("pairs/s %.0f  ms/step %.1f  e2e %.0f  launches %d conv TF/s %.0f  by_kind %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], r.get("achieved", 0), {k: (round(v["tflops"]), round(v["ms_per_step"], 1)) for k, v in r.get("by_kind", {}).items()}))
PY
}
IIC_CONV_HALO=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_exp3_base.json 2> gpurun_out/bench_exp3_base.err; tail -2 gpurun_out/bench_exp3_base.err; summ gpurun_out/bench_exp3_base.json
IIC_CONV_HALO=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_exp3_halo.json 2> gpurun_out/bench_exp3_halo.err; tail -2 gpurun_out/bench_exp3_halo.err; summ gpurun_out/bench_exp3_halo.json
