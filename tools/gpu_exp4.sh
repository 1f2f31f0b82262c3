#!/bin/bash
# experiment round 4: halo wgrad, quad stem + fused stats, batched packing
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/exp4_tests.log 2>&1; echo "kernel tests rc=$?"; tail -30 gpurun_out/exp4_tests.log
IIC_CONV_HALO=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "conv or exact or adjoint" > gpurun_out/exp4_halo_tests.log 2>&1; echo "halo(forced) conv tests rc=$?"; tail -25 gpurun_out/exp4_halo_tests.log
timeout 300 python -m pytest tests/test_gpu_parity_nets.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/exp4_nets.log 2>&1; echo "nets rc=$?"; tail -8 gpurun_out/exp4_nets.log
timeout 300 python tools/conv_sweep.py 1408 2>&1 | head -1
summ() {
python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d.get("roofline", {})
print("pairs/s %.0f  ms/step %.1f  e2e %.0f  launches %d conv TF/s %.0f  by_kind %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], r.get("achieved", 0), {k: (round(v["tflops"]), round(v["ms_per_step"], 1)) for k, v in r.get("by_kind", {}).items()}))
print("other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])})
PY
}
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_exp4.json 2> gpurun_out/bench_exp4.err; tail -2 gpurun_out/bench_exp4.err; summ gpurun_out/bench_exp4.json
