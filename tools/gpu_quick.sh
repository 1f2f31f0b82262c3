#!/bin/bash
# quick perf iteration: build, conv kernel tests, conv sweep, bench (no ncu)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "conv or exact" > gpurun_out/quick_tests.log 2>&1; echo "conv tests rc=$?"; tail -4 gpurun_out/quick_tests.log
timeout 600 python tools/conv_sweep.py ${SWEEP_N:-1408} > gpurun_out/conv_sweep.txt 2>&1; cat gpurun_out/conv_sweep.txt
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
r = d.get("roofline", {})
print("pairs/s %.0f  ms/step %.1f  e2e %.0f  conv TF/s %.0f frac %.3f share %.2f  by_kind %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("achieved", 0), r.get("frac", 0), r.get("conv_share_of_step", 0), {k: (round(v["tflops"]), round(v["ms_per_step"], 1)) for k, v in r.get("by_kind", {}).items()}))
print("other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])})
PY
