#!/bin/bash
# Round-2 GPU session G (1 GPU): dedicated tcgen05 stem wgrad, pre-split 3xTF32 weights.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/g_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/g_t0) )) s] $*"; }
summ() {
python - "$1" <<'PY'
import json, sys
try:
  d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
except Exception as e:
  print("no bench json:", e); sys.exit(0)
r = d.get("roofline", {})
print("%s [%s] pairs/s %.0f  ms/step %.2f  e2e %.0f  loss %.3e launches %d conv TF/s %.0f frac %.3f by_kind %s" % (d["config"]["name"], d["dtype"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"], d["gpu_launches"], r.get("achieved", 0), r.get("frac", 0), {k: (round(v["tflops"]), round(v["ms_per_step"], 2)) for k, v in r.get("by_kind", {}).items()}))
print("  other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])}, "clocks:", (d.get("clocks") or {}).get("sm_mhz"))
PY
}
IIC_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv_tf32.py -m gpu -q --tb=short -p no:cacheprovider --timeout 200 > $O/g_tests_kernels.log 2>&1
stamp "1 kernel + tf32 conv tests (incl. unvalidated) rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/g_tests_kernels.log | tail -20; grep -E "^E  " $O/g_tests_kernels.log | sort | uniq -c | sort -rn | head -12
timeout 300 python -m pytest tests/test_gpu_precision.py tests/test_gpu_step.py -m gpu -q --tb=short -p no:cacheprovider --timeout 250 > $O/g_tests_prec.log 2>&1
stamp "2 precision + step tests rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/g_tests_prec.log | tail -20; grep -E "^E  " $O/g_tests_prec.log | sort | uniq -c | sort -rn | head -12
timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/g_bench_default.json 2> $O/g_bench_default.err; stamp "3 bench default rc=$?"; tail -1 $O/g_bench_default.err | cut -c1-200; summ $O/g_bench_default.json
IIC_STEM_WGRAD_TC=1 timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/g_bench_stemtc.json 2> $O/g_bench_stemtc.err; stamp "4 bench stem wgrad tcgen05 rc=$?"; tail -1 $O/g_bench_stemtc.err | cut -c1-200; summ $O/g_bench_stemtc.json
timeout 300 python bench.py --precision tf32x3 --steps 3 --no-cpu-baseline --also '' > $O/g_bench_x3.json 2> $O/g_bench_x3.err; stamp "5 bench tf32x3 (pre-split weights) rc=$?"; tail -1 $O/g_bench_x3.err | cut -c1-200; summ $O/g_bench_x3.json
IIC_STEM_WGRAD_TC=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/g_smoke.log 2>&1; stamp "6 smoke (all modes, stem TC) rc=$?"; tail -4 $O/g_smoke.log
timeout 200 python tools/conv_sweep.py 352 tf32x3 > $O/g_sweep_x3.txt 2>&1; stamp "7 sweep tf32x3 rc=$?"; tail -12 $O/g_sweep_x3.txt
