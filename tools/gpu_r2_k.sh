#!/bin/bash
# Round-2 GPU session K (1 GPU): masked addend (the masked residual gradient is never written), suite with bn_bitmask default.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/k_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/k_t0) )) s] $*"; }
summ() {
python - "$1" <<'PY'
import json, sys
try:
  d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
except Exception as e:
  print("no bench json:", e); sys.exit(0)
r = d.get("roofline", {})
print("%s [%s] pairs/s %.0f  ms/step %.2f  e2e %.0f  loss %.3e launches %d conv TF/s %.0f frac %.3f whole %.3f by_kind %s" % (d["config"]["name"], d["dtype"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"], d["gpu_launches"], r.get("achieved", 0), r.get("frac", 0), r.get("whole_step_frac", 0), {k: (round(v["tflops"]), round(v["ms_per_step"], 2)) for k, v in r.get("by_kind", {}).items()}))
print("  other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])}, "clocks:", (d.get("clocks") or {}).get("sm_mhz"))
PY
}
IIC_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests -m "gpu and unvalidated" -q --tb=short -p no:cacheprovider --timeout 200 > $O/k_tests_unvalidated.log 2>&1
stamp "1 unvalidated (masked addend) rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/k_tests_unvalidated.log | tail -20; grep -E "^E  " $O/k_tests_unvalidated.log | sort | uniq -c | sort -rn | head -12
timeout 700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/k_tests_serial.log 2>&1
stamp "2 suite as the driver runs it (bn_bitmask default) rc=$?"; tail -5 $O/k_tests_serial.log
for v in "" "IIC_MASKED_ADDEND=1" "" "IIC_MASKED_ADDEND=1"; do
  f=$(echo "x$v" | tr ' =' '__')
  env $v timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/k_bench_$f.json 2> $O/k_bench_$f.err; stamp "3 bench [$v] rc=$?"; tail -1 $O/k_bench_$f.err | cut -c1-200; summ $O/k_bench_$f.json
done
IIC_MASKED_ADDEND=1 timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_precision.py tests/test_gpu_parity_nets.py tests/test_gpu_batch.py -m gpu -x -q -p no:cacheprovider > $O/k_tests_masked.log 2>&1
stamp "4 step / precision / net tests with masked_addend rc=$?"; tail -4 $O/k_tests_masked.log
IIC_MASKED_ADDEND=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/k_smoke.log 2>&1; stamp "5 smoke rc=$?"; tail -4 $O/k_smoke.log
