#!/bin/bash
# Round-2 GPU session L (1 GPU): stem_bwd_dy, dgrad_s2_mt (and whatever is still marked unvalidated).
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/l_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/l_t0) )) s] $*"; }
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
for k, v in (r.get("by_layer") or {}).items():
  if k.startswith("dgrad") and " s2 " in k: print("    %-36s x%-3d %7.3f ms %6.0f TF" % (k, v["launches"], v["ms_per_step"], v["tflops"]))
PY
}
IIC_RUN_UNVALIDATED=1 timeout 400 python -m pytest tests -m "gpu and unvalidated" -q --tb=short -p no:cacheprovider --timeout 200 > $O/l_tests_unvalidated.log 2>&1
stamp "1 unvalidated rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/l_tests_unvalidated.log | tail -20; grep -E "^E  " $O/l_tests_unvalidated.log | sort | uniq -c | sort -rn | head -12
for v in "" "IIC_STEM_BWD_DY=1" "IIC_DGRAD_S2_MT=1" "$EXTRA" "IIC_STEM_BWD_DY=1 IIC_DGRAD_S2_MT=1 $EXTRA" ""; do
  f=$(echo "x$v" | tr ' =' '__')
  env $v timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' --layer-table > $O/l_bench_$f.json 2> $O/l_bench_$f.err; stamp "2 bench [$v] rc=$?"; tail -1 $O/l_bench_$f.err | cut -c1-200; summ $O/l_bench_$f.json
done
env IIC_STEM_BWD_DY=1 IIC_DGRAD_S2_MT=1 $EXTRA timeout 700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/l_tests_all.log 2>&1
stamp "3 whole suite with the new switches on rc=$?"; tail -4 $O/l_tests_all.log
env IIC_STEM_BWD_DY=1 IIC_DGRAD_S2_MT=1 $EXTRA timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/l_smoke.log 2>&1; stamp "4 smoke rc=$?"; tail -4 $O/l_smoke.log
