#!/bin/bash
# Round-2 GPU session I (1 GPU): new defaults (wgrad_mt, stem wgrad tcgen05, parallel BN fold), halo addend by TMA, full suite.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/i_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/i_t0) )) s] $*"; }
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
  if k.startswith("dgrad 3x3 s1 64"): print("    %-36s x%-3d %7.3f ms %6.0f TF" % (k, v["launches"], v["ms_per_step"], v["tflops"]))
PY
}
timeout 700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/i_tests_serial.log 2>&1
stamp "1 suite as the driver runs it rc=$?"; tail -5 $O/i_tests_serial.log
IIC_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests -m "gpu and unvalidated" -q --tb=short -p no:cacheprovider --timeout 200 > $O/i_tests_unvalidated.log 2>&1
stamp "2 unvalidated (halo addend by TMA) rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/i_tests_unvalidated.log | tail -20; grep -E "^E  " $O/i_tests_unvalidated.log | sort | uniq -c | sort -rn | head -12
for v in "" "IIC_HALO_ADDEND_TMA=1" ""; do
  f=$(echo "x$v" | tr ' =' '__')
  env $v timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' --layer-table > $O/i_bench_$f.json 2> $O/i_bench_$f.err; stamp "3 bench [$v] rc=$?"; tail -1 $O/i_bench_$f.err | cut -c1-200; summ $O/i_bench_$f.json
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/i_smoke.log 2>&1; stamp "4 smoke rc=$?"; tail -4 $O/i_smoke.log
timeout 300 python bench.py > $O/i_bench_default_full.json 2> $O/i_bench_default_full.err; stamp "5 bench default (as the driver runs it) rc=$?"; tail -1 $O/i_bench_default_full.err | cut -c1-200; summ $O/i_bench_default_full.json
timeout 200 python bench.py --pairs-per-gpu 88 --graph --steps 20 --no-cpu-baseline --also "" --no-roofline > $O/i_bench_88g.json 2> $O/i_bench_88g.err; stamp "6 bench 88 pairs graph rc=$?"; summ $O/i_bench_88g.json
timeout 200 python bench.py --graph --steps 10 --no-cpu-baseline --also "" --no-roofline > $O/i_bench_graph.json 2> $O/i_bench_graph.err; stamp "7 bench 704 pairs graph rc=$?"; summ $O/i_bench_graph.json
