#!/bin/bash
# Round-2 GPU session Q (1 GPU): the driver's end-of-round sequence on the final tree.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/q_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/q_t0) )) s] $*"; }
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
pm = d.get("precision_modes") or {}
for k, v in pm.items(): print("  also %s: %.0f pairs/s, %.1f ms/step, conv frac %.3f" % (k, v["value"], v["ms_per_step"], (v.get("roofline") or {}).get("frac", 0)))
cb = d.get("cpu_baseline")
if cb: print("  cpu_baseline:", cb)
PY
}
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > $O/q_tests.log 2>&1
stamp "1 pytest tests/ -x -q -m gpu rc=$?"; tail -4 $O/q_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/q_smoke.log 2>&1; stamp "2 smoke() rc=$?"; tail -4 $O/q_smoke.log
timeout 400 python bench.py --impl reference > $O/q_ref.json 2> $O/q_ref.err; stamp "3 bench.py --impl reference rc=$?"; tail -1 $O/q_ref.json | cut -c1-400
timeout 400 python bench.py > $O/q_bench.json 2> $O/q_bench.err; stamp "4 bench.py rc=$?"; tail -1 $O/q_bench.err | cut -c1-200; summ $O/q_bench.json
for c in c2 c3 c5; do
  timeout 200 python bench.py --config $c --steps 10 --no-cpu-baseline --also '' > $O/q_bench_$c.json 2> $O/q_bench_$c.err; stamp "5 bench $c rc=$?"; summ $O/q_bench_$c.json
done
timeout 200 python bench.py --pairs-per-gpu 88 --steps 20 --no-cpu-baseline --also '' --no-roofline > $O/q_bench_88.json 2> $O/q_bench_88.err; stamp "6 bench 88 pairs eager rc=$?"; summ $O/q_bench_88.json
timeout 200 python bench.py --pairs-per-gpu 88 --graph --steps 20 --no-cpu-baseline --also '' --no-roofline > $O/q_bench_88g.json 2> $O/q_bench_88g.err; stamp "6b bench 88 pairs graph rc=$?"; summ $O/q_bench_88g.json
timeout 300 python bench.py --precision tf32x3 --steps 3 --no-cpu-baseline --also '' > $O/q_bench_x3.json 2> $O/q_bench_x3.err; stamp "7 bench tf32x3 rc=$?"; summ $O/q_bench_x3.json
