#!/bin/bash
# Round-2 GPU session T (1 GPU): final tree -- full GPU suite as the driver runs it, smoke, bench, c5 with the 64-tap stem wgrad.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/t_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/t_t0) )) s] $*"; }
summ() {
python - "$1" <<'PY'
import json, sys
try:
  d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
except Exception as e:
  print("no bench json:", e); sys.exit(0)
r = d.get("roofline") or {}
print("%s [%s] pairs/s %.0f  ms/step %.2f  e2e %.0f  loss %.3e launches %d conv TF/s %.0f frac %.3f whole %.3f" % (d["config"]["name"], d["dtype"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"], d["gpu_launches"], r.get("achieved", 0), r.get("frac", 0), r.get("whole_step_frac", 0)))
if r: print("  other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])}, "clocks:", (d.get("clocks") or {}).get("sm_mhz"))
PY
}
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > $O/t_tests.log 2>&1
stamp "1 pytest tests/ -x -q -m gpu rc=$?"; tail -4 $O/t_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/t_smoke.log 2>&1; stamp "2 smoke() rc=$?"; tail -3 $O/t_smoke.log
timeout 400 python bench.py > $O/t_bench.json 2> $O/t_bench.err; stamp "3 bench.py rc=$?"; summ $O/t_bench.json
timeout 200 python bench.py --config c5 --steps 10 --no-cpu-baseline --also '' > $O/t_bench_c5.json 2> $O/t_bench_c5.err; stamp "4 bench c5 rc=$?"; summ $O/t_bench_c5.json
timeout 200 python bench.py --config c2 --steps 20 --no-cpu-baseline --also '' > $O/t_bench_c2.json 2> $O/t_bench_c2.err; stamp "5 bench c2 rc=$?"; summ $O/t_bench_c2.json
