#!/bin/bash
# Round-2 GPU session S (1 GPU): split-K head GEMMs; the small configurations from a CUDA graph.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/s_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/s_t0) )) s] $*"; }
summ() {
python - "$1" <<'PY'
import json, sys
try:
  d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
except Exception as e:
  print("no bench json:", e); sys.exit(0)
r = d.get("roofline") or {}
print("%s [%s] pairs/s %.0f  ms/step %.2f  e2e %.0f  loss %.3e launches %d" % (d["config"]["name"], d["dtype"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"], d["gpu_launches"]))
if r: print("  other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])})
PY
}
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_nets.py tests/test_gpu_step.py -m gpu -x -q -p no:cacheprovider -k "heads or 6c or step" > $O/s_tests.log 2>&1
stamp "1 heads / 6c / step tests rc=$?"; tail -3 $O/s_tests.log
timeout 200 python bench.py --config c2 --steps 20 --no-cpu-baseline --also '' > $O/s_bench_c2.json 2> $O/s_bench_c2.err; stamp "2 bench c2 rc=$?"; summ $O/s_bench_c2.json
timeout 200 python bench.py --config c2 --graph --steps 20 --no-cpu-baseline --also '' --no-roofline > $O/s_bench_c2g.json 2> $O/s_bench_c2g.err; stamp "2b bench c2 graph rc=$?"; summ $O/s_bench_c2g.json
timeout 200 python bench.py --config c3 --graph --steps 20 --no-cpu-baseline --also '' --no-roofline > $O/s_bench_c3g.json 2> $O/s_bench_c3g.err; stamp "3 bench c3 graph rc=$?"; summ $O/s_bench_c3g.json
timeout 200 python bench.py --config c5 --graph --steps 10 --no-cpu-baseline --also '' --no-roofline > $O/s_bench_c5g.json 2> $O/s_bench_c5g.err; stamp "4 bench c5 graph rc=$?"; tail -1 $O/s_bench_c5g.err | cut -c1-200; summ $O/s_bench_c5g.json
timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/s_bench_c4.json 2> $O/s_bench_c4.err; stamp "5 bench c4 rc=$?"; summ $O/s_bench_c4.json
