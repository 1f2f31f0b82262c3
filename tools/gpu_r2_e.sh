#!/bin/bash
# Round-2 GPU session E (1 GPU): driver-style suite, remaining unvalidated tests, wgrad-on-a-second-stream A/B.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/e_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/e_t0) )) s] $*"; }
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
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/e_tests_serial.log 2>&1
stamp "1 suite, serial, defaults, as the driver runs it rc=$?"; tail -4 $O/e_tests_serial.log
IIC_RUN_UNVALIDATED=1 timeout 400 python -m pytest tests -m "gpu and unvalidated" -q --tb=short -p no:cacheprovider --timeout 300 > $O/e_tests_unvalidated.log 2>&1
stamp "2 remaining unvalidated tests rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/e_tests_unvalidated.log | tail -20; grep -E "^E  " $O/e_tests_unvalidated.log | sort | uniq -c | sort -rn | head -12
for v in "" IIC_BN_BWD_CTAS=1 "IIC_BN_BWD_CTAS=1 IIC_WGRAD_STREAM=1" IIC_WGRAD_STREAM=1 ""; do
  f=$(echo "x$v" | tr ' =' '__')
  env $v timeout 200 python bench.py --steps 10 --no-cpu-baseline --no-roofline --also '' > $O/e_bench_$f.json 2> $O/e_bench_$f.err; stamp "3 bench [$v] rc=$?"; tail -1 $O/e_bench_$f.err | cut -c1-200; summ $O/e_bench_$f.json
done
IIC_BN_BWD_CTAS=1 IIC_WGRAD_STREAM=1 timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/e_bench_wstream_roof.json 2> $O/e_bench_wstream_roof.err; stamp "4 bench wgrad stream with breakdown rc=$?"; summ $O/e_bench_wstream_roof.json
IIC_BN_BWD_CTAS=1 IIC_WGRAD_STREAM=1 timeout 200 python bench.py --graph --steps 10 --no-cpu-baseline --no-roofline --also '' > $O/e_bench_wstream_graph.json 2> $O/e_bench_wstream_graph.err; stamp "5 bench wgrad stream + graph rc=$?"; tail -1 $O/e_bench_wstream_graph.err | cut -c1-200; summ $O/e_bench_wstream_graph.json
IIC_BN_BWD_CTAS=1 IIC_WGRAD_STREAM=1 timeout 200 python bench.py --config c5 --steps 5 --no-cpu-baseline --no-roofline --also '' > $O/e_bench_c5_wstream.json 2> $O/e_bench_c5_wstream.err; stamp "6 bench c5 wgrad stream rc=$?"; summ $O/e_bench_c5_wstream.json
timeout 200 python bench.py --config c5 --steps 5 --no-cpu-baseline --no-roofline --also '' > $O/e_bench_c5.json 2> $O/e_bench_c5.err; stamp "6b bench c5 rc=$?"; summ $O/e_bench_c5.json
