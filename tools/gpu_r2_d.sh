#!/bin/bash
# Round-2 GPU session D (1 GPU): the suite as the driver runs it (serial, defaults), the remaining `unvalidated` tests,
# smoke() as the driver runs it, the official bench line, the fused wgrad-unpack switch, bench c5.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/d_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/d_t0) )) s] $*"; }
summ() {
python - "$1" <<'PY'
import json, sys
try:
  d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
except Exception as e:
  print("no bench json:", e); sys.exit(0)
r = d.get("roofline", {})
print("%s [%s] pairs/s %.0f  ms/step %.2f  e2e %.0f  loss %.3e launches %d conv TF/s %.0f frac %.3f by_kind %s" % (d["config"]["name"], d["dtype"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"], d["gpu_launches"], r.get("achieved", 0), r.get("frac", 0), {k: (round(v["tflops"]), round(v["ms_per_step"], 2)) for k, v in r.get("by_kind", {}).items()}))
print("  other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])})
for m, e in d.get("precision_modes", {}).items():
  rr = e.get("roofline", {})
  print("  mode %s: pairs/s %.0f ms/step %.1f conv TF/s %.0f frac %.3f by_kind %s" % (m, e["value"], e["ms_per_step"], rr.get("achieved", 0), rr.get("frac", 0), {k: round(v["tflops"]) for k, v in rr.get("by_kind", {}).items()}))
print("  traffic:", r.get("traffic"), "clocks:", d.get("clocks"), "cpu:", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
PY
}
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/d_tests_serial.log 2>&1
stamp "1 suite, serial, defaults, as the driver runs it rc=$?"; tail -4 $O/d_tests_serial.log
IIC_RUN_UNVALIDATED=1 timeout 400 python -m pytest tests -m "gpu and unvalidated" -q --tb=short -p no:cacheprovider -n 4 --timeout 300 > $O/d_tests_unvalidated.log 2>&1
stamp "2 remaining unvalidated tests rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/d_tests_unvalidated.log | tail -20; grep -E "^E  " $O/d_tests_unvalidated.log | sort | uniq -c | sort -rn | head -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/d_smoke.log 2>&1; stamp "3 smoke rc=$?"; tail -5 $O/d_smoke.log
timeout 400 python bench.py > $O/d_bench.json 2> $O/d_bench.err; stamp "4 bench default rc=$?"; tail -2 $O/d_bench.err; summ $O/d_bench.json
IIC_WGRAD_FUSED=1 timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/d_bench_wgrad_fused.json 2> $O/d_bench_wgrad_fused.err; stamp "5 bench, wgrad writes the torch layout rc=$?"; summ $O/d_bench_wgrad_fused.json
timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/d_bench_again.json 2> $O/d_bench_again.err; stamp "5b bench default again (A/B) rc=$?"; summ $O/d_bench_again.json
timeout 200 python bench.py --graph --steps 10 --no-cpu-baseline --also '' > $O/d_bench_graph.json 2> $O/d_bench_graph.err; stamp "5c bench --graph rc=$?"; summ $O/d_bench_graph.json
IIC_BN_BWD_CTAS=1 timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/d_bench_bnctas1.json 2> $O/d_bench_bnctas1.err; stamp "5d bench, bn_bwd 1 CTA/SM rc=$?"; summ $O/d_bench_bnctas1.json
IIC_BN_BWD_CTAS=1 IIC_WGRAD_STREAM=1 timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/d_bench_wstream.json 2> $O/d_bench_wstream.err; stamp "5e bench, wgrad on a second stream rc=$?"; tail -2 $O/d_bench_wstream.err; summ $O/d_bench_wstream.json
IIC_WGRAD_STREAM=1 timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/d_bench_wstream2.json 2> $O/d_bench_wstream2.err; stamp "5f bench, wgrad on a second stream, bn_bwd 2 CTAs/SM rc=$?"; summ $O/d_bench_wstream2.json
timeout 200 python bench.py --config c5 --steps 5 --no-cpu-baseline > $O/d_bench_c5.json 2> $O/d_bench_c5.err; stamp "6 bench c5 rc=$?"; summ $O/d_bench_c5.json
timeout 200 python bench.py --impl reference --steps 3 > $O/d_ref.json 2> $O/d_ref.err; stamp "7 reference arm rc=$?"; tail -c 600 $O/d_ref.json
