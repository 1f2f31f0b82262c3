#!/bin/bash
# Round-2 GPU session H (1 GPU): wgrad with 2-3 tiles per dy k-block, stem wgrad with four builder groups.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/h_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/h_t0) )) s] $*"; }
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
for k, v in (r.get("by_layer") or {}).items():
  if k.startswith("wgrad"): print("    %-36s x%-3d %7.3f ms %6.0f TF" % (k, v["launches"], v["ms_per_step"], v["tflops"]))
PY
}
IIC_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider --timeout 200 -k "wgrad or stem" > $O/h_tests_kernels.log 2>&1
stamp "1 wgrad + stem tests (incl. unvalidated) rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/h_tests_kernels.log | tail -20; grep -E "^E  " $O/h_tests_kernels.log | sort | uniq -c | sort -rn | head -12
for v in "" "IIC_WGRAD_MT=1" "IIC_STEM_WGRAD_TC=1" "IIC_WGRAD_MT=1 IIC_STEM_WGRAD_TC=1"; do
  f=$(echo "x$v" | tr ' =' '__')
  env $v timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' --layer-table > $O/h_bench_$f.json 2> $O/h_bench_$f.err; stamp "2 bench [$v] rc=$?"; tail -1 $O/h_bench_$f.err | cut -c1-200; summ $O/h_bench_$f.json
done
IIC_WGRAD_MT=1 IIC_STEM_WGRAD_TC=1 timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/h_tests_all.log 2>&1
stamp "3 whole suite with both switches on rc=$?"; tail -4 $O/h_tests_all.log
IIC_WGRAD_MT=1 IIC_STEM_WGRAD_TC=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/h_smoke.log 2>&1; stamp "4 smoke rc=$?"; tail -4 $O/h_smoke.log
IIC_WGRAD_MT=1 timeout 200 python tools/conv_sweep.py 1408 bf16 > $O/h_sweep_mt.txt 2>&1; stamp "5 sweep bf16 wgrad_mt rc=$?"; tail -12 $O/h_sweep_mt.txt
for c in c2 c3 c5; do
  IIC_WGRAD_MT=1 IIC_STEM_WGRAD_TC=1 timeout 200 python bench.py --config $c --steps 10 --no-cpu-baseline --also '' > $O/h_bench_$c.json 2> $O/h_bench_$c.err; stamp "6 bench $c rc=$?"; tail -1 $O/h_bench_$c.err | cut -c1-200; summ $O/h_bench_$c.json
done
