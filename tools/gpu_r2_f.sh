#!/bin/bash
# Round-2 GPU session F (1 GPU): stem wgrad on the tensor cores, 3xTF32 raw-hi operand, live per-layer conv table, final suite.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/f_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/f_t0) )) s] $*"; }
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
  print("    %-36s x%-3d %7.3f ms %6.0f TF" % (k, v["launches"], v["ms_per_step"], v["tflops"]))
PY
}
IIC_RUN_UNVALIDATED=1 timeout 400 python -m pytest tests -m "gpu and unvalidated" -q --tb=short -p no:cacheprovider --timeout 300 > $O/f_tests_unvalidated.log 2>&1
stamp "1 unvalidated tests (stem TC wgrad, raw hi) rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/f_tests_unvalidated.log | tail -20; grep -E "^E  " $O/f_tests_unvalidated.log | sort | uniq -c | sort -rn | head -12
timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' --layer-table > $O/f_bench_layers.json 2> $O/f_bench_layers.err; stamp "2 bench + layer table rc=$?"; tail -1 $O/f_bench_layers.err | cut -c1-200; summ $O/f_bench_layers.json
IIC_STEM_WGRAD_TC=1 timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/f_bench_stemtc.json 2> $O/f_bench_stemtc.err; stamp "3 bench stem wgrad TC rc=$?"; tail -1 $O/f_bench_stemtc.err | cut -c1-200; summ $O/f_bench_stemtc.json
for v in 0 1; do
  IIC_TF32X3_RAW_HI=$v timeout 300 python bench.py --precision tf32x3 --steps 3 --no-cpu-baseline --also '' > $O/f_bench_x3_raw$v.json 2> $O/f_bench_x3_raw$v.err; stamp "4 bench tf32x3 raw_hi=$v rc=$?"; tail -1 $O/f_bench_x3_raw$v.err | cut -c1-200; summ $O/f_bench_x3_raw$v.json
done
IIC_TF32X3_RAW_HI=1 IIC_SMOKE_MODES=tf32x3 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke_raw.log 2>&1; stamp "5 smoke tf32x3 raw hi rc=$?"; tail -3 $O/f_smoke_raw.log
IIC_TF32X3_RAW_HI=1 timeout 200 python tools/conv_sweep.py 352 tf32x3 > $O/f_sweep_x3_raw.txt 2>&1; stamp "6 sweep tf32x3 raw rc=$?"; tail -2 $O/f_sweep_x3_raw.txt
timeout 200 python bench.py --pairs-per-gpu 88 --graph --steps 20 --no-cpu-baseline --also "" --no-roofline > $O/f_bench_88_layers.json 2> $O/f_bench_88_layers.err; stamp "7 bench 88 pairs graph rc=$?"; summ $O/f_bench_88_layers.json
timeout 200 python bench.py --pairs-per-gpu 88 --steps 20 --no-cpu-baseline --also '' --layer-table > $O/f_bench_88e_layers.json 2> $O/f_bench_88e_layers.err; stamp "7b bench 88 pairs eager rc=$?"; summ $O/f_bench_88e_layers.json
timeout 700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/f_tests_serial.log 2>&1
stamp "8 suite as the driver runs it rc=$?"; tail -4 $O/f_tests_serial.log
