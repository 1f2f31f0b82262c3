#!/bin/bash
# Round-2 GPU session M (1 GPU): 32-bit index arithmetic in the pooling / stem-backward kernels; stem_bwd_dy re-measured.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/m_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/m_t0) )) s] $*"; }
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
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider --timeout 200 -k "stem or pool or sobel or maxpool" > $O/m_tests.log 2>&1
stamp "1 stem / pool / sobel kernel tests rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/m_tests.log | tail -8
for v in "" "IIC_STEM_BWD_DY=1" "" "IIC_STEM_BWD_DY=1"; do
  f=$(echo "x$v" | tr ' =' '__')
  env $v timeout 200 python bench.py --steps 10 --no-cpu-baseline --also '' > $O/m_bench_$f.json 2> $O/m_bench_$f.err; stamp "2 bench [$v] rc=$?"; tail -1 $O/m_bench_$f.err | cut -c1-200; summ $O/m_bench_$f.json
done
env IIC_STEM_BWD_DY=1 timeout 300 python -m pytest tests/test_gpu_step.py tests/test_gpu_precision.py tests/test_gpu_parity_nets.py -m gpu -x -q -p no:cacheprovider > $O/m_tests_dy.log 2>&1
stamp "3 step / precision / nets with stem_bwd_dy rc=$?"; tail -3 $O/m_tests_dy.log
