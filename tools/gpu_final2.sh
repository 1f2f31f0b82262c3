#!/bin/bash
# Second one-call validation: the variants added after v9 (dgrad addend prefetch, interleaved quad stem, fused stem
# backward, two-tile N=128 work items, TMA-store halo epilogue), each switchable; full suite with the defaults, a
# fallback suite if anything fails, the official bench line, one bench per variant switched off, smoke, sweep, ncu.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/final2_t0
[ -f iic_b200/lib/libiic_b200.so ] || python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
stamp() { echo "[$(( $(date +%s) - $(cat $O/final2_t0) )) s] $*"; }
OLD="IIC_DGRAD_PREFETCH=0 IIC_STEM_QUAD=1 IIC_STEM_BWD_FUSED=0 IIC_TC2_MT2=0 IIC_CONV_HALO_STORE=0"
summ() {
python - "$1" <<'PY'
import json, sys
try:
  d = json.load(open(sys.argv[1]))
except Exception as e:
  print("no bench json:", e); sys.exit(0)
r = d.get("roofline", {})
print("pairs/s %.0f  ms/step %.2f  e2e %.0f  loss %.3e launches %d conv TF/s %.0f  by_kind %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"], d["gpu_launches"], r.get("achieved", 0), {k: (round(v["tflops"]), round(v["ms_per_step"], 2)) for k, v in r.get("by_kind", {}).items()}))
print("other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])})
print("variants:", d.get("kernel_variants"), "clocks:", d.get("clocks"))
PY
}
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -n 4 --timeout 200 > $O/final2_tests_new.log 2>&1
rcA=$?; stamp "A full suite (new defaults) rc=$rcA"; grep -E "^(FAILED|ERROR)|passed|failed" $O/final2_tests_new.log | tail -30
timeout 200 python bench.py > $O/bench2_new.json 2> $O/bench2_new.err; stamp "B bench default rc=$?"; tail -2 $O/bench2_new.err; summ $O/bench2_new.json
if [ $rcA -ne 0 ]; then
  env $OLD timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -n 4 --timeout 200 > $O/final2_tests_old.log 2>&1
  stamp "C full suite (v9 variants) rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/final2_tests_old.log | tail -30
fi
env $OLD timeout 120 python bench.py --steps 5 --no-cpu-baseline > $O/bench2_old.json 2> $O/bench2_old.err; stamp "D bench v9 variants rc=$?"; summ $O/bench2_old.json
for v in IIC_CONV_HALO_STORE=0 IIC_STEM_BWD_FUSED=0 IIC_DGRAD_PREFETCH=0 IIC_TC2_MT2=0 IIC_STEM_QUAD=1; do
  env $v timeout 120 python bench.py --steps 5 --no-cpu-baseline > $O/bench2_$v.json 2> $O/bench2_$v.err; stamp "E bench $v rc=$?"; summ $O/bench2_$v.json
done
timeout 120 python __graft_entry__.py smoke > $O/smoke2.log 2>&1; stamp "G smoke rc=$?"; tail -2 $O/smoke2.log
timeout 120 python tools/conv_sweep.py 1408 > $O/conv_sweep_final2.txt 2>&1; stamp "F sweep rc=$?"; cat $O/conv_sweep_final2.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel|stem_bwd_wgrad" -s 5 -c 4 -o $O/prof_final2 \
   python bench.py --steps 1 --warmup 1 --pairs-per-gpu 176 --no-cpu-baseline --no-roofline > $O/ncu_full2.log 2>&1
stamp "I ncu full rc=$?"
