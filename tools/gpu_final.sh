#!/bin/bash
# One-call validation of the round: full GPU suite + smoke + bench with the default kernel variants, then the
# conservative variant set (halo wgrad / quad stem / fused stem statistics / batched packing off) for comparison,
# then A/B benches and the ncu launch list.  Every stage has its own timeout and writes to gpurun_out/ as it goes,
# so a clamped call still returns what finished.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/final_t0
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
[ -f iic_b200/lib/libiic_b200.so ] || python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
stamp() { echo "[$(( $(date +%s) - $(cat $O/final_t0) )) s] $*"; }
SAFE="IIC_CONV_HALO_WGRAD=0 IIC_STEM_QUAD=0 IIC_STEM_STATS=0 IIC_PACK_BATCHED=0"
summ() {
python - "$1" <<'PY'
import json, sys
try:
  d = json.load(open(sys.argv[1]))
except Exception as e:
  print("no bench json:", e); sys.exit(0)
r = d.get("roofline", {})
print("pairs/s %.0f  ms/step %.2f  e2e %.0f  launches %d conv TF/s %.0f  by_kind %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], r.get("achieved", 0), {k: (round(v["tflops"]), round(v["ms_per_step"], 2)) for k, v in r.get("by_kind", {}).items()}))
print("other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])})
print("variants:", d.get("kernel_variants"), "clocks:", d.get("clocks"))
PY
}
# A: full suite, default variants
timeout 420 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -n 4 --timeout 240 > $O/final_tests_new.log 2>&1
stamp "A full suite (default variants) rc=$?"; tail -15 $O/final_tests_new.log
# B: official bench line, default variants
timeout 240 python bench.py > $O/bench_new.json 2> $O/bench_new.err; stamp "B bench default rc=$?"; tail -2 $O/bench_new.err; summ $O/bench_new.json
# C: bench, conservative variants
env $SAFE timeout 200 python bench.py --steps 5 --no-cpu-baseline > $O/bench_safe.json 2> $O/bench_safe.err; stamp "C bench conservative rc=$?"; tail -2 $O/bench_safe.err; summ $O/bench_safe.json
# D: conservative variants: kernel + net parity tests
env $SAFE timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_nets.py -m gpu -q --tb=short -p no:cacheprovider -n 4 --timeout 240 > $O/final_tests_safe.log 2>&1
stamp "D conservative kernel+nets tests rc=$?"; tail -8 $O/final_tests_safe.log
# E: smoke
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; stamp "E smoke rc=$?"; tail -3 $O/smoke.log
# F: ncu launch list of the bench command (small batch; shares, not absolutes)
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 1 --pairs-per-gpu 176 --no-cpu-baseline --no-roofline > $O/ncu_bench.log 2>&1
stamp "F ncu list rc=$?"; wc -l $O/launches.csv
# G: A/B benches
IIC_CONV_HALO_WGRAD=0 timeout 200 python bench.py --steps 5 --no-cpu-baseline > $O/bench_nohw.json 2> $O/bench_nohw.err; stamp "G1 bench halo wgrad off rc=$?"; summ $O/bench_nohw.json
IIC_CONV_HALO=0 timeout 200 python bench.py --steps 5 --no-cpu-baseline > $O/bench_nohalo.json 2> $O/bench_nohalo.err; stamp "G2 bench halo off rc=$?"; summ $O/bench_nohalo.json
# H: per-layer conv sweep, default variants
timeout 200 python tools/conv_sweep.py 1408 > $O/conv_sweep_final.txt 2>&1; stamp "H sweep rc=$?"; cat $O/conv_sweep_final.txt
# I: full ncu capture of the halo kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 5 -c 6 -o $O/prof_halo \
   python bench.py --steps 1 --warmup 1 --pairs-per-gpu 176 --no-cpu-baseline --no-roofline > $O/ncu_full.log 2>&1
stamp "I ncu full rc=$?"
