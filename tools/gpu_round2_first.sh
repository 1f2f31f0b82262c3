#!/bin/bash
# First GPU call of round 2: (1) the suite as the driver runs it (serial), (2) the tests of everything written after the
# last GPU session of round 1 (IIC_RUN_UNVALIDATED=1: bn_bitmask, stem_bwd_v2, evaluation kernels), (3) the official bench
# line, (4) one bench per not-yet-measured switch, (5) ncu: launch list + full captures of the kernels that decide the
# next steps (both versions of the fused stem backward, the bit-mask BatchNorm backward, the N=128 two-tile kernel).
#   gpurun --timeout 900 -- bash tools/gpu_round2_first.sh
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/r2_t0
[ -f iic_b200/lib/libiic_b200.so ] || python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
stamp() { echo "[$(( $(date +%s) - $(cat $O/r2_t0) )) s] $*"; }
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
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2_tests_serial.log 2>&1
stamp "1 suite, serial, as the driver runs it rc=$?"; tail -3 $O/r2_tests_serial.log
IIC_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests -m "gpu and unvalidated" -q --tb=short -p no:cacheprovider -n 4 --timeout 200 > $O/r2_tests_unvalidated.log 2>&1
stamp "2 unvalidated variants rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/r2_tests_unvalidated.log | tail -30
timeout 300 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; stamp "3 bench default rc=$?"; tail -2 $O/r2_bench.err; summ $O/r2_bench.json
for v in IIC_BN_BITMASK=1 IIC_CONV_HALO_STATS=1 IIC_STEM_BWD_FUSED=1 "IIC_STEM_BWD_FUSED=1 IIC_STEM_BWD_V2=1"; do
  f=$(echo "$v" | tr ' =' '__')
  env $v timeout 200 python bench.py --steps 5 --no-cpu-baseline > $O/r2_bench_$f.json 2> $O/r2_bench_$f.err; stamp "4 bench $v rc=$?"; summ $O/r2_bench_$f.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches.csv \
   python bench.py --steps 1 --warmup 1 --pairs-per-gpu 176 --no-cpu-baseline --no-roofline > $O/r2_ncu_list.log 2>&1
stamp "5a ncu launch list rc=$?"; python tools/ncu_launch_table.py $O/r2_launches.csv "launch list" "" 2>/dev/null | head -30
IIC_STEM_BWD_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stem_bwd" -c 4 -o $O/r2_prof_stem_v1 \
   python bench.py --steps 1 --warmup 0 --pairs-per-gpu 176 --no-cpu-baseline --no-roofline > $O/r2_ncu_stem_v1.log 2>&1
stamp "5b ncu stem bwd v1 rc=$?"
IIC_STEM_BWD_FUSED=1 IIC_STEM_BWD_V2=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stem_bwd" -c 4 -o $O/r2_prof_stem_v2 \
   python bench.py --steps 1 --warmup 0 --pairs-per-gpu 176 --no-cpu-baseline --no-roofline > $O/r2_ncu_stem_v2.log 2>&1
stamp "5c ncu stem bwd v2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel" -s 30 -c 6 -o $O/r2_prof_tc2 \
   python bench.py --steps 1 --warmup 0 --pairs-per-gpu 176 --no-cpu-baseline --no-roofline > $O/r2_ncu_tc2.log 2>&1
stamp "5d ncu tc2 rc=$?"
timeout 200 python tools/conv_sweep.py 1408 > $O/r2_conv_sweep.txt 2>&1; stamp "6 sweep rc=$?"; cat $O/r2_conv_sweep.txt
timeout 200 python tools/seg_step.py 15 A > $O/r2_seg_A.json 2>&1; stamp "7 seg step rc=$?"; cat $O/r2_seg_A.json
