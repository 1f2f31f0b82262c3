#!/bin/bash
# Round-2 GPU session C (1 GPU): validation of the fixes after session A (tf32 wgrad through shared-memory transposes,
# forward segmentation joint on bf16 three-term operands, re-set test tolerances), end-to-end precision of every mode,
# the official bench line with its precision modes, ncu evidence for the new kernels, smoke().
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/c_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/c_t0) )) s] $*"; }
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
print("  clocks:", d.get("clocks"), "cpu:", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
PY
}
IIC_RUN_UNVALIDATED=1 timeout 500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -n 4 --timeout 400 > $O/c_tests.log 2>&1
stamp "1 suite (incl. unvalidated) rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|crashed" $O/c_tests.log | tail -40; grep -E "^E  " $O/c_tests.log | sort | uniq -c | sort -rn | head -25
timeout 60 tools/umma_bf16_mn_probe > $O/c_bf16_mn_probe.txt 2>&1; stamp "1b bf16 MN-major overlapped probe rc=$?"; cat $O/c_bf16_mn_probe.txt
timeout 500 python tools/precision_probe.py --sz 32 --pairs 64 --steps 40 --modes bf16,tf32,tf32x3,fp32 > $O/c_prec32.json 2> $O/c_prec32.err
stamp "2 precision 32x32 rc=$?"; tail -2 $O/c_prec32.err; python - <<'PY'
import json
try:
  d = json.load(open("gpurun_out/c_prec32.json"))
  for k, v in d.items():
    if isinstance(v, dict) and "grad_rel_l2_total" in v:
      print(k, {a: (round(b, 6) if isinstance(b, float) else b) for a, b in v.items() if a != "worst"})
    elif k.startswith("trajectory_rel"):
      print(k, v)
  print("oracle loss", d.get("oracle_loss"), "trajectory last:", {m: t[-1] for m, t in d.get("trajectory", {}).items()})
except Exception as e:
  print("no precision json", e)
PY
timeout 400 python bench.py --also tf32x3,tf32 > $O/c_bench.json 2> $O/c_bench.err; stamp "3 bench default (+ tf32x3, tf32 modes) rc=$?"; tail -2 $O/c_bench.err; summ $O/c_bench.json
IIC_WGRAD_FUSED=1 timeout 150 python bench.py --steps 10 --no-cpu-baseline > $O/c_bench_wgrad_fused.json 2> $O/c_bench_wgrad_fused.err; stamp "3b bench, wgrad writes the torch layout rc=$?"; summ $O/c_bench_wgrad_fused.json
timeout 200 python __graft_entry__.py smoke > $O/c_smoke.log 2>&1; stamp "4 smoke rc=$?"; tail -6 $O/c_smoke.log
IIC_SMOKE_MODES=fp32,tf32x3,bf16 timeout 200 python __graft_entry__.py smoke > $O/c_smoke_all.log 2>&1; stamp "4b smoke incl. tf32x3 rc=$?"; tail -6 $O/c_smoke_all.log
timeout 100 python tools/seg_step.py 15 A > $O/c_seg_default.json 2>&1; stamp "5 seg step, default (tensor-core backward, SIMT joint) rc=$?"; tail -1 $O/c_seg_default.json
IIC_SEG_JOINT_TC=1 timeout 100 python tools/seg_step.py 15 A > $O/c_seg_tc.json 2>&1; stamp "5b seg step, + bf16 tensor-core joint rc=$?"; tail -1 $O/c_seg_tc.json
IIC_SEG_CORR_TC=0 timeout 100 python tools/seg_step.py 15 A > $O/c_seg_simt.json 2>&1; stamp "5c seg step, all SIMT rc=$?"; tail -1 $O/c_seg_simt.json
timeout 150 python bench.py --config c5 --steps 5 --no-cpu-baseline > $O/c_bench_c5.json 2> $O/c_bench_c5.err; stamp "6 bench c5 default rc=$?"; summ $O/c_bench_c5.json
IIC_SEG_JOINT_TC=1 timeout 150 python bench.py --config c5 --steps 5 --no-cpu-baseline > $O/c_bench_c5_tc.json 2> $O/c_bench_c5_tc.err; stamp "6b bench c5 + tensor-core joint rc=$?"; summ $O/c_bench_c5_tc.json
timeout 100 python tools/conv_sweep.py 352 tf32x3 > $O/c_conv_sweep_tf32x3.txt 2>&1; stamp "7 sweep tf32x3 rc=$?"; tail -12 $O/c_conv_sweep_tf32x3.txt
timeout 100 python tools/conv_sweep.py 352 tf32 > $O/c_conv_sweep_tf32.txt 2>&1; stamp "7b sweep tf32 rc=$?"; tail -3 $O/c_conv_sweep_tf32.txt
timeout 100 python bench.py --pairs-per-gpu 88 --steps 10 --no-cpu-baseline > $O/c_bench_88.json 2> $O/c_bench_88.err; stamp "7c bench 88 pairs (breakdown) rc=$?"; summ $O/c_bench_88.json
timeout 250 ncu --set full --clock-control none --import-source on -k regex:"conv_tf32_kernel" -s 30 -c 9 -o $O/c_prof_tf32 \
   python bench.py --precision tf32x3 --pairs-per-gpu 176 --steps 1 --warmup 0 --no-cpu-baseline --no-roofline > $O/c_ncu_tf32.log 2>&1
stamp "8 ncu conv_tf32 rc=$?"
IIC_SEG_JOINT_TC=1 timeout 250 ncu --set full --clock-control none --import-source on -k regex:"seg_joint_tc_kernel|seg_corr_tc_kernel|seg_joint_kernel" -c 6 -o $O/c_prof_seg \
   python tools/seg_step.py 15 A > $O/c_ncu_seg.log 2>&1
stamp "8b ncu seg tc rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c_launches.csv \
   python bench.py --steps 1 --warmup 1 --pairs-per-gpu 176 --no-cpu-baseline --no-roofline > $O/c_ncu_list.log 2>&1
stamp "9 ncu launch list (176 pairs) rc=$?"; python tools/ncu_launch_table.py $O/c_launches.csv "launch list" "" 2>/dev/null | head -14
