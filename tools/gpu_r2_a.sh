#!/bin/bash
# Round-2 GPU session A (one box, ~15 min): (1) whole GPU suite incl. the variants written after round 1's last GPU
# session and this round's new kernels (tf32 / 3xTF32 convolutions, step functions, batch assembly), (2) end-to-end
# precision of every mode on the well-conditioned fixture, (3) the official bench line, (4) not-yet-measured switches,
# (5) the other BASELINE configs, (6) ncu launch list + DRAM traffic of the dominant conv kernel at the bench batch.
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/a_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/a_t0) )) s] $*"; }
summ() {
python - "$1" <<'PY'
import json, sys
try:
  d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
except Exception as e:
  print("no bench json:", e); sys.exit(0)
r = d.get("roofline", {})
print("%s pairs/s %.0f  ms/step %.2f  e2e %.0f  loss %.3e launches %d conv TF/s %.0f frac %.3f by_kind %s" % (d["config"]["name"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"], d["gpu_launches"], r.get("achieved", 0), r.get("frac", 0), {k: (round(v["tflops"]), round(v["ms_per_step"], 2)) for k, v in r.get("by_kind", {}).items()}))
print("  other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])})
for m, e in d.get("precision_modes", {}).items():
  rr = e.get("roofline", {})
  print("  mode %s: pairs/s %.0f ms/step %.1f conv TF/s %.0f frac %.3f" % (m, e["value"], e["ms_per_step"], rr.get("achieved", 0), rr.get("frac", 0)))
print("  clocks:", d.get("clocks"), "cpu:", (d.get("cpu_baseline") or {}).get("value"))
PY
}
IIC_RUN_UNVALIDATED=1 timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -n 4 --timeout 300 > $O/a_tests.log 2>&1
stamp "1 suite (incl. unvalidated + new) rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|crashed" $O/a_tests.log | tail -40
timeout 60 tools/umma_sw64_probe > $O/a_sw64_probe.txt 2>&1; stamp "1b SWIZZLE_64B overlapped-operand probe rc=$?"; cat $O/a_sw64_probe.txt
timeout 200 python tools/precision_probe.py --sz 32 --pairs 64 --steps 40 --modes bf16,tf32,tf32x3,fp32 --fp64 > $O/a_prec32.json 2> $O/a_prec32.err
stamp "2a precision 32x32 rc=$?"; tail -2 $O/a_prec32.err; python - <<'PY'
import json
try:
  d = json.load(open("gpurun_out/a_prec32.json"))
  for k, v in d.items():
    if isinstance(v, dict) and "grad_rel_l2_total" in v:
      print(k, {a: (round(b, 6) if isinstance(b, float) else b) for a, b in v.items() if a != "worst"})
    elif k.startswith("trajectory_rel"):
      print(k, v)
  print("oracle loss", d.get("oracle_loss"))
except Exception as e:
  print("no precision json", e)
PY
timeout 300 python bench.py --arena --rgb-input --also tf32x3 > $O/a_bench.json 2> $O/a_bench.err; stamp "3 bench default rc=$?"; tail -2 $O/a_bench.err; summ $O/a_bench.json
timeout 150 python bench.py --steps 5 --no-cpu-baseline > $O/a_bench_r1path.json 2> $O/a_bench_r1path.err; stamp "4b bench, round-1 path (no arena, grey input) rc=$?"; summ $O/a_bench_r1path.json
for c in c2 c3 c4-strong c5; do
  timeout 150 python bench.py --arena --rgb-input --config $c --steps 5 --no-cpu-baseline --also '' > $O/a_bench_$c.json 2> $O/a_bench_$c.err; stamp "5 bench $c rc=$?"; tail -1 $O/a_bench_$c.err; summ $O/a_bench_$c.json
done
timeout 150 python bench.py --arena --rgb-input --graph --steps 10 --no-cpu-baseline --no-roofline --also '' > $O/a_bench_graph.json 2> $O/a_bench_graph.err; stamp "4a bench c4, CUDA graph rc=$?"; tail -2 $O/a_bench_graph.err; summ $O/a_bench_graph.json
timeout 150 python bench.py --arena --rgb-input --graph --config c2 --steps 10 --no-cpu-baseline --no-roofline --also '' > $O/a_bench_graph_c2.json 2> $O/a_bench_graph_c2.err; stamp "4a bench c2, CUDA graph rc=$?"; tail -2 $O/a_bench_graph_c2.err; summ $O/a_bench_graph_c2.json
timeout 150 python bench.py --arena --rgb-input --graph --pairs-per-gpu 88 --steps 10 --no-cpu-baseline --no-roofline --also '' > $O/a_bench_graph_88.json 2> $O/a_bench_graph_88.err; stamp "4a bench c4 88 pairs, CUDA graph rc=$?"; summ $O/a_bench_graph_88.json
timeout 150 python bench.py --arena --rgb-input --pairs-per-gpu 88 --steps 10 --no-cpu-baseline --no-roofline --also '' > $O/a_bench_eager_88.json 2> $O/a_bench_eager_88.err; stamp "4a bench c4 88 pairs, eager rc=$?"; summ $O/a_bench_eager_88.json
timeout 100 python tools/conv_sweep.py 352 tf32x3 > $O/a_conv_sweep_tf32x3.txt 2>&1; stamp "7b sweep tf32x3 rc=$?"; tail -12 $O/a_conv_sweep_tf32x3.txt
IIC_SEG_JOINT_TC=1 timeout 100 python tools/seg_step.py 15 A > $O/a_seg_tc.json 2>&1; stamp "8 seg step, tensor-core joint rc=$?"; tail -1 $O/a_seg_tc.json
timeout 100 python tools/seg_step.py 15 A > $O/a_seg_simt.json 2>&1; stamp "8b seg step, SIMT joint rc=$?"; tail -1 $O/a_seg_simt.json
for v in IIC_BN_BITMASK=1 IIC_CONV_HALO_STATS=1 "IIC_STEM_BWD_FUSED=1 IIC_STEM_BWD_V2=1" ; do
  f=$(echo "$v" | tr ' =' '__')
  env $v timeout 120 python bench.py --arena --rgb-input --steps 5 --no-cpu-baseline --also '' > $O/a_bench_$f.json 2> $O/a_bench_$f.err; stamp "4 bench $v rc=$?"; summ $O/a_bench_$f.json
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/a_launches.csv \
   python bench.py --arena --rgb-input --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --also '' > $O/a_ncu_list.log 2>&1
stamp "6a ncu launch list rc=$?"; python tools/ncu_launch_table.py $O/a_launches.csv "launch list" "" 2>/dev/null | head -40
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel|conv_halo" -s 40 -c 12 -o $O/a_prof_conv \
   python bench.py --arena --rgb-input --steps 1 --warmup 0 --no-cpu-baseline --no-roofline --also '' > $O/a_ncu_conv.log 2>&1
stamp "6b ncu conv kernels rc=$?"
timeout 200 python tools/precision_probe.py --sz 96 --pairs 32 --steps 0 --modes bf16,tf32,tf32x3 > $O/a_prec96.json 2> $O/a_prec96.err
stamp "2b precision 96x96 rc=$?"; tail -2 $O/a_prec96.err; grep -E "grad_rel_l2_total|grad_cos_min|loss_rel|out_max" $O/a_prec96.json | head -20
timeout 100 python tools/conv_sweep.py 1408 > $O/a_conv_sweep.txt 2>&1; stamp "7 sweep rc=$?"; tail -12 $O/a_conv_sweep.txt
timeout 100 python tools/conv_sweep.py 352 tf32 > $O/a_conv_sweep_tf32.txt 2>&1; stamp "7c sweep tf32 rc=$?"; tail -12 $O/a_conv_sweep_tf32.txt
