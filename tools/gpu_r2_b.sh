#!/bin/bash
# Round-2 GPU session B (multi-GPU, gpurun --gpus N): NCCL parity tests, weak and strong scaling with the overlapped
# in-place gradient all-reduce against round 1's cat + all-reduce path, --verify (N ranks == one-device emulation).
#   gpurun --gpus 2 --timeout 900 -- bash tools/gpu_r2_b.sh 2
N=${1:-2}
QUICK=${2:-}
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/b_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/b_t0) )) s] $*"; }
one() {
python - "$1" <<'PY'
import json, sys
try:
  d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
  print("  %s N=%d pairs/gpu %d: %.0f pairs/s, %.2f ms/step, e2e %.0f, parity_ok=%s %s" % (d["config"]["name"], d["n_gpus"], d["config"]["pairs_per_gpu"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("parity_ok"), {k: v for k, v in (d.get("verify") or {}).items() if k in ("loss", "loss_emulated", "grad_sq_rel_err", "ranks_agree")}))
  t = d.get("allreduce_timeline")
  if t: print("    all-reduce: backward ends %.2f ms, gradients ready %.2f ms, exposed %.2f ms, busy %.2f ms, buckets %s" % (t["backward_end_ms"], t["gradients_ready_ms"], t["exposed_ms"], t["allreduce_busy_ms"], [(r["bucket"], round(r["start_ms"], 1), round(r["end_ms"], 1)) for r in t["buckets"]]))
except Exception as e:
  print("  no json:", e)
PY
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ -z "$QUICK" ]; then
IIC_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > $O/b_multi_tests.log 2>&1
stamp "1 tests/test_gpu_multi.py rc=$?"; tail -5 $O/b_multi_tests.log
fi
timeout 200 $TR bench.py --gpus $N --arena --verify --steps 10 --no-roofline > $O/b_bench_arena_n$N.json 2> $O/b_bench_arena_n$N.err
stamp "2 weak c4, arena + overlap + verify rc=$?"; tail -2 $O/b_bench_arena_n$N.err; one $O/b_bench_arena_n$N.json
timeout 200 $TR bench.py --gpus $N --no-arena --verify --steps 10 --no-roofline > $O/b_bench_r1_n$N.json 2> $O/b_bench_r1_n$N.err
stamp "3 weak c4, round-1 gradient path rc=$?"; one $O/b_bench_r1_n$N.json
timeout 200 $TR bench.py --gpus $N --arena --config c4-strong --steps 10 --no-roofline > $O/b_bench_strong_arena_n$N.json 2> $O/b_bench_strong_arena_n$N.err
stamp "4 strong c4 (704 global), arena rc=$?"; one $O/b_bench_strong_arena_n$N.json
timeout 200 $TR bench.py --gpus $N --no-arena --config c4-strong --steps 10 --no-roofline > $O/b_bench_strong_r1_n$N.json 2> $O/b_bench_strong_r1_n$N.err
stamp "5 strong c4 (704 global), round-1 path rc=$?"; one $O/b_bench_strong_r1_n$N.json
timeout 200 $TR bench.py --gpus $N --arena --graph --config c4-strong --steps 10 --no-roofline > $O/b_bench_strong_graph_n$N.json 2> $O/b_bench_strong_graph_n$N.err
stamp "5b strong c4 (704 global), arena + CUDA graph rc=$?"; tail -2 $O/b_bench_strong_graph_n$N.err; one $O/b_bench_strong_graph_n$N.json
if [ -z "$QUICK" ]; then
timeout 200 $TR bench.py --gpus $N --arena --config c3 --steps 10 --no-roofline > $O/b_bench_c3_n$N.json 2> $O/b_bench_c3_n$N.err
stamp "6 c3 (660 global) rc=$?"; one $O/b_bench_c3_n$N.json
fi
timeout 200 $TR bench.py --gpus $N --impl reference --steps 2 > $O/b_ref_n$N.json 2> $O/b_ref_n$N.err
stamp "7 reference arm under torchrun rc=$?"; python -c "
import json; d=json.loads(open('$O/b_ref_n$N.json').read().strip().splitlines()[-1]); print('  reference: %.1f pairs/s on %d threads' % (d['value'], d['cpu_baseline']['cores']))" 2>/dev/null
