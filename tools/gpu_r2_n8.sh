#!/bin/bash
# Round-2 GPU session N8 (8 GPUs): the driver's invocation at N = 8 on the final build, strong scaling from a graph.
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/n8_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/n8_t0) )) s] $*"; }
one() {
python - "$1" <<'PY'
import json, sys
try:
  d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
  print("  %s N=%d pairs/gpu %d: %.0f pairs/s, %.2f ms/step, e2e %.0f, parity_ok=%s %s" % (d["config"]["name"], d["n_gpus"], d["config"]["pairs_per_gpu"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("parity_ok"), {k: v for k, v in (d.get("verify") or {}).items() if k in ("loss", "loss_emulated", "grad_sq_rel_err", "ranks_agree", "why")}))
  t = d.get("allreduce_timeline")
  if t: print("    all-reduce: backward ends %.2f ms, gradients ready %.2f ms, exposed %.2f ms, busy %.2f ms" % (t["backward_end_ms"], t["gradients_ready_ms"], t["exposed_ms"], t["allreduce_busy_ms"]))
except Exception as e:
  print("  no json:", e)
PY
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/n8_bench_default.json 2> $O/n8_bench_default.err
stamp "1 bench.py --gpus $N (the driver's invocation) rc=$?"; tail -2 $O/n8_bench_default.err | cut -c1-300; one $O/n8_bench_default.json
timeout 200 $TR bench.py --gpus $N --config c4-strong --graph --steps 20 --no-roofline > $O/n8_bench_strong_graph.json 2> $O/n8_bench_strong_graph.err
stamp "2 strong c4 (88 pairs/GPU), graph rc=$?"; tail -1 $O/n8_bench_strong_graph.err | cut -c1-300; one $O/n8_bench_strong_graph.json
timeout 200 $TR bench.py --gpus $N --config c4-strong --steps 20 --no-roofline > $O/n8_bench_strong.json 2> $O/n8_bench_strong.err
stamp "3 strong c4, eager rc=$?"; one $O/n8_bench_strong.json
