#!/bin/bash
# Round-2 GPU session N (2 GPUs): the driver's own invocation (default flags) under torchrun with the conditioned --verify,
# NCCL parity tests, strong scaling from a graph.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/n_t0
stamp() { echo "[$(( $(date +%s) - $(cat $O/n_t0) )) s] $*"; }
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
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/n_bench_default_n$N.json 2> $O/n_bench_default_n$N.err
stamp "1 bench.py --gpus $N (the driver's invocation) rc=$?"; tail -2 $O/n_bench_default_n$N.err | cut -c1-300; one $O/n_bench_default_n$N.json
timeout 300 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > $O/n_multi_tests.log 2>&1
stamp "2 tests/test_gpu_multi.py rc=$?"; tail -3 $O/n_multi_tests.log
timeout 200 $TR bench.py --gpus $N --config c4-strong --graph --steps 10 --no-roofline > $O/n_bench_strong_graph_n$N.json 2> $O/n_bench_strong_graph_n$N.err
stamp "3 strong c4, graph rc=$?"; tail -1 $O/n_bench_strong_graph_n$N.err | cut -c1-300; one $O/n_bench_strong_graph_n$N.json
timeout 200 $TR bench.py --gpus $N --config c3 --steps 10 --no-roofline > $O/n_bench_c3_n$N.json 2> $O/n_bench_c3_n$N.err
stamp "4 c3 rc=$?"; one $O/n_bench_c3_n$N.json
timeout 200 $TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $O/n_ref_n$N.json 2> $O/n_ref_n$N.err
stamp "5 reference arm rc=$?"; tail -1 $O/n_ref_n$N.json | cut -c1-300
