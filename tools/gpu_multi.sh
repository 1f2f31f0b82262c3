#!/bin/bash
# N-GPU run: NCCL parity test + bench under torch.distributed.run
N=${N:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/multi_test.log 2>&1; echo "multi test rc=$?"; tail -5 gpurun_out/multi_test.log
for n in $N; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  echo "bench n=$n rc=$?"; tail -2 gpurun_out/bench_n$n.err | cut -c1-300
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$n.json").read().strip().splitlines()[-1])
print("N=%d pairs/s %.0f ms/step %.1f e2e %.0f conv frac %.3f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
PY
done
