#!/bin/bash
# experiment round: descriptor probe, BN kernel tests, net parity, bench with / without the merged BN launches
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
[ -x tools/umma_shift_probe ] && { timeout 120 tools/umma_shift_probe > gpurun_out/umma_shift_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/umma_shift_probe.txt; }
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "bn or stats" > gpurun_out/exp_tests.log 2>&1; echo "bn tests rc=$?"; tail -15 gpurun_out/exp_tests.log
timeout 600 python -m pytest tests/test_gpu_parity_nets.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/exp_nets.log 2>&1; echo "nets rc=$?"; tail -8 gpurun_out/exp_nets.log
summ() {
python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d.get("roofline", {})
print("pairs/s %.0f  ms/step %.1f  e2e %.0f  launches %d conv TF/s %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], r.get("achieved", 0)))
print("other:", {k: round(v["ms_per_step"], 2) for k, v in sorted(r.get("other_kernels_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])})
PY
}
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_exp_on.json 2> gpurun_out/bench_exp_on.err; tail -2 gpurun_out/bench_exp_on.err; summ gpurun_out/bench_exp_on.json
IIC_BN_MERGED=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_exp_off.json 2> gpurun_out/bench_exp_off.err; tail -2 gpurun_out/bench_exp_off.err; summ gpurun_out/bench_exp_off.json
timeout 300 python tools/bn_sweep.py > gpurun_out/bn_sweep.txt 2>&1; cat gpurun_out/bn_sweep.txt
