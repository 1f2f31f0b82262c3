#!/bin/bash
# bench + launch list (+ optional full ncu capture of the conv kernel) on the B200 box
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
PAIRS=${PAIRS:-704}
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 --pairs-per-gpu $PAIRS > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${NCU_LIST:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 1 --pairs-per-gpu ${NCU_PAIRS:-176} --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
fi
if [ "${NCU_FULL:-0}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s ${NCU_SKIP:-150} -c ${NCU_COUNT:-4} -o gpurun_out/prof_conv \
     python bench.py --steps 1 --warmup 1 --pairs-per-gpu ${NCU_PAIRS:-176} --no-cpu-baseline --no-roofline > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
