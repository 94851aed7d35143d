#!/bin/bash
# Round 2, GPU call J (1 GPU): the whole GPU suite, smoke, the bench line, its ncu launch list, C5 through the drop-in.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== full suite"
timeout 2400 python -m pytest tests -q -m gpu --timeout 1500 > gpurun_out/j_tests_full.log 2>&1
echo "rc=$?" >> gpurun_out/j_tests_full.log
tail -8 gpurun_out/j_tests_full.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/j_smoke.log; tail -2 gpurun_out/j_smoke.log
echo "== bench (both arms)"
timeout 900 python bench.py --impl reference > gpurun_out/j_bench_reference.json 2> gpurun_out/j_bench_reference.err; echo "rc=$?" >> gpurun_out/j_bench_reference.err
timeout 900 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "rc=$?" >> gpurun_out/j_bench.err
tail -c 300 gpurun_out/j_bench.json; tail -2 gpurun_out/j_bench.err; tail -c 600 gpurun_out/j_bench_reference.json
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/j_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/j_ncu_bench.log 2>&1
wc -l gpurun_out/j_bench_launches.csv
echo "== C5 through the drop-in: 10 vs 40 steps"
D=rebound_b200/_dropin
: > gpurun_out/j_c5_dropin.log
for dr in 1 0; do
  for st in 10 40; do
    echo "sheet steps=$st REBOUND_B200_DEVICE_RESOLVE=$dr" >> gpurun_out/j_c5_dropin.log
    REBOUND_B200_DEVICE_RESOLVE=$dr REBOUND_B200_RESOLVE_TRACE=1 timeout 900 $D/driver_dropin sheet /dev/null 2655 $st 2>&1 | tail -2 >> gpurun_out/j_c5_dropin.log
  done
done
for st in 10 40; do echo "sheet_hb steps=$st" >> gpurun_out/j_c5_dropin.log; timeout 900 $D/driver_dropin sheet_hb /dev/null 2655 $st 2>&1 | tail -1 >> gpurun_out/j_c5_dropin.log; done
cat gpurun_out/j_c5_dropin.log
