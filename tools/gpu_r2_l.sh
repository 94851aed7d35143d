#!/bin/bash
# Round 2, GPU call L (1 GPU): shipped group walk (paired evaluation + exact criterion): tests, smoke, shape A/B, bench line, ncu.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tests (direct, tree, sharded-on-one-GPU, simulation mirror)"
timeout 700 python -m pytest tests/test_gpu_direct.py tests/test_gpu_tree.py tests/test_gpu_sharded_local.py tests/test_gpu_simulation.py tests/test_gpu_random.py -q -m gpu --timeout 600 -x > gpurun_out/l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/l_tests.log; tail -4 gpurun_out/l_tests.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/l_smoke.log; tail -2 gpurun_out/l_smoke.log
echo "== shapes"
timeout 600 python tools/gw_ab.py -qstuv c4_20fast c4_22fast > gpurun_out/l_ab.jsonl 2> gpurun_out/l_ab.err
cut -c1-200 gpurun_out/l_ab.jsonl
timeout 300 python tools/gw_ab.py - c4_24fast c5_20fast >> gpurun_out/l_ab.jsonl 2>> gpurun_out/l_ab.err
tail -2 gpurun_out/l_ab.jsonl | cut -c1-200
echo "== bench"
timeout 900 python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "rc=$?" >> gpurun_out/l_bench.err
tail -c 300 gpurun_out/l_bench.json; tail -2 gpurun_out/l_bench.err
echo "== ncu --set full, group walk at 2^20 and 2^24"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:walk_group_kernel -c 1 -o gpurun_out/l_walk_group_20 -f python tools/measure.py c4_20fast > gpurun_out/l_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:walk_group_kernel -c 1 -o gpurun_out/l_walk_group_24 -f python tools/measure.py c4_24fast > gpurun_out/l_ncu2.log 2>&1
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/l_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/l_ncu_bench.log 2>&1
wc -l gpurun_out/l_bench_launches.csv
ls -la gpurun_out/l_* | head -20
