#!/bin/bash
# Round 2, GPU call A: new kernels (group walk, direct fast), native exchange on one GPU, bench, first profiles.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
echo "== targeted tests" 
timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_direct.py tests/test_gpu_sharded_local.py -q -m gpu -x --timeout 600 > gpurun_out/a_tests_new.log 2>&1
echo "rc=$?" >> gpurun_out/a_tests_new.log
tail -5 gpurun_out/a_tests_new.log
echo "== bench"
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "rc=$?" >> gpurun_out/a_bench.err
tail -c 1500 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
echo "== measure"
timeout 600 python tools/measure.py c1 c1fast c3s_17 c3s_17fast c4_20 c4_20fast c4_22fast c5_20 c5_20fast > gpurun_out/a_measure.log 2>&1
tail -12 gpurun_out/a_measure.log
echo "== ncu walk_group + direct_fast"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_group_kernel -c 1 -o gpurun_out/a_walk_group -f python tools/measure.py c4_20fast > gpurun_out/a_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:direct_fast_kernel -s 2 -c 1 -o gpurun_out/a_direct_fast -f python tools/measure.py c3s_17fast > gpurun_out/a_ncu2.log 2>&1
echo "== full suite"
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/a_tests_full.log 2>&1
echo "rc=$?" >> gpurun_out/a_tests_full.log
tail -8 gpurun_out/a_tests_full.log
