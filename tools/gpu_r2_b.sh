#!/bin/bash
# Round 2, GPU call B: group walk with abort, sharded tree build / group handle / host logic on one GPU, profiles, bench.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== targeted tests"
timeout 1500 python -m pytest tests/test_gpu_tree.py tests/test_gpu_sharded_local.py tests/test_gpu_group.py -q -m gpu --timeout 600 -x > gpurun_out/b_tests_new.log 2>&1
echo "rc=$?" >> gpurun_out/b_tests_new.log
tail -15 gpurun_out/b_tests_new.log
echo "== measure"
timeout 600 python tools/measure.py d17 d17fast c4_20fast c4_22fast c4_24fast c5_20fast > gpurun_out/b_measure.log 2>&1
tail -8 gpurun_out/b_measure.log
echo "== ncu walk_group + direct_fast"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_group_kernel -c 1 -o gpurun_out/b_walk_group -f python tools/measure.py c4_20fast > gpurun_out/b_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:direct_fast_kernel -s 1 -c 1 -o gpurun_out/b_direct_fast -f python tools/measure.py d17fast > gpurun_out/b_ncu2.log 2>&1
echo "== bench"
timeout 900 python bench.py > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "rc=$?" >> gpurun_out/b_bench.err
tail -c 600 gpurun_out/b_bench.json; tail -3 gpurun_out/b_bench.err
echo "== rest of the new tests"
timeout 1800 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_hostlogic.py -q -m gpu --timeout 1500 > gpurun_out/b_tests_rest.log 2>&1
echo "rc=$?" >> gpurun_out/b_tests_rest.log
tail -12 gpurun_out/b_tests_rest.log
