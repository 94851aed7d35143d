#!/bin/bash
# Round 2, GPU call M (1 GPU): the tests call L did not reach (it stopped at the first failure), with the walk variants fixed.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_sharded_local.py tests/test_gpu_simulation.py tests/test_gpu_random.py tests/test_gpu_group.py -q -m gpu --timeout 600 > gpurun_out/m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/m_tests.log; tail -6 gpurun_out/m_tests.log
