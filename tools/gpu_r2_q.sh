#!/bin/bash
# Round 2, GPU call Q (1 GPU): the whole GPU suite and smoke on the final code.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 760 python -m pytest tests -q -m gpu --timeout 700 > gpurun_out/q_tests_full.log 2>&1
echo "rc=$?" >> gpurun_out/q_tests_full.log
tail -8 gpurun_out/q_tests_full.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/q_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/q_smoke.log; tail -2 gpurun_out/q_smoke.log
