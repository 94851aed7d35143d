#!/bin/bash
# Round 2, GPU call C: group-walk shapes (A/B), group handle + drop-in over devices, lazy heartbeat on the real engine.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== group walk variants"
: > gpurun_out/c_variants.log
for v in a b c d e f g; do
  echo "variant $v" >> gpurun_out/c_variants.log
  REBOUND_B200_GW_VARIANT=$v timeout 300 python tools/measure.py c4_20fast c4_22fast c5_20fast 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['case'], 'walk_ms', d['kernel_ms_per_step'].get('treewalk'), 'groups', d['walk_stats']['groups'], 'entries', d['walk_stats']['group_entries'])
" >> gpurun_out/c_variants.log
done
cat gpurun_out/c_variants.log
echo "== tests"
timeout 1800 python -m pytest tests/test_gpu_group.py tests/test_gpu_tree.py "tests/test_gpu_hostlogic.py::test_host_side_call_sequences_on_the_real_engine" -q -m gpu --timeout 900 > gpurun_out/c_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c_tests.log
tail -15 gpurun_out/c_tests.log
