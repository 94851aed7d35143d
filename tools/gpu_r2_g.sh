#!/bin/bash
# Round 2, GPU call G (1 GPU): key-range walk of sharded runs (LOCAL transport), C5 through the drop-in in steady state,
# C1 FAST with two particles per lane, DRAM traffic of the headline kernels at full size.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== sharded tests on one GPU"
timeout 1200 python -m pytest tests/test_gpu_sharded_local.py "tests/test_gpu_group.py::test_group_handle_steps_host_bitwise" "tests/test_gpu_group.py::test_group_handle_call_by_call" -q -m gpu --timeout 600 > gpurun_out/g_tests.log 2>&1
echo "rc=$?" >> gpurun_out/g_tests.log
tail -8 gpurun_out/g_tests.log
echo "== C5 through the drop-in: 10 vs 40 steps"
D=rebound_b200/_dropin
: > gpurun_out/g_c5_dropin.log
for dr in 1 0; do
  for st in 10 40; do
    echo "sheet steps=$st REBOUND_B200_DEVICE_RESOLVE=$dr" >> gpurun_out/g_c5_dropin.log
    REBOUND_B200_DEVICE_RESOLVE=$dr REBOUND_B200_RESOLVE_TRACE=1 timeout 900 $D/driver_dropin sheet /dev/null 2655 $st 2>&1 | tail -4 >> gpurun_out/g_c5_dropin.log
  done
done
echo "sheet_hb steps=40 (heartbeat installed, lazy host copy)" >> gpurun_out/g_c5_dropin.log
timeout 900 $D/driver_dropin sheet_hb /dev/null 2655 40 2>&1 | tail -1 >> gpurun_out/g_c5_dropin.log
echo "reference (OpenMP off: serial build), sheet steps=2 at root 939 (N~2^17)" >> gpurun_out/g_c5_dropin.log
timeout 900 $D/driver_ref sheet /dev/null 939 2 2>&1 | tail -1 >> gpurun_out/g_c5_dropin.log
cat gpurun_out/g_c5_dropin.log
echo "== C1 FAST: one vs two particles per lane"
for ipt in 1 2; do REBOUND_B200_FAST_IPT=$ipt timeout 120 python tools/measure.py c1fast 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print('ipt', $ipt, d['case'], d['kernel_ms_per_step'], '%.4g' % d['interactions_per_s'])
    except Exception: pass
"; done | tee gpurun_out/g_c1_ipt.log
echo "== ncu: DRAM traffic at full size (group walk 2^24, strict walk 2^24)"
timeout 900 ncu --set full --clock-control none -k regex:walk_group_kernel -c 1 -o gpurun_out/g_walk_group_24 -f python tools/measure.py c4_24fast > gpurun_out/g_ncu1.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:walk_rec_kernel -c 1 --csv --log-file gpurun_out/g_walk_rec_24.csv python tools/measure.py c4_24 > gpurun_out/g_ncu2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/g_launches_c4_22fast.csv python tools/measure.py c4_22fast > gpurun_out/g_ncu3.log 2>&1
tail -3 gpurun_out/g_walk_rec_24.csv
