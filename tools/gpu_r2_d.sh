#!/bin/bash
# Round 2, GPU call D: exact device-side resolve + lazy heartbeat through the drop-in, final group walk, C5 through the drop-in.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tests"
timeout 2400 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_hostlogic.py tests/test_gpu_tree.py tests/test_gpu_resolve.py tests/test_gpu_group.py -q -m gpu --timeout 1200 > gpurun_out/d_tests.log 2>&1
echo "rc=$?" >> gpurun_out/d_tests.log
tail -15 gpurun_out/d_tests.log
echo "== measure"
timeout 600 python tools/measure.py c4_20fast c4_22fast c4_24fast c5_20fast c5_20 > gpurun_out/d_measure.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/d_measure.log"):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["case"], d["kernel_ms_per_step"], d["walk_stats"]["groups"], d["walk_stats"]["group_entries"])
PY
echo "== C5 through the drop-in (reb_simulation_steps, N ~ 2^20, 10 steps)"
D=rebound_b200/_dropin
for scen in sheet sheet_hb; do
  for dr in 1 0; do
    echo "scenario $scen REBOUND_B200_DEVICE_RESOLVE=$dr" >> gpurun_out/d_c5_dropin.log
    REBOUND_B200_DEVICE_RESOLVE=$dr REBOUND_B200_PIN=0 timeout 600 $D/driver_dropin $scen gpurun_out/d_c5_${scen}_$dr.bin 2655 10 >> gpurun_out/d_c5_dropin.log 2>&1
  done
done
cmp gpurun_out/d_c5_sheet_1.bin gpurun_out/d_c5_sheet_0.bin && echo "sheet: device resolve == host resolve (bitwise)" >> gpurun_out/d_c5_dropin.log
cmp gpurun_out/d_c5_sheet_hb_1.bin gpurun_out/d_c5_sheet_hb_0.bin && echo "sheet_hb: device resolve == host resolve (bitwise)" >> gpurun_out/d_c5_dropin.log
rm -f gpurun_out/d_c5_*.bin
cat gpurun_out/d_c5_dropin.log
echo "== ncu walk_group (final)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_group_kernel -c 1 -o gpurun_out/d_walk_group -f python tools/measure.py c4_20fast > gpurun_out/d_ncu1.log 2>&1
echo "== bench"
timeout 900 python bench.py > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "rc=$?" >> gpurun_out/d_bench.err
tail -c 400 gpurun_out/d_bench.json; tail -3 gpurun_out/d_bench.err
