#!/bin/bash
# Round 2, GPU call H (2 GPUs): NCCL path after the padded all-gather of ragged ranges and the open-boundary probe.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 600 > gpurun_out/h_tests.log 2>&1
echo "rc=$?" >> gpurun_out/h_tests.log
tail -6 gpurun_out/h_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-configs > gpurun_out/h_bench2.json 2> gpurun_out/h_bench2.err
echo "rc=$?" >> gpurun_out/h_bench2.err
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/h_bench2.json") if l.startswith("{")][0]
print(d["n_gpus"], "%.4g" % d["value"], round(d["ms_per_step"], 2), d["kernel_ms_per_step"], d["exchange"])
PY
tail -2 gpurun_out/h_bench2.err
