#!/bin/bash
# Round 2, GPU call I (8 GPUs): the headline workload sharded over 8 and 4 ranks.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/i_smi.txt 2>&1
run() { # n tag extra...
  n=$1; tag=$2; shift; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/i_bench_$tag.json 2> gpurun_out/i_bench_$tag.err
  echo "rc=$?" >> gpurun_out/i_bench_$tag.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/i_bench_$tag.json") if l.startswith("{")][0]
    print("$tag", d["n_gpus"], "%.4g" % d["value"], round(d["ms_per_step"], 2), d["kernel_ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "strict %.4g" % d.get("strict", {}).get("value", 0), d.get("exchange"))
    for k, b in d.get("configs", {}).items(): print("   ", k, "%.4g" % b["value"], round(b["ms_per_step"], 2), b.get("exchange"))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/i_bench_$tag.err").read()[-1500:])
PY
}
run 8 n8
run 8 n8_replicated --shard-build 0 --no-configs
run 4 n4 --no-configs
