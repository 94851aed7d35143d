#!/bin/bash
# Round 2, GPU call P (N GPUs, N = $1): the headline workload sharded over N ranks with the shipped group walk.
N=${1:-8}; shift
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/p_bench_n$N.json 2> gpurun_out/p_bench_n$N.err
echo "rc=$?" >> gpurun_out/p_bench_n$N.err
python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/p_bench_n$N.json") if l.startswith("{")][0]
    print("n$N", d["n_gpus"], "%.4g" % d["value"], round(d["ms_per_step"], 2), d["kernel_ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "strict %.4g" % d.get("strict", {}).get("value", 0), d.get("exchange"))
    for k, b in d.get("configs", {}).items(): print("   ", k, "%.4g" % b["value"], round(b["ms_per_step"], 2), b.get("exchange"))
except Exception as e:
    print("n$N failed", e); print(open("gpurun_out/p_bench_n$N.err").read()[-1500:])
PY
