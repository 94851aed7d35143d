#!/bin/bash
# Round 2, GPU call E (2 GPUs): NCCL transport -- multi-process tests, group handle on distinct devices, bench at N=2.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/e_smi.txt 2>&1
echo "== tests (2 GPUs)"
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -q -m gpu --timeout 900 > gpurun_out/e_tests.log 2>&1
echo "rc=$?" >> gpurun_out/e_tests.log
tail -12 gpurun_out/e_tests.log
echo "== bench N=2 (sharded build)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e_bench2.json 2> gpurun_out/e_bench2.err
echo "rc=$?" >> gpurun_out/e_bench2.err
tail -c 1200 gpurun_out/e_bench2.json; tail -3 gpurun_out/e_bench2.err
echo "== bench N=2 (replicated build)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --shard-build 0 --no-configs > gpurun_out/e_bench2_repl.json 2> gpurun_out/e_bench2_repl.err
echo "rc=$?" >> gpurun_out/e_bench2_repl.err
tail -c 600 gpurun_out/e_bench2_repl.json; tail -3 gpurun_out/e_bench2_repl.err
echo "== drop-in on 2 GPUs: C4-like disc through reb_simulation_steps"
D=rebound_b200/_dropin
for devs in 0 0,1; do
  echo "REBOUND_B200_DEVICES=$devs" >> gpurun_out/e_dropin.log
  REBOUND_B200_DEVICES=$devs timeout 600 $D/driver_dropin disc gpurun_out/e_disc_$devs.bin 4194303 5 >> gpurun_out/e_dropin.log 2>&1
done
cmp gpurun_out/e_disc_0.bin gpurun_out/e_disc_0,1.bin && echo "disc 2^22: 1 GPU == 2 GPUs (bitwise)" >> gpurun_out/e_dropin.log
rm -f gpurun_out/e_disc_*.bin
cat gpurun_out/e_dropin.log
