#!/bin/bash
# Round 2, GPU call K (1 GPU): group-walk variants (paired evaluation, exact group criterion, cheaper pair term): parity
# class tests, then A/B timings at 2^20 / 2^22, then 2^24 for the default and the candidates.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== FAST tests, default kernels (new 16-instruction pair term in the direct kernels and the default group walk)"
timeout 600 python -m pytest tests/test_gpu_tree.py tests/test_gpu_direct.py tests/test_gpu_sharded_local.py -q -m gpu -k "fast or Fast or FAST" --timeout 500 > gpurun_out/k_tests_default.log 2>&1; echo "rc=$?" >> gpurun_out/k_tests_default.log; tail -4 gpurun_out/k_tests_default.log
for v in m q h l; do
  echo "== FAST tree tests, REBOUND_B200_GW_VARIANT=$v"
  REBOUND_B200_GW_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_tree.py -q -m gpu -k "fast" --timeout 500 > gpurun_out/k_tests_$v.log 2>&1; echo "rc=$?" >> gpurun_out/k_tests_$v.log; tail -4 gpurun_out/k_tests_$v.log
done
echo "== A/B 2^20, 2^22"
timeout 900 python tools/gw_ab.py -hijklmnopq c4_20fast c4_22fast > gpurun_out/k_ab.jsonl 2> gpurun_out/k_ab.err
cat gpurun_out/k_ab.jsonl | cut -c1-220
echo "== 2^24"
timeout 600 python tools/gw_ab.py -hmq c4_24fast > gpurun_out/k_ab24.jsonl 2>> gpurun_out/k_ab.err
cat gpurun_out/k_ab24.jsonl | cut -c1-220
echo "== direct FAST"
timeout 300 python tools/measure.py c1fast d17fast c3s_18fast > gpurun_out/k_direct.jsonl 2>> gpurun_out/k_ab.err
cut -c1-300 gpurun_out/k_direct.jsonl
tail -3 gpurun_out/k_ab.err
