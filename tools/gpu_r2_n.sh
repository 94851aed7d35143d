#!/bin/bash
# Round 2, GPU call N (1 GPU): full-size parity tests (C1..C4 at BASELINE.json's sizes) with the shipped FAST kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu --timeout 800 --durations=8 > gpurun_out/n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/n_tests.log; tail -16 gpurun_out/n_tests.log
