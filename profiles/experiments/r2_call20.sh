#!/bin/bash
# fused training tape: parity tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tape.py tests/test_gpu_training.py -x -q > gpurun_out/r2_pytest_tape.log 2>&1; echo "rc=$?"
tail -60 gpurun_out/r2_pytest_tape.log
