#!/bin/bash
# round 2, GPU call 2: first run of the split-fp16 engine: unit tests, headline gating tests, perf of the GEMM alone
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engines.py -x -q -k "split" > gpurun_out/r2_pytest_h3.log 2>&1; echo "h3 pytest rc=$?"
tail -5 gpurun_out/r2_pytest_h3.log
timeout 600 python profiles/experiments/r2_h3_perf.py > gpurun_out/r2_h3_perf.log 2>&1; echo "perf rc=$?"
cat gpurun_out/r2_h3_perf.log | tail -8
timeout 1200 python -m pytest tests/test_gpu_headline.py -x -q > gpurun_out/r2_pytest_headline.log 2>&1; echo "headline rc=$?"
tail -15 gpurun_out/r2_pytest_headline.log
