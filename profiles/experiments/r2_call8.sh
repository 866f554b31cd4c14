#!/bin/bash
# round 2, GPU call 8: persistent cooperative decoder (guarded by timeouts: a grid-barrier bug would hang the kernel)
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_headline.py -m gpu -x -q -k "persistent or graphed" > gpurun_out/r2_pytest_pd.log 2>&1; echo "pd pytest rc=$?"
tail -30 gpurun_out/r2_pytest_pd.log
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_c8.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest_c8.log
timeout 900 python bench.py > gpurun_out/r2_bench_c8.json 2> gpurun_out/r2_bench_c8.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_c8.err
