#!/bin/bash
# round 2, GPU call 6: graphs, RL step, bf16 attention; full test-suite; full bench with graph + config-1 + RL graph lines
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_c6.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2_pytest_c6.log
python bench.py > gpurun_out/r2_bench_c6.json 2> gpurun_out/r2_bench_c6.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_c6.err
python bench.py --gemm-mode 5 --train-steps 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_bf16_c6.json 2> gpurun_out/r2_bench_bf16_c6.err; echo "bench bf16 rc=$?"
