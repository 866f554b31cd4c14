#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_c24.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_c24.log
python profiles/experiments/r2_xe_graph_once.py 1 5 2>&1 | tail -1
python profiles/experiments/r2_rl_graph_once.py 1 5 2>&1 | tail -1
python profiles/experiments/r2_rl_graph_once.py 5 5 2>&1 | tail -1
