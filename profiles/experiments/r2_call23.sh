#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tape.py tests/test_gpu_training.py -x -q > gpurun_out/r2_pytest_tape.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_tape.log
python profiles/experiments/r2_xe_graph_once.py 1 5 2>&1 | tail -1
python profiles/experiments/r2_xe_graph_once.py 5 5 2>&1 | tail -1
python profiles/experiments/r2_rl_graph_once.py 1 5 2>&1 | tail -1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_rl_graph_launches.csv python profiles/experiments/r2_rl_graph_once.py 1 1 > gpurun_out/r2_rl_graph_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2_rl_graph_ncu.log
