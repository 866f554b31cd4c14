#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_engines.py tests/test_gpu_training.py tests/test_gpu_tape.py -x -q 2>&1 | tail -4
python profiles/experiments/r2_xe_graph_once.py 1 5 2>&1 | tail -1
python profiles/experiments/r2_xe_graph_once.py 5 5 2>&1 | tail -1
python profiles/experiments/r2_rl_graph_once.py 1 5 2>&1 | tail -1
