#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py tests/test_gpu_tape.py tests/test_gpu_overlap_sync.py -x -q 2>&1 | tail -6
python profiles/experiments/r2_phase_timeline.py xe 1 2>&1 | tail -23
python profiles/experiments/r2_rl_graph_once.py 1 5 2>&1 | tail -1
