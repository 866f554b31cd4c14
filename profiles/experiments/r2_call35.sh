#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py tests/test_gpu_tape.py -x -q -k "rl" 2>&1 | tail -4
python profiles/experiments/r2_phase_timeline.py rl 1 2>&1 | tail -27
