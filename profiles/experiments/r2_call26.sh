#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_overlap_sync.py tests/test_gpu_tape.py -x -q 2>&1 | tail -15
