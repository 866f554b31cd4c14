#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_training.py -x -q -k "config2_size and as_written or adam or graphed_xe" > gpurun_out/r2_sanitizer_memcheck_train_full.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/r2_sanitizer_memcheck_train_full.log
