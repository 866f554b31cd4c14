#!/bin/bash
# compute-sanitizer memcheck over the new training kernels (fused tape, strided criteria, multi-source cell backward, vector
# red / Adam, maxout cells) on the tiny configurations
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tape.py -x -q -k "xe or rl or maxout" > gpurun_out/r2_sanitizer_memcheck_tape.log 2>&1; echo "memcheck rc=$?"
tail -12 gpurun_out/r2_sanitizer_memcheck_tape.log
