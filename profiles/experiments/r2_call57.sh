#!/bin/bash
# round 2, GPU call 57: full GPU suite on the final commit
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_c57.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_c57.log
