#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_c30.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2_pytest_c30.log
