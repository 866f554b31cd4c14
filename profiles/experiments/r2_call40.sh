#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python profiles/experiments/r2_h3_cluster4.py > gpurun_out/r2_h3_cluster4.log 2>&1; echo "rc=$?"
tail -40 gpurun_out/r2_h3_cluster4.log
