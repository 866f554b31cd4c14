#!/bin/bash
set -x
mkdir -p gpurun_out
python profiles/experiments/r2_phase_timeline.py xe 1 > gpurun_out/r2_phase_xe.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_phase_xe.log | tail -40
python profiles/experiments/r2_phase_timeline.py rl 1 > gpurun_out/r2_phase_rl.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_phase_rl.log | tail -40
