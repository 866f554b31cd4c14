#!/bin/bash
# round 2, GPU call 52: ncu launch list (gpu__time_duration.sum, natural clocks) of the default bench command's decode step on the final build
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_final_5000img.csv python bench.py --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline --no-bf16 --graph 0 > gpurun_out/r2_launches_final_5000img.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2_launches_final_5000img.log | cut -c1-600
wc -l gpurun_out/r2_launches_final_5000img.csv
