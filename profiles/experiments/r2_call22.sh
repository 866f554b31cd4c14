#!/bin/bash
set -x
mkdir -p gpurun_out
python profiles/experiments/r2_xe_graph_once.py 1 5 2>&1 | tail -2
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_xe_graph_launches.csv python profiles/experiments/r2_xe_graph_once.py 1 1 > gpurun_out/r2_xe_graph_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_xe_graph_ncu.log
wc -l gpurun_out/r2_xe_graph_launches.csv
