#!/bin/bash
# round 2, GPU call 3: engine mode 4 (split fp16) as the default through the whole path
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_mode4.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_pytest_mode4.log
python bench.py > gpurun_out/r2_bench_mode4.json 2> gpurun_out/r2_bench_mode4.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_mode4.err
B="python bench.py --images 1024 --chunk 1024 --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm_h3_kernel<1" -s 0 -c 1 -f -o gpurun_out/r2_h3_score $B > gpurun_out/r2_ncu_h3a.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_mode4.csv $B > gpurun_out/r2_ncu_h3d.log 2>&1
