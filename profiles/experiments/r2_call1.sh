#!/bin/bash
# round 2, GPU call 1: baseline of the shipped build -- gpu tests, ncu --set full of the three tensor kernels, launch list, bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest0.log 2>&1; echo "pytest rc=$?"
B="python bench.py --gemm-mode 3 --images 1024 --chunk 1024 --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2p_kernel -s 0 -c 1 -f -o gpurun_out/r2_tc2p_score $B > gpurun_out/r2_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 0 -c 1 -f -o gpurun_out/r2_tc2_gates $B > gpurun_out/r2_ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc2p_kernel<2>" -s 0 -c 1 -f -o gpurun_out/r2_tc2p_vocab $B > gpurun_out/r2_ncu_c.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches0.csv $B > gpurun_out/r2_ncu_d.log 2>&1
python bench.py > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_pytest0.log
