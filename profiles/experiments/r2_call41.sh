#!/bin/bash
# ncu --set full of the single-pass bf16 kernel (engine mode 5) on the stage-1 att_2_att_h launch: what bounds it?
set -x
mkdir -p gpurun_out
B="python bench.py --gemm-mode 5 --images 1024 --chunk 1024 --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_h3_kernel -s 6 -c 1 -f -o gpurun_out/r2_h3_bf16_score $B > gpurun_out/r2_ncu_h3_bf16.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_ncu_h3_bf16.log
ls -la gpurun_out/r2_h3_bf16_score.ncu-rep
