#!/bin/bash
# round 2, GPU call 43: ncu --set full of the relay-free 2-CTA kernel on the stage-1 att_2_att_h launch (resnet encoder, 1024 images), bf16 and split fp16
set -x
mkdir -p gpurun_out
for M in 5 4; do
B="python bench.py --gemm-mode $M --images 1024 --chunk 1024 --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_h3_kernel -s 6 -c 1 -f -o gpurun_out/r2_h3_direct_m$M $B > gpurun_out/r2_ncu_h3_direct_m$M.log 2>&1; echo "ncu rc=$?"
done
ls -la gpurun_out/*.ncu-rep
