#!/bin/bash
# round 2, GPU call 45: ncu --set full of the bf16-feature attention step (mode 5) and of the logits GEMM with the vocabulary epilogue (mode 4)
set -x
mkdir -p gpurun_out
B="python bench.py --gemm-mode 5 --images 1024 --chunk 1024 --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_step_kernel -s 8 -c 1 -f -o gpurun_out/r2_attn_bf16 $B > gpurun_out/r2_ncu_attn_bf16.log 2>&1; echo "ncu rc=$?"
B="python bench.py --gemm-mode 4 --images 1024 --chunk 1024 --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline --no-bf16"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_h3_kernel.2..3 -s 4 -c 1 -f -o gpurun_out/r2_h3_logits $B > gpurun_out/r2_ncu_h3_logits.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_ncu_h3_logits.log
ls -la gpurun_out/*.ncu-rep
