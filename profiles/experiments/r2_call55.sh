#!/bin/bash
# round 2, GPU call 55: ncu --set full of the gates GEMM, the logits GEMM (vocabulary epilogue) and the bf16-feature attention step on the final build
set -x
mkdir -p gpurun_out
B4="python bench.py --gemm-mode 4 --images 1024 --chunk 1024 --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline --no-bf16 --graph 0"
B5="python bench.py --gemm-mode 5 --images 1024 --chunk 1024 --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline --graph 0"
N="ncu --set full --clock-control none --import-source on --kernel-name-base mangled"
timeout 300 $N -k regex:gemm_h3_kernelILi0ELi3ELi2 --launch-skip 6 --launch-count 1 -f -o gpurun_out/r2_final_gates $B4 > gpurun_out/r2_ncu_final_gates.log 2>&1; echo "gates rc=$?"
timeout 300 $N -k regex:gemm_h3_kernelILi2ELi3ELi2 --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_final_logits $B4 > gpurun_out/r2_ncu_final_logits.log 2>&1; echo "logits rc=$?"
timeout 300 $N -k regex:attention_step_kernelILb1ELi2 --launch-skip 0 --launch-count 1 -f -o gpurun_out/r2_final_attn_bf16 $B5 > gpurun_out/r2_ncu_final_attn_bf16.log 2>&1; echo "attn rc=$?"
ls -la gpurun_out/r2_final_*.ncu-rep
