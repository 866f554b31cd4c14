#!/bin/bash
# round 2, GPU call 47: per-launch durations of the attention step (ncu launch list) for fp32 features, bf16 8-byte loads, bf16 16-byte loads
set -x
mkdir -p gpurun_out
B="python bench.py --images 5000 --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline --no-bf16 --graph 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attention_step_kernel -c 400 --csv --log-file gpurun_out/r2_attn_launches_fp32.csv $B --gemm-mode 4 > /dev/null 2>&1; echo "rc=$?"
RFN_ATT_BF16_WIDE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attention_step_kernel -c 400 --csv --log-file gpurun_out/r2_attn_launches_bf16_8B.csv $B --gemm-mode 5 > /dev/null 2>&1; echo "rc=$?"
RFN_ATT_BF16_WIDE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attention_step_kernel -c 400 --csv --log-file gpurun_out/r2_attn_launches_bf16_16B.csv $B --gemm-mode 5 > /dev/null 2>&1; echo "rc=$?"
wc -l gpurun_out/r2_attn_launches_*.csv
