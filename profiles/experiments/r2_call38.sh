#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py --gemm-mode 5 --train-steps 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_n1_bf16_b.json 2> gpurun_out/r2_bench_n1_bf16_b.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1_bf16_b.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], d['clocks'])
print('gemm', d['roofline_gemm']['achieved'], d['roofline_gemm']['frac'], 'attn', d['roofline_attention']['achieved'], d['roofline_attention']['frac'])
print(d['kernel_time_shares'])
PY
