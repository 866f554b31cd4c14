#!/bin/bash
set -x
mkdir -p gpurun_out
for w in 1 0; do
RFN_ATT16_WIDE=$w python bench.py --gemm-mode 5 --images 2048 --chunk 2048 --train-steps 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_bf16_att_w$w.json 2> gpurun_out/r2_bf16_att_w$w.err; echo "rc=$?"
done
python - <<'PY'
import json
for w in (1, 0):
    d = json.loads(open(f'gpurun_out/r2_bf16_att_w{w}.json').read().strip().splitlines()[-1])
    print(w, 'value', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], 'attn', d['roofline_attention']['achieved'], d['roofline_attention']['avg_launch_ms'], 'gemm', d['roofline_gemm']['achieved'])
PY
