#!/bin/bash
# round 2, GPU call 48: A/B of the bf16 context sum (8-byte vs 16-byte loads) inside the bench on ONE box, alternating
set -x
mkdir -p gpurun_out
for i in 1 2; do for W in 0 1; do
RFN_ATT_BF16_WIDE=$W timeout 300 python bench.py --gemm-mode 5 --train-steps 0 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_c48_w${W}_$i.json 2> gpurun_out/r2_bench_c48.err; echo "bench rc=$?"
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_c48_w*.json')):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['roofline_gemm']['achieved'], d['roofline_attention']['achieved'], d['roofline_attention']['avg_launch_ms'], d['clocks']['sm_mhz'])
PY
