#!/bin/bash
# the driver's round-end sequence on one GPU: smoke, the default bench line, the reference arm
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_c29.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke_c29.log
( time python bench.py ) > gpurun_out/r2_bench_n1_c29.json 2> gpurun_out/r2_bench_n1_c29.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1_c29.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1_c29.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d['clocks'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'attn', d['roofline_attention']['achieved'])
print('shares', d['kernel_time_shares'])
for k in ('xe_train', 'rl_train'):
    x = d[k]
    print(k, x['ms_per_step'], x['deduplicated'], x['eager']['ms_per_step'])
print('ens', d['ensemble'])
print('cfg1', d['config1_latency'])
PY
