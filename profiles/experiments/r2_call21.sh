#!/bin/bash
set -x
mkdir -p gpurun_out
python profiles/experiments/r2_train_bench.py both 5 1 > gpurun_out/r2_train_fused.json 2> gpurun_out/r2_train_fused.err; echo "rc=$?"
tail -5 gpurun_out/r2_train_fused.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_train_fused.json'))
for k in ('xe', 'rl'):
    x = d[k]
    print(k, 'graph', x['ms_per_step'], 'dedup', x['deduplicated']['ms_per_step'], 'eager', x['eager']['ms_per_step'], x.get('kernel_ms_and_launches'))
PY
