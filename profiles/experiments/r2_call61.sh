#!/bin/bash
# round 2, GPU call 61: the default bench line on the final commit
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench_c61.json 2> gpurun_out/r2_bench_c61.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_c61.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config']['cuda_graph'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline_attention']['frac'], 'cpu', d['cpu_baseline'], 'bf16', d['bf16_mode']['value'], 'xe', d['xe_train']['ms_per_step'], 'rl', d['rl_train']['ms_per_step'])
PY
