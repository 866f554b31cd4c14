#!/bin/bash
# eight GPUs, one box: the driver's scaling sequence N = 1, 8 back to back (inference line + the training lines with the fused tape)
set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_scale_b_n1.json 2> gpurun_out/r2_scale_b_n1.err; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale_b_n8.json 2> gpurun_out/r2_scale_b_n8.err; echo "n8 rc=$?"
tail -3 gpurun_out/r2_scale_b_n8.err
python - <<'PY'
import json
for f in ('n1', 'n8'):
    try:
        d = json.loads(open(f'gpurun_out/r2_scale_b_{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', d['value'], d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), 'clk', d['clocks']['sm_mhz'])
        for k in ('xe_train', 'rl_train'):
            x = d[k]
            print('  ', k, x['value'], x['ms_per_step'], 'dedup', x['deduplicated']['ms_per_step'], 'eager', x['eager']['ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
