#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py --images 625 --chunk 625 --steps 10 --warmup 3 --train-steps 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_625_half.json 2> gpurun_out/r2_bench_625_half.err; echo "625 rc=$?"
python bench.py --images 1250 --chunk 1250 --steps 6 --warmup 3 --train-steps 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_1250_half.json 2> gpurun_out/r2_bench_1250_half.err; echo "1250 rc=$?"
python bench.py --train-steps 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_5000_half.json 2> gpurun_out/r2_bench_5000_half.err; echo "5000 rc=$?"
python - <<'PY'
import json
for f in ('625_half', '1250_half', '5000_half'):
    try:
        p = json.load(open(f'gpurun_out/r2_bench_{f}.json'))
        print(f, p['value'], p['ms_per_step'], p['config'].get('eager'), p['clocks']['sm_mhz'], p['roofline_gemm']['avg_launch_ms'])
    except Exception as e:
        print(f, 'ERR', e)
PY
