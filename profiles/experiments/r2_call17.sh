#!/bin/bash
# round 2, GPU call 17: programmatic dependent launch: correctness, then the 8-GPU shard size (625 images) on one GPU with PDL off / on
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_pytest_c17.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_c17.log
for pdl in 0 1; do
  RFN_PDL=$pdl python bench.py --images 625 --chunk 625 --steps 10 --warmup 3 --train-steps 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_625_pdl$pdl.json 2> gpurun_out/r2_bench_625_pdl$pdl.err; echo "625 pdl=$pdl rc=$?"
done
RFN_PDL=1 python bench.py --train-steps 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_5000_pdl1.json 2> gpurun_out/r2_bench_5000_pdl1.err; echo "5000 rc=$?"
RFN_PDL=0 python bench.py --train-steps 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_5000_pdl0.json 2> gpurun_out/r2_bench_5000_pdl0.err; echo "5000 pdl0 rc=$?"
python - <<'PY'
import json
for f in ('625_pdl0', '625_pdl1', '5000_pdl1', '5000_pdl0'):
    try:
        p = json.load(open(f'gpurun_out/r2_bench_{f}.json'))
        print(f, p['value'], p['ms_per_step'], p['config'].get('eager'), p['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
