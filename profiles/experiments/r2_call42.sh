#!/bin/bash
# round 2, GPU call 42: the 2-CTA kernels without the per-k-block relay (both CTAs' TMA bytes counted on the leader's barrier)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engines.py -m gpu -x -q > gpurun_out/r2_pytest_c42.log 2>&1; echo "engines pytest rc=$?"
tail -5 gpurun_out/r2_pytest_c42.log
timeout 200 python profiles/experiments/r2_h3_perf2.py direct 2>&1 | tail -8
timeout 300 python bench.py --train-steps 0 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_c42_m4.json 2> gpurun_out/r2_bench_c42_m4.err; echo "bench m4 rc=$?"
timeout 300 python bench.py --gemm-mode 5 --train-steps 0 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_c42_m5.json 2> gpurun_out/r2_bench_c42_m5.err; echo "bench m5 rc=$?"
python - <<'PY'
import json
for m in (4, 5):
    try:
        d = json.loads(open(f'gpurun_out/r2_bench_c42_m{m}.json').read().strip().splitlines()[-1])
        print(m, d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_attention']['achieved'], d['clocks'], d['kernel_time_shares'])
    except Exception as e:
        print(m, 'ERR', e)
PY
