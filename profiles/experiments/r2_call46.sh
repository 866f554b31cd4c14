#!/bin/bash
# round 2, GPU call 46: bf16 context sum with 16-byte loads + pointer-increment addressing (own kernel instantiation); bf16-mode tests
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_headline.py tests/test_gpu_engines.py -m gpu -x -q -k "bf16 or mode5 or attention" > gpurun_out/r2_pytest_c46.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_c46.log
timeout 300 python bench.py --train-steps 0 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_c46.json 2> gpurun_out/r2_bench_c46.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_c46.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_c46.json',):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline_attention']['achieved'], d['clocks'])
        b = d.get('bf16_mode')
        if b: print('  bf16', b['value'], b['ms_per_step'], b['roofline_gemm']['achieved'], b['roofline_gemm']['frac'], b['roofline_attention']['achieved'], b['roofline_attention']['frac'], b['captions_equal_to_fp32_mode'], b['kernel_time_shares'])
    except Exception as e:
        print(f, 'ERR', e)
PY
