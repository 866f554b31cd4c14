#!/bin/bash
# round 2, GPU call 44: bench line with the bf16 sub-line (bf16 context sum unrolled by 8), and the 8-GPU shard size (625 images) on one GPU
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --train-steps 0 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_c44.json 2> gpurun_out/r2_bench_c44.err; echo "bench rc=$?"
timeout 300 python bench.py --images 625 --chunk 625 --train-steps 0 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_c44_625.json 2> gpurun_out/r2_bench_c44_625.err; echo "bench 625 rc=$?"
tail -3 gpurun_out/r2_bench_c44.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_c44.json', 'gpurun_out/r2_bench_c44_625.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline_attention']['achieved'], d['clocks'])
        print('  shares', d['kernel_time_shares'])
        b = d.get('bf16_mode')
        if b: print('  bf16', b['value'], b['ms_per_step'], b['roofline_gemm']['achieved'], b['roofline_gemm']['frac'], b['roofline_attention']['achieved'], b['roofline_attention']['frac'], b['captions_equal_to_fp32_mode'])
    except Exception as e:
        print(f, 'ERR', e)
PY
