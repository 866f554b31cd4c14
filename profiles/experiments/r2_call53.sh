#!/bin/bash
# round 2, GPU call 53: final validation of the build -- full GPU suite, smoke(), the default bench line
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_c53.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_c53.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_c53.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/r2_smoke_c53.log
timeout 900 python bench.py > gpurun_out/r2_bench_c53.json 2> gpurun_out/r2_bench_c53.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_c53.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_c53.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config']['cuda_graph'], d['config']['other_issue_mode'])
print('e2e', d['e2e'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'attn', d['roofline_attention']['achieved'], d['roofline_attention']['frac'])
print('shares', d['kernel_time_shares'])
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
print('cpu', d['cpu_baseline'])
b = d['bf16_mode']; print('bf16', b['value'], b['roofline_gemm']['frac'], b['roofline_attention']['frac'], b['captions_equal_to_fp32_mode'])
print('xe', d['xe_train']['ms_per_step'], d['xe_train']['value'], d['xe_train']['deduplicated'])
print('rl', d['rl_train']['ms_per_step'], d['rl_train']['value'])
print('ens', d['ensemble']); print('cfg1', d['config1_latency']); print('ciderd', d['ciderd_reward'])
PY
