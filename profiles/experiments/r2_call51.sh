#!/bin/bash
# round 2, GPU call 51 (2 GPUs): the bench line as the driver launches it for N = 2 (issue-mode trial, bf16 sub-line, training lines under NCCL)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_c51_n2.json 2> gpurun_out/r2_bench_c51_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2_bench_c51_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_c51_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config']['cuda_graph'], d['config']['other_issue_mode'], d['e2e'] and d['e2e']['value'], d['xe_train'] and d['xe_train']['ms_per_step'], d['rl_train'] and d['rl_train']['ms_per_step'], d['bf16_mode'] and (d['bf16_mode']['value'], d['bf16_mode']['captions_equal_to_fp32_mode']))
PY
