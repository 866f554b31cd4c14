#!/bin/bash
# two GPUs: data-parallel gradient parity incl. the overlapped all-reduce with the fused tape; XE / RL lines at N=2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddp_nccl.py -x -q -s > gpurun_out/r2_ddp_nccl_2gpu_b.log 2>&1; echo "ddp rc=$?"
tail -8 gpurun_out/r2_ddp_nccl_2gpu_b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n2_b.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n2_b.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'])
for k in ('xe_train', 'rl_train'):
    x = d[k]
    print(k, x['ms_per_step'], x['deduplicated'], x['eager']['ms_per_step'])
PY
