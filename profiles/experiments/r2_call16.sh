#!/bin/bash
# round 2, GPU call 16 (8 GPUs): full suite on one GPU, smoke, then the scaling points N = 1, 8 on the same box
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_c16.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_c16.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_c16.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke_c16.log
timeout 900 python bench.py --train-steps 0 --no-cpu-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 --train-steps 0 --no-cpu-baseline > gpurun_out/r2_scale_n8.json 2> gpurun_out/r2_scale_n8.err; echo "n8 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 5 --warmup 3 --train-steps 0 --no-cpu-baseline > gpurun_out/r2_scale_n4.json 2> gpurun_out/r2_scale_n4.err; echo "n4 rc=$?"
python - <<'PY'
import json
for n in (1, 4, 8):
    try:
        p = json.load(open(f'gpurun_out/r2_scale_n{n}.json'))
        print(n, p['value'], p['ms_per_step'], p['e2e']['value'], p['config'].get('eager'), p['clocks'])
    except Exception as e:
        print(n, 'ERR', e)
PY
