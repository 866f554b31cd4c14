#!/bin/bash
# round 2, GPU call 14 (8 GPUs): the scaling point the driver measures at round end
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --train-steps 2 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?"
tail -3 gpurun_out/r2_bench_n8.err
