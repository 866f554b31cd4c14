#!/bin/bash
# round 2, GPU call 7 (2 GPUs): 2-rank NCCL gradient parity test, launch-overhead probe, N=2 bench
set -x
mkdir -p gpurun_out
nvidia-smi -L
python profiles/experiments/r2_launch_overhead.py > gpurun_out/r2_launch_overhead.log 2>&1; cat gpurun_out/r2_launch_overhead.log
timeout 900 python -m pytest tests/test_gpu_ddp_nccl.py -m gpu -x -q -s > gpurun_out/r2_ddp_nccl_2gpu.log 2>&1; echo "ddp rc=$?"
tail -8 gpurun_out/r2_ddp_nccl_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --train-steps 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2_bench_n2.err
