#!/bin/bash
# round 2, GPU call 15: e2e chunk-size sweep (PCIe-bound leg) + quick regression of the changed launch config
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_engines.py -m gpu -x -q > gpurun_out/r2_pytest_c15.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_c15.log
for c in 125 250 500; do
  python bench.py --train-steps 0 --no-cpu-baseline --e2e-chunk $c --steps 3 --warmup 3 > gpurun_out/r2_bench_e2e_$c.json 2> gpurun_out/r2_bench_e2e_$c.err; echo "chunk $c rc=$?"
  python -c "
import json; p=json.load(open('gpurun_out/r2_bench_e2e_$c.json')); print($c, p['value'], p['ms_per_step'], p['e2e']['value'], p['e2e']['ms_per_step'], p['e2e']['h2d_gbs_per_gpu'], p['e2e']['h2d_copy_peak_gbs'])"
done
