#!/bin/bash
# round 2, GPU call 4: mode 4 with the weight cache: full gpu tests, smoke, bench, near-tie audit
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_c4.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_pytest_c4.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_c4.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke_c4.log
python bench.py > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_c4.err
timeout 900 python profiles/experiments/r2_neartie_audit.py 5000 32 > gpurun_out/r2_neartie_audit.log 2>&1; echo "audit rc=$?"
tail -20 gpurun_out/r2_neartie_audit.log
