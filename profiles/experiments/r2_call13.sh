#!/bin/bash
# round 2, GPU call 13: full suite with the persistent decoder, sanitizers over the round-2 kernels, full bench
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_c13.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest_c13.log
python profiles/experiments/r2_sanitizer_cases.py > gpurun_out/r2_sanitizer_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/r2_sanitizer_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python profiles/experiments/r2_sanitizer_cases.py > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER_CASES_OK|Error|hazard" gpurun_out/r2_sanitizer_$tool.log | head -8
done
timeout 900 python bench.py > gpurun_out/r2_bench_c13.json 2> gpurun_out/r2_bench_c13.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_c13.err
