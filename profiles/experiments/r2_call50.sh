#!/bin/bash
# round 2, GPU call 50: single-pass top-3 in the vocabulary epilogue of the split-fp16 / bf16 GEMM; fast-exp A/B; sanitizers over the relay-free kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engines.py tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_pytest_c50.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_c50.log
for F in 0 1 0 1; do
RFN_H3_FAST_EXP=$F timeout 300 python bench.py --train-steps 0 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_c50_f${F}_$RANDOM.json 2> gpurun_out/r2_bench_c50.err; echo "bench rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_c50_f*.json')):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    b = d['bf16_mode']
    print(f, d['value'], d['ms_per_step'], d['kernel_time_shares']['gemm_logit'], round(d['kernel_time_shares']['gemm_logit']*d['ms_per_step'],2), d['seq_checksum'], '| bf16', b['value'], b['kernel_time_shares']['gemm_logit'], b['roofline_attention']['frac'], b['roofline_gemm']['frac'])
PY
for tool in memcheck racecheck synccheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 5 python profiles/experiments/r2_sanitizer_cases.py > gpurun_out/r2_sanitizer2_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER_CASES_OK|Error|hazard" gpurun_out/r2_sanitizer2_$tool.log | head -6
done
