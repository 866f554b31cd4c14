#!/bin/bash
# round 2, GPU call 56: the build after rfn_set_att_bf16_variant replaced the environment read: smoke, bf16 / engine tests, variants give the same captions
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python -m pytest tests/test_gpu_headline.py tests/test_gpu_engines.py -m gpu -x -q > gpurun_out/r2_pytest_c56.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r2_pytest_c56.log
for W in 0 1 2; do
RFN_ATT_BF16_WIDE=$W timeout 200 python bench.py --gemm-mode 5 --images 1250 --chunk 1250 --train-steps 0 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $W', d['seq_checksum'], d['ms_per_step'], d['roofline_attention']['achieved'])"
done
