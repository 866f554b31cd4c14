#!/bin/bash
# round 2, GPU call 54: does the nvidia-smi polling during the timed region slow the step?  (main region vs the unsampled second measurement)
set -x
mkdir -p gpurun_out
for i in 1 2; do for S in 1 0; do
RFN_BENCH_SAMPLER=$S timeout 300 python bench.py --graph 2 --train-steps 0 --no-e2e --no-cpu-baseline --no-bf16 > gpurun_out/r2_bench_c54_s${S}_$i.json 2> gpurun_out/r2_bench_c54.err; echo "bench rc=$?"
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_c54_s*.json')):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'main(graph)', d['ms_per_step'], 'other', d['config']['other_issue_mode']['mode'], d['config']['other_issue_mode']['ms_per_step'], d['clocks'])
PY
