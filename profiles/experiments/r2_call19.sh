#!/bin/bash
# where the at:: glue kernels of the XE step come from (torch.profiler), as written and de-duplicated
set -x
mkdir -p gpurun_out
python profiles/experiments/r2_train_glue.py 1 > gpurun_out/r2_train_glue_1.log 2>&1; echo "glue1 rc=$?"
tail -120 gpurun_out/r2_train_glue_1.log
