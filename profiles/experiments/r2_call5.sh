#!/bin/bash
# round 2, GPU call 5: new tests (eval oracle, ensemble greedy, beam 16, bf16 mode), ncu of the split-fp16 kernels, bf16 bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_eval_utils.py tests/test_gpu_headline.py tests/test_ingest.py "tests/test_gpu_parity.py::test_other_beam_widths" "tests/test_gpu_parity.py::test_max_beam_width_and_bad_arguments" -m gpu -x -q > gpurun_out/r2_pytest_c5.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_pytest_c5.log
B="python bench.py --images 1024 --chunk 1024 --steps 1 --warmup 1 --train-steps 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_h3_kernel -s 6 -c 1 -f -o gpurun_out/r2_h3_score $B > gpurun_out/r2_ncu_h3a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_h3_kernel -s 7 -c 1 -f -o gpurun_out/r2_h3_gates $B > gpurun_out/r2_ncu_h3b.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:attention_step_kernel -s 0 -c 1 -f -o gpurun_out/r2_attn_step $B > gpurun_out/r2_ncu_h3c.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:split_rows_kernel -s 0 -c 1 -f -o gpurun_out/r2_split $B > gpurun_out/r2_ncu_h3e.log 2>&1
python bench.py --gemm-mode 5 --train-steps 0 --no-cpu-baseline > gpurun_out/r2_bench_bf16.json 2> gpurun_out/r2_bench_bf16.err; echo "bench bf16 rc=$?"
tail -2 gpurun_out/r2_bench_bf16.err
