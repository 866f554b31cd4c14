#!/bin/bash
# round 2, GPU call 58: the eval-driver tests against the reference-written fixtures, smoke() on the final library
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_eval_utils.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
