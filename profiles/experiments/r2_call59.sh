#!/bin/bash
# round 2, GPU call 59: the review-cell mirror against the fixture written by the reference class
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "review_net_core" 2>&1 | tail -3
