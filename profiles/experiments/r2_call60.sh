#!/bin/bash
# round 2, GPU call 60: the device reward path against the fixture written by the reference's compute_reward
timeout 200 python -m pytest tests/test_ciderd.py -m gpu -x -q 2>&1 | tail -3
