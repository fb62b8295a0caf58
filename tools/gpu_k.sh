#!/bin/bash
export TB_FF_SPIN_LIMIT_MS=3000
timeout 1200 python -m pytest tests/test_rollout_gpu.py -m gpu -x -q 2>&1 | tail -5
python tools/time_policy_rollout.py 16384 31 2>&1 | tail -7
