#!/bin/bash
# GPU box: the whole -m gpu suite, then the per-episode timing with per-CTA server roles and with server SMs
export TB_FF_SPIN_LIMIT_MS=3000
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for rep in 1 2; do for st in 0 12; do echo -n "sm_stride $st: "; TB_FF_SERVER_SM_STRIDE=$st timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1; done; done
