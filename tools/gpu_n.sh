#!/bin/bash
export TB_FF_SPIN_LIMIT_MS=3000
timeout 2400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "racket_court or racket_rests" -s 2>&1 | tail -12
for rep in 1 2; do for v in prev base; do echo -n "$v: "; TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1; done; done
