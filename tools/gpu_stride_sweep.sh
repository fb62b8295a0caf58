#!/bin/bash
# server SM stride sweep of one build: tools/gpu_stride_sweep.sh <variant> "<strides>"
export TB_FF_SPIN_LIMIT_MS=1500
for rep in 1 2; do for s in $2; do
  echo -n "$1 stride $s: "; TB_FF_SERVER_SM_STRIDE=$s TB_LIB_PATH=$PWD/build/variants/lib_$1.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1
done; done
