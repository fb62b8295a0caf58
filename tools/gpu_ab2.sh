#!/bin/bash
export TB_FF_SPIN_LIMIT_MS=1500
for rep in 1 2 3; do for v in $1; do
  echo -n "$v: "; TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py ${2:-f64} 1048576 3 2>&1 | tail -1
done; done
