#!/bin/bash
export TB_FF_SPIN_LIMIT_MS=3000
for rep in 1 2; do
 for v in base minb5 minb6; do for st in 12 16; do
  echo -n "$v sm_stride $st: "; TB_FF_SERVER_SM_STRIDE=$st TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1
 done; done
done
