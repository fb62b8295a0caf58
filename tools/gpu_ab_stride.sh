#!/bin/bash
# A/B of builds at given server SM strides (a flight-side change shows where the flights bound the launch: stride 8-9; a
# server-side one where the servers do: stride 14-16):  tools/gpu_ab_stride.sh "<variants>" "<strides>"
export TB_FF_SPIN_LIMIT_MS=1500
for s in $2; do for rep in 1 2; do for v in $1; do
  echo -n "stride $s $v: "; TB_FF_SERVER_SM_STRIDE=$s TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1 | cut -c1-75
done; done; done
