#!/bin/bash
# server-SM stride sweep (runtime knob), f64 1 Mi envs
export TB_FF_SPIN_LIMIT_MS=1500
for rep in 1 2; do for st in 0 4 6 8 12 16; do
  echo -n "sm_stride $st: "; TB_FF_SERVER_SM_STRIDE=$st TB_LIB_PATH=$PWD/build/variants/lib_base.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1
done; done | tee gpurun_out/r2d_sm.log
TB_FF_SERVER_SM_STRIDE=8 TB_FF_DIAG_DUMP=1 TB_LIB_PATH=$PWD/build/variants/lib_diag.so timeout 300 python tools/time_kernels.py f64 1048576 > gpurun_out/r2d_diag.log 2>&1
grep -v "late landing" gpurun_out/r2d_diag.log | tail -16
TB_FF_SERVER_SM_STRIDE=8 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
