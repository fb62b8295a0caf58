#!/bin/bash
export TB_FF_SPIN_LIMIT_MS=1500
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for rep in 1 2; do
  echo -n "r1: "; TB_LIB_PATH=$PWD/build/variants/lib_r1.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1
  for st in 0 8 12; do
  echo -n "base sm_stride $st: "; TB_FF_SERVER_SM_STRIDE=$st TB_LIB_PATH=$PWD/build/variants/lib_base.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1
done; done | tee gpurun_out/r2f_ab.log
TB_FF_SERVER_SM_STRIDE=0 TB_FF_DIAG_DUMP=1 TB_LIB_PATH=$PWD/build/variants/lib_diag.so timeout 300 python tools/time_kernels.py f64 1048576 > gpurun_out/r2f_diag.log 2>&1
grep -v "late landing" gpurun_out/r2f_diag.log | tail -16
