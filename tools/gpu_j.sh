#!/bin/bash
export TB_FF_SPIN_LIMIT_MS=1500
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
  for v in r1 base stride3 stride4; do echo -n "$v: "; TB_FF_SERVER_SM_STRIDE=0 TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1; done
  for st in 8 12 16; do echo -n "base sm_stride $st: "; TB_FF_SERVER_SM_STRIDE=$st TB_LIB_PATH=$PWD/build/variants/lib_base.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1; done
done | tee gpurun_out/r2j_ab.log
TB_FF_SERVER_SM_STRIDE=0 TB_FF_DIAG_DUMP=1 TB_LIB_PATH=$PWD/build/variants/lib_diag.so timeout 300 python tools/time_kernels.py f64 1048576 > gpurun_out/r2j_diag.log 2>&1
grep -v "late landing" gpurun_out/r2j_diag.log | tail -16
echo "--- f32"; for v in r1 base; do echo -n "$v: "; TB_FF_SERVER_SM_STRIDE=0 TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f32 1048576 3 2>&1 | tail -1; done
echo "--- 16384 envs"; for v in r1 base; do echo -n "$v: "; TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f64 16384 5 2>&1 | tail -1; done
