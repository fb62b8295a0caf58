#!/bin/bash
# fast-forward diagnostics of instrumented builds: tools/gpu_diag.sh "diag diagnl" [ring]
export TB_FF_SPIN_LIMIT_MS=1500 TB_FF_DIAG_DUMP=1 TB_RING=${2:-32}
for v in ${1:-diag}; do
  echo "=== $v"
  TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_kernels.py f64 1048576 > gpurun_out/diag_$v.log 2>&1
  grep -v "late landing" gpurun_out/diag_$v.log | tail -22
done
