#!/bin/bash
# full ncu capture of the fast-forward launch (26th step) for given variants: tools/gpu_ncu_ff.sh "base r1" tag
export TB_FF_SPIN_LIMIT_MS=20000
for v in $1; do
  TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:ff_kernel -s 25 -c 1 -f -o gpurun_out/prof_${2:-r2}_${v}_ff \
     python tools/prof_swing.py f64 1048576 > gpurun_out/ncu_${2:-r2}_${v}.log 2>&1
  tail -n 2 gpurun_out/ncu_${2:-r2}_${v}.log
done
