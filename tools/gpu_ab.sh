#!/bin/bash
# parity suite + A/B timing of variants + diag.   usage: tools/gpu_ab.sh "variants" [tag] [pytest: yes|no]
export TB_FF_SPIN_LIMIT_MS=1500
TAG=${2:-ab}
mkdir -p gpurun_out
if [ "${3:-yes}" = yes ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log; fi
for rep in 1 2; do for v in $1; do
  echo -n "$v: "; TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1
done; done | tee gpurun_out/${TAG}_ab.log
if [ -f build/variants/lib_diag.so ]; then
TB_FF_DIAG_DUMP=1 TB_LIB_PATH=$PWD/build/variants/lib_diag.so timeout 300 python tools/time_kernels.py f64 1048576 > gpurun_out/${TAG}_diag.log 2>&1
grep -v "late landing" gpurun_out/${TAG}_diag.log | tail -19
grep "late landing" gpurun_out/${TAG}_diag.log | sort -t- -k2 -n | head -8
fi
