#!/bin/bash
# round 2, GPU call A: parity suite on the new ff_kernel, then A/B timing of kernel variants (f64, 1 Mi envs)
export TB_FF_SPIN_LIMIT_MS=1500
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2a_pytest.log
V="${1:-r1 base stride1 stride4 min4 min16}"
for rep in 1 2; do for v in $V; do
  echo -n "$v: "; TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1
done; done | tee gpurun_out/r2a_ab.log
TB_FF_DIAG_DUMP=1 TB_LIB_PATH=$PWD/build/variants/lib_diag.so timeout 300 python tools/time_kernels.py f64 1048576 > gpurun_out/r2a_diag.log 2>&1
grep -v "late landing" gpurun_out/r2a_diag.log | tail -40
