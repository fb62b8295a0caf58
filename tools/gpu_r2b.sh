#!/bin/bash
# round 2, GPU call B: parity suite on ff_kernel v3 (lean contact path, overlapped dense finishing), A/B timing (f64, 1 Mi envs)
export TB_FF_SPIN_LIMIT_MS=1500
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2b_pytest.log
V="${1:-r1 base nolean unroll1 stride4 stride8}"
for rep in 1 2; do for v in $V; do
  echo -n "$v: "; TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1
done; done | tee gpurun_out/r2b_ab.log
echo "--- 4-batch ring (round-1 workload)"
for v in r1 base; do echo -n "$v: "; TB_RING=4 TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1; done | tee -a gpurun_out/r2b_ab.log
TB_FF_DIAG_DUMP=1 TB_LIB_PATH=$PWD/build/variants/lib_diag.so timeout 300 python tools/time_kernels.py f64 1048576 > gpurun_out/r2b_diag.log 2>&1
grep -v "late landing" gpurun_out/r2b_diag.log | tail -30
grep "late landing" gpurun_out/r2b_diag.log | sort -t- -k2 -n | head -12
