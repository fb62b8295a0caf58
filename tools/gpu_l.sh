#!/bin/bash
export TB_FF_SPIN_LIMIT_MS=3000
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for rep in 1 2; do for v in prev base; do
  echo -n "$v: "; TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f64 1048576 3 2>&1 | tail -1
done; done
for v in prev base; do echo -n "hit $v: "; TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_hit.py f64 1048576 2>&1 | tail -1; done
for v in prev base; do echo -n "f32 $v: "; TB_LIB_PATH=$PWD/build/variants/lib_$v.so timeout 300 python tools/time_steps.py f32 1048576 3 2>&1 | tail -1; done
