#!/bin/bash
# A/B of kernel variants built into build/variants/lib_<name>.so (TB_LIB_PATH override), interleaved runs.
# usage: tools/ab_variants.sh "base guard" [reps] [precision] [n_envs]
V=${1:-base}; R=${2:-3}; P=${3:-f64}; N=${4:-1048576}
for r in $(seq $R); do for v in $V; do
  echo -n "$v: "; TB_LIB_PATH=$PWD/build/variants/lib_$v.so python tools/time_kernels.py $P $N 2>&1 | grep "^25 \|^12 " | tr '\n' ' '; echo
done; done
