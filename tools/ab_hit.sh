#!/bin/bash
# A/B of Tennisbot-v0 step variants (tennisbot_rl_b200/variants/lib_<name>.so), interleaved.  usage: ab_hit.sh "base hf3" reps prec n
V=${1:-base}; R=${2:-2}; P=${3:-f64}; N=${4:-1048576}
for r in $(seq $R); do for v in $V; do
  echo -n "$v: "; TB_LIB_PATH=$PWD/tennisbot_rl_b200/variants/lib_$v.so python tools/time_hit.py $P $N 2>&1 | tail -1
done; done
