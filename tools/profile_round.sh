#!/bin/bash
# Run on the GPU box (gpurun): launch list of a short bench run + full captures of the two step kernels (f64, 1 Mi envs).
# usage: tools/profile_round.sh r1   (then summarise with tools/ncu_summary.py / ncu_lines.py into profiles/)
R=${1:-r1}
set -x
python bench.py --steps 26 --warmup 26 --skip-cpu-baseline > gpurun_out/bench_plain_$R.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 26 --warmup 26 --skip-cpu-baseline > gpurun_out/ncu_launch_$R.log 2>&1
python tools/prof_swing.py f64 1048576 > gpurun_out/plain_f64_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ff_kernel -s 25 -c 1 -o gpurun_out/prof_${R}_ff_f64 \
    python tools/prof_swing.py f64 1048576 > gpurun_out/ncu_ff_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 1 -o gpurun_out/prof_${R}_step_f64 \
    python tools/prof_swing.py f64 1048576 > gpurun_out/ncu_step_$R.log 2>&1
tail -n 2 gpurun_out/ncu_ff_$R.log gpurun_out/ncu_step_$R.log
# Tennisbot-v0: one step_kernel launch once the episodes have desynchronised (1 Mi envs, f64)
python tools/prof_swing.py f64 1048576 Tennisbot-v0 1300 > gpurun_out/plain_hit_f64_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1250 -c 1 -o gpurun_out/prof_${R}_hit_step_f64 \
    python tools/prof_swing.py f64 1048576 Tennisbot-v0 1300 > gpurun_out/ncu_hit_$R.log 2>&1
tail -n 2 gpurun_out/ncu_hit_$R.log
