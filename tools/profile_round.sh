#!/bin/bash
# Run on the GPU box (gpurun): the numbers and captures of a round.  usage: tools/profile_round.sh r2
R=${1:-r2}
export TB_FF_SPIN_LIMIT_MS=20000
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${R}_bench_k20.json 2> gpurun_out/${R}_bench_k20.err
python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err
python bench.py --precision f32 --skip-cpu-baseline --extras off > gpurun_out/${R}_bench_f32.json 2>/dev/null
python bench.py --env Tennisbot-v0 --skip-cpu-baseline --extras off > gpurun_out/${R}_bench_hit.json 2>/dev/null
python bench.py --env Tennisbot-v0 --envs-per-gpu 65536 --skip-cpu-baseline --extras off > gpurun_out/${R}_bench_hit_65536.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
# A/B on identical work: round 1's kernels vs this round's, on the i.i.d. action ring and on round 1's 4-batch ring
for ring in 32 4; do for v in r1 base; do
  echo -n "ring $ring $v: "; TB_RING=$ring TB_LIB_PATH=$PWD/build/variants/lib_$v.so python tools/time_steps.py f64 1048576 3 2>&1 | tail -1
done; done > gpurun_out/${R}_ab.txt
echo -n "racket_court_contact=1: " >> gpurun_out/${R}_ab.txt; TB_PARAMS=racket_court_contact=1 python tools/time_steps.py f64 1048576 2 2>&1 | tail -1 >> gpurun_out/${R}_ab.txt
python tools/time_e2e.py 1048576 > gpurun_out/${R}_e2e_modes.txt 2>&1
cat gpurun_out/${R}_ab.txt
# launch list of a short bench run, then full captures of the two step kernels (each only after the plain run exited 0)
python bench.py --steps 26 --warmup 26 --skip-cpu-baseline --extras off > gpurun_out/bench_plain_$R.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 26 --warmup 26 --skip-cpu-baseline --extras off > gpurun_out/ncu_launch_$R.log 2>&1
python tools/prof_swing.py f64 1048576 > gpurun_out/plain_f64_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ff_kernel -s 25 -c 1 -f -o gpurun_out/prof_${R}_ff_f64 \
    python tools/prof_swing.py f64 1048576 > gpurun_out/ncu_ff_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 1 -f -o gpurun_out/prof_${R}_step_f64 \
    python tools/prof_swing.py f64 1048576 > gpurun_out/ncu_step_$R.log 2>&1
python tools/prof_swing.py f64 1048576 Tennisbot-v0 1300 > gpurun_out/plain_hit_f64_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1250 -c 1 -f -o gpurun_out/prof_${R}_hit_step_f64 \
    python tools/prof_swing.py f64 1048576 Tennisbot-v0 1300 > gpurun_out/ncu_hit_$R.log 2>&1
tail -n 1 gpurun_out/ncu_ff_$R.log gpurun_out/ncu_step_$R.log gpurun_out/ncu_hit_$R.log
