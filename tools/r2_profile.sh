#!/usr/bin/env bash
# Round-2 profile capture (run under gpurun; every ncu command follows a plain run of the same command that exited 0).
set -x
cd "$(dirname "$0")/.."
for what in fused fused_f64 fused_c3; do
  python tools/prof_fused.py $what 3 > gpurun_out/r2_prof_plain_$what.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_$what python tools/prof_fused.py $what 3 > gpurun_out/r2_prof_ncu_$what.log 2>&1
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2_launches_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2_launches_ncu.log 2>&1
tail -n 2 gpurun_out/r2_prof_ncu_fused.log; tail -n 2 gpurun_out/r2_launches_ncu.log
