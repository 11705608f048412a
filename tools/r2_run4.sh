SMC_REPS=4 python tools/bench_raw.py c2 > gpurun_out/r2_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o gpurun_out/r2_prof_step_c2 env SMC_REPS=4 python tools/bench_raw.py c2 > gpurun_out/r2_ncu4.log 2>&1
tail -3 gpurun_out/r2_ncu4.log
