{
python tools/timeline.py c2 c2x8
SMC_SEGMENTS=1 python tools/timeline.py c2
SMC_GRID_CAP=592 python tools/bench_raw.py c2
SMC_GRID_CAP=444 python tools/bench_raw.py c2
SMC_SEGMENTS=1 SMC_GRID_CAP=592 python tools/bench_raw.py c2
} > gpurun_out/r2_ab5.log 2>&1
cat gpurun_out/r2_ab5.log
