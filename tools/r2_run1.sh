set -x
python -m pytest tests/test_gpu_fused.py tests/test_gpu_paths.py tests/test_gpu_canaries.py -x -q > gpurun_out/r2_pytest_a.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_a.log; tail -5 gpurun_out/r2_pytest_a.log
{
SMC_LIB=tools/tune/lib_r1.so python tools/bench_raw.py c2 c2x8 c3 c4s c2s8
python tools/bench_raw.py c2 c2x8 c3 c4s c2s8
SMC_TARGET_TILES=32768 python tools/bench_raw.py c2 c2s8
SMC_TARGET_TILES=8192 python tools/bench_raw.py c2
SMC_STATIC_SCHEDULE=1 python tools/bench_raw.py c2
SMC_SEPARATE_FINALIZE=1 python tools/bench_raw.py c2
SMC_NORM=1 SMC_LIB=tools/tune/lib_r1.so python tools/bench_raw.py c2
SMC_NORM=1 python tools/bench_raw.py c2
SMC_SCHEME=1 SMC_LIB=tools/tune/lib_r1.so python tools/bench_raw.py c2
SMC_SCHEME=1 python tools/bench_raw.py c2
} > gpurun_out/r2_ab1.log 2>&1
cat gpurun_out/r2_ab1.log
