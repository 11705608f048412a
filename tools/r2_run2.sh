python -m pytest tests/test_gpu_fused.py tests/test_gpu_paths.py tests/test_gpu_canaries.py -x -q > gpurun_out/r2_pytest_b.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_b.log; tail -3 gpurun_out/r2_pytest_b.log
{
SMC_LIB=tools/tune/lib_r1.so python tools/bench_raw.py c2 c2x8 c2s8 c3
for v in l0o0 l0o1 l0o2 l1o0 l1o1 l1o2 l1o2c4 l1o2c6 l1o2u2 l0o2c4; do SMC_LIB=tools/tune/lib_v_$v.so python tools/bench_raw.py c2 c2x8 c2s8; done
python tools/bench_raw.py c3 c4s
SMC_SEGMENTS=1 python tools/bench_raw.py c2 c2x8 c2s8
SMC_SEG_ITEMS=2048 python tools/bench_raw.py c2 c2x8 c2s8
SMC_SEG_ITEMS=512 python tools/bench_raw.py c2 c2x8 c2s8
SMC_SEGMENTS=4 python tools/bench_raw.py c2 c2s8
SMC_STATIC_SCHEDULE=1 python tools/bench_raw.py c2
} > gpurun_out/r2_ab2.log 2>&1
grep -v "^+" gpurun_out/r2_ab2.log | cut -c1-200
