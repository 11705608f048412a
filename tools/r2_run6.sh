python -m pytest tests/test_gpu_fused.py tests/test_gpu_paths.py tests/test_gpu_canaries.py -x -q > gpurun_out/r2_pytest_c.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_c.log; tail -3 gpurun_out/r2_pytest_c.log
{
SMC_LIB=tools/tune/lib_r1.so python tools/bench_raw.py c2 c2x8 c2s8 c3 c4s
python tools/bench_raw.py c2 c2x8 c2s8 c3 c4s
for v in o0 o1 inl c4 c6 u2; do SMC_LIB=tools/tune/lib_v_$v.so python tools/bench_raw.py c2 c2x8 c2s8; done
SMC_TARGET_TILES=32768 python tools/bench_raw.py c2 c2s8
SMC_TARGET_TILES=8192 python tools/bench_raw.py c2 c2s8
SMC_NORM=1 python tools/bench_raw.py c2
SMC_NORM=1 SMC_LIB=tools/tune/lib_r1.so python tools/bench_raw.py c2
} > gpurun_out/r2_ab6.log 2>&1
grep -v "^+" gpurun_out/r2_ab6.log | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['lib'][:24].ljust(24), d['shape'].ljust(5), d['norm'], d['ms_min'], d['ms_med'], d['env'])
"
