python -m pytest tests/test_gpu_fused.py tests/test_gpu_paths.py tests/test_gpu_canaries.py tests/test_gpu_normals.py -x -q > gpurun_out/r2_pytest_d.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_d.log; tail -12 gpurun_out/r2_pytest_d.log
{
SMC_LIB=tools/tune/lib_r1.so python tools/bench_raw.py c2 c2x8 c2s8 c3 c3t2 c3t3
python tools/bench_raw.py c2 c2x8 c2s8 c3 c3t2 c3t3
SMC_SHORT_GROUPED=0 python tools/bench_raw.py c3
for v in c6 c6u2 u2 c6o0 c6o1 c6inl; do SMC_LIB=tools/tune/lib_v_$v.so python tools/bench_raw.py c2 c2x8 c2s8; done
for t in 4096 8192 12288 24576; do SMC_TARGET_TILES=$t SMC_LIB=tools/tune/lib_v_c6.so python tools/bench_raw.py c2 c2s8; done
for t in 4096 8192 12288; do SMC_TARGET_TILES=$t SMC_LIB=tools/tune/lib_v_c6u2.so python tools/bench_raw.py c2 c2s8; done
} > gpurun_out/r2_ab7.log 2>&1
grep -v "^+" gpurun_out/r2_ab7.log | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['lib'][:24].ljust(24), d['shape'].ljust(5), d['norm'], d['ms_min'], d['ms_med'], d['env'])
"
