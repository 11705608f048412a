python -m pytest tests/test_gpu_normals.py tests/test_gpu_fused.py -x -q -k "float64 or f64 or prec" > gpurun_out/r2_pytest_f.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_f.log; tail -3 gpurun_out/r2_pytest_f.log
tools/tune/pipe_overlap_f64 > gpurun_out/r2_pipe_overlap_f64.txt 2>&1; cat gpurun_out/r2_pipe_overlap_f64.txt
{
python tools/bench_raw.py c2f64 c4s
for v in f64full f64c5 f64c3 f64u2; do SMC_LIB=tools/tune/lib_v_$v.so python tools/bench_raw.py c2f64 c4s; done
SMC_SCHEME=1 python tools/bench_raw.py c2f64
} > gpurun_out/r2_ab10.log 2>&1
grep -v "^+" gpurun_out/r2_ab10.log | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['lib'][:24].ljust(24), d['shape'].ljust(5), d['norm'], d['ms_min'], d['ms_med'], d['env'])
"
