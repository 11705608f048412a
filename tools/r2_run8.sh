python -m pytest tests/test_gpu_fused.py tests/test_gpu_paths.py tests/test_gpu_canaries.py tests/test_gpu_normals.py -x -q > gpurun_out/r2_pytest_e.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_e.log; tail -3 gpurun_out/r2_pytest_e.log
python tools/bench_raw.py c2 c2x8 c2s8 c3 c3t2 c3t3 c4s > gpurun_out/r2_ab8.log 2>&1; cat gpurun_out/r2_ab8.log
(time python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_n1.log 2>&1; tail -c 6000 gpurun_out/r2_bench_n1.log
