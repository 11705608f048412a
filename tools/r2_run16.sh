python -m pytest tests/test_gpu_fused.py tests/test_gpu_canaries.py tests/test_gpu_trainer.py -x -q > gpurun_out/r2_pytest_g.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_g.log; tail -3 gpurun_out/r2_pytest_g.log
python tools/bench_raw.py c3 c3t2 c3t3 c2 > gpurun_out/r2_ab16.log 2>&1; cat gpurun_out/r2_ab16.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_pytest_multi2.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_multi2.log; tail -4 gpurun_out/r2_pytest_multi2.log
