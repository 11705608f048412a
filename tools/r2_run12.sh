(time python -m pytest tests/test_gpu_stream_battery.py -x -q) > gpurun_out/r2_pytest_battery.log 2>&1; tail -15 gpurun_out/r2_pytest_battery.log
