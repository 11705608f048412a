(time python -m pytest tests -q -m gpu) > gpurun_out/r2_pytest_all.log 2>&1; tail -8 gpurun_out/r2_pytest_all.log
