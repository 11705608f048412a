(time python -m pytest tests -x -q -m gpu) > gpurun_out/r2_pytest_all.log 2>&1; tail -12 gpurun_out/r2_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
