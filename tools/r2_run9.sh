timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_pytest_multi.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_multi.log; tail -5 gpurun_out/r2_pytest_multi.log
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/r2_bench_n2.log 2>&1; tail -c 4500 gpurun_out/r2_bench_n2.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --collective nccl --no-c5cut) > gpurun_out/r2_bench_n2_nccl.log 2>&1; tail -c 1500 gpurun_out/r2_bench_n2_nccl.log
python bench.py --steps 20 --warmup 5 --no-workloads --no-cpu-baseline > gpurun_out/r2_bench_n1b.log 2>&1; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n1b.log') if l.startswith('{')][0]); print({k:d[k] for k in ('value','ms_per_step','e2e','sustained')})"
