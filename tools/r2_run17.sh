for n in 8 4; do
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5) > gpurun_out/r2_bench_n$n.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n$n.log') if l.startswith('{')][0])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','collective','kernels_per_step','clocks')})
print('e2e', d['e2e']['ms_per_step'], 'sustained', d['sustained']['ms_per_step'])
print('parity', d['parity_check'])
print('strong', json.dumps(d['strong']))
PY
tail -4 gpurun_out/r2_bench_n$n.log | grep real
done
