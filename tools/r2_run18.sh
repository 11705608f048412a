(time python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_n1_final.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n1_final.log') if l.startswith('{')][0])
for k in ('value','ms_per_step','e2e','sustained','clocks','gpu_launches','kernels_per_step'): print(k, d[k])
print('roofline', d['roofline']['frac'], d['roofline']['detail']['fp32_issue']['frac'])
print(json.dumps(d['workloads'], indent=1))
g=d['gpu_reference_baseline']; print({k:g[k] for k in ('raw','normalize','ratio')})
print(d['cpu_baseline'])
PY
grep real gpurun_out/r2_bench_n1_final.log
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
