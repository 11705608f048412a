#!/usr/bin/env bash
# Final scaling record of the round (run under `gpurun --gpus 8`): bench.py exactly as the driver launches it.
cd "$(dirname "$0")/.."
run() { # N extra-flags...
  local n=$1; shift
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 "$@"
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus "$n" --steps 20 --warmup 5 "$@"; fi
}
run 8 --no-workloads --no-cpu-baseline > gpurun_out/r2_final2_n8.log 2> gpurun_out/r2_final2_n8.err
run 4 --no-workloads --no-cpu-baseline --no-c5cut > gpurun_out/r2_final2_n4.log 2> gpurun_out/r2_final2_n4.err
run 2 --no-workloads --no-cpu-baseline --no-c5cut > gpurun_out/r2_final2_n2.log 2> gpurun_out/r2_final2_n2.err
run 1 --no-workloads --no-cpu-baseline > gpurun_out/r2_final2_n1.log 2> gpurun_out/r2_final2_n1.err
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r2_final2_multi_tests.log 2>&1
tail -n 3 gpurun_out/r2_final2_multi_tests.log
for n in 8 4 2 1; do python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2_final2_n{n}.log").read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["sustained"]["ms_per_step"], (d.get("parity_check") or {}).get("ok"),
          json.dumps(d.get("strong"))[:700])
except Exception as e:
    print(n, "FAILED", e)
PY
done
