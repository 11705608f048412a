#!/usr/bin/env python
"""B-ref-CPU-sim baseline (BASELINE.json configs[0]): the reference's kernel
algorithm under Numba's CUDA simulator on the host CPU — fp64, 1 contract, 12 timesteps,
network_size=16, batches_per_mc_run=64 — followed by numpy.fft.fft(axis=1) + mean(axis=0).
The simulator runs threads under the GIL, i.e. effectively ONE core.  Not the product.

usage: NUMBA_ENABLE_CUDASIM=1 python baseline/ref_cpu_cudasim.py   (the env var is set below if absent)
"""
import json
import math
import os
import time

os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
import numpy as np
from numba import cuda


@cuda.jit
def path_kernel(io, timesteps, dt, X0, r, d, v):
    idx = cuda.grid(1)
    if idx < io.shape[1]:
        sdt = math.sqrt(dt)
        x = X0
        mu = r - d - 0.5 * v * v
        for i in range(timesteps):
            x *= math.exp(mu * dt + v * (io[i, idx] * sdt))
            io[i, idx] = x


def main():
    T, N, B = 12, 16, 64
    P = N * B
    X0, K, Tm, r, d, v = 100.0, 100.0, 1.0, 0.05, 0.0, 0.2
    reps, total = 0, 0.0
    while total < 5.0:
        z = np.random.default_rng(42 + reps).standard_normal((T, P))
        t0 = time.perf_counter()
        path_kernel[(P + 255) // 256, 256](z, T, Tm / T, X0, r, d, v)
        put = math.exp(-r * Tm) * np.maximum(K - z[-1], 0.0)
        cf = np.mean(np.fft.fft(put.reshape(B, N), axis=1), axis=0)
        total += time.perf_counter() - t0
        reps += 1
    print(json.dumps({"impl": "ref_cpu_cudasim", "config": "c1: fp64, T=12, N=16, B=64", "reps": reps, "seconds": total,
                      "path_steps_per_sec": reps * T * P / total, "cf_estimates_per_sec": reps / total,
                      "host_cores": os.cpu_count(), "cores_used": 1, "put_price": float(cf[0].real) / N}))


if __name__ == "__main__":
    main()
