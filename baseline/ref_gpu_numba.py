#!/usr/bin/env python
"""B-ref-GPU baseline ("Numba + torch-as-CuPy"): the reference's pipeline shape on
one B200, for the ">= 10x the reference" comparison.  NOT the product and not imported by it.

The unmodified reference cannot run here (CuPy is not installed and not in the wheelhouse), so this
script restates its per-contract flow 1:1 with the same launch structure and host syncs:
  normals    torch.randn((T, P))                         for cp.random...standard_normal (async_normals.py:214-215)
  paths      a Numba @cuda.jit kernel, one thread = one path, float64 scalar arguments, in place
             (the algorithm of gbm.py:241-257; block size 256, grid = ceil(P/256), gbm.py:101-103)
  engine     linspace/exp (gbm.py:429-431), optional mean+rescale (:437-438), payoff (:467-474)
  CF         torch.fft.fft(mat, dim=1).mean(dim=0)       for cp.mean(cp.fft.fft(mat, axis=1), axis=0)
             (gbm_trainer.py:814-817; torch.fft is cuFFT, the library CuPy calls)
  syncs      after the kernel, after normalisation, after the payoff (gbm.py:434,442,476)

usage: python baseline/ref_gpu_numba.py [--T 252 --N 128 --B 65536 --reps 5 --normalize 0]
Prints one JSON line with path-steps/s.
"""
import argparse
import json
import math
import time

import torch
from numba import cuda


@cuda.jit
def path_kernel(io, timesteps, dt, X0, r, d, v, log_flag):
    idx = cuda.grid(1)
    if idx < io.shape[1]:
        sdt = math.sqrt(dt)
        x = X0
        if log_flag:
            mu = r - d - 0.5 * v * v
            for i in range(timesteps):
                x *= math.exp(mu * dt + v * (io[i, idx] * sdt))
                io[i, idx] = x
        else:
            mu = r - d
            for i in range(timesteps):
                dw = io[i, idx] * sdt
                x = abs(x + mu * x * dt + v * x * dw)
                io[i, idx] = x


def one_contract(T, N, B, normalize, dtype, X0=100.0, K=100.0, Tm=1.0, r=0.05, d=0.0, v=0.2):
    P = N * B
    sims = torch.randn((T, P), dtype=dtype, device="cuda")
    torch.cuda.synchronize()  # the generator's stream.synchronize() before hand-out (async_normals.py:230)
    dt = Tm / T
    path_kernel[(P + 255) // 256, 256](cuda.as_cuda_array(sims), T, dt, X0, r, d, v, True)
    times = torch.linspace(dt, Tm, T, dtype=dtype, device="cuda")
    forwards = X0 * torch.exp((r - d) * times)
    df = torch.exp(-r * times)
    cuda.synchronize()
    if normalize:
        row_means = sims.mean(dim=1, keepdim=True).squeeze()
        sims *= (forwards / row_means).unsqueeze(1)
    torch.cuda.synchronize()
    terminal = sims[-1]
    put = df[-1] * torch.clamp(K - terminal, min=0)
    call = df[-1] * torch.clamp(terminal - K, min=0)  # noqa: F841  (the reference computes both)
    torch.cuda.synchronize()
    cf = torch.fft.fft(put.reshape(B, N), dim=1).mean(dim=0)
    torch.cuda.synchronize()
    return cf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=252)
    ap.add_argument("--N", type=int, default=128)
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--normalize", type=int, default=0)
    ap.add_argument("--dtype", default="float32")
    a = ap.parse_args()
    dtype = getattr(torch, a.dtype)
    for _ in range(2):
        cf = one_contract(a.T, a.N, a.B, a.normalize, dtype)
    t0 = time.perf_counter()
    for _ in range(a.reps):
        cf = one_contract(a.T, a.N, a.B, a.normalize, dtype)
    dt = (time.perf_counter() - t0) / a.reps
    print(json.dumps({"impl": "ref_gpu_numba_torch", "T": a.T, "N": a.N, "B": a.B, "dtype": a.dtype, "normalize": a.normalize,
                      "ms_per_contract": dt * 1e3, "path_steps_per_sec": a.T * a.N * a.B / dt, "put_price": float(cf[0].real) / a.N,
                      "device": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
