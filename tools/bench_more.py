#!/usr/bin/env python
"""Secondary measurements (single GPU): the other BASELINE.json configs and kernel variants.
Prints one JSON line per measurement.  Not the driver's bench (that is bench.py)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from spectralmc_b200 import _cabi

dev = torch.device("cuda", 0)
CANON = (100.0, 100.0, 1.0, 0.05, 0.0, 0.2)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def fused(name, C, T, N, B, dtype, scheme=_cabi.SMC_LOG_EULER, norm=_cabi.SMC_RAW, reps=5):
    rows = np.tile(np.asarray(CANON), (C, 1))
    if C > 1:
        from oracle.sobol import sobol_contracts  # bench-side input synthesis only
        rows = sobol_contracts(C, seed=42)
    contracts = torch.tensor(rows, dtype=torch.float64, device=dev)
    args = _cabi.make_fused_args(contracts, C, T, N, B, dtype, scheme, norm, 7, 0)
    ws = torch.empty(_cabi.LIB.smc_cf_fused_workspace_bytes(_cabi.byref(args)) + 256, dtype=torch.uint8, device=dev)
    ms = timeit(lambda: _cabi.cf_fused(args, dev, dtype, ws), reps=reps)
    steps = float(C) * T * N * B
    print(json.dumps({"what": name, "C": C, "T": T, "N": N, "B": B, "dtype": str(dtype), "ms": ms, "path_steps_per_sec": steps / ms * 1e3,
                      "cf_estimates_per_sec": C / ms * 1e3, "workspace_MB": ws.numel() / 1e6}), flush=True)


f32, f64 = torch.float32, torch.float64
fused("c2 RAW log-euler", 1, 252, 128, 65536, f32)
fused("c2 NORMALIZE log-euler", 1, 252, 128, 65536, f32, norm=_cabi.SMC_NORMALIZE)
fused("c2 RAW simple-euler", 1, 252, 128, 65536, f32, scheme=_cabi.SMC_SIMPLE_EULER)
fused("c2 RAW log-euler stepwise-exp", 1, 252, 128, 65536, f32, scheme=_cabi.SMC_LOG_EULER_STEPWISE)
fused("c2 x 8 contracts RAW", 8, 252, 128, 65536, f32, reps=3)
fused("c3 trainer-test size (1024 Sobol contracts, T=1,N=16,B=4096)", 1024, 1, 16, 4096, f32)
fused("c3 trainer-test size NORMALIZE", 1024, 1, 16, 4096, f32, norm=_cabi.SMC_NORMALIZE)
fused("c1 shape fp64 (1 contract, T=12,N=16,B=64)", 1, 12, 16, 64, f64)
fused("c2 fp64 RAW", 1, 252, 128, 65536, f64, reps=3)
fused("c4 fp64: 512 Sobol contracts, T=365, N=256, B=4096", 512, 365, 256, 4096, f64, reps=2)

# CF methods on a materialised payoff matrix (c2 shape)
mat = torch.rand((65536, 128), dtype=f32, device=dev)
for method, name in ((_cabi.SMC_CF_MEAN_THEN_FFT, "cf_fft_mean mean-then-FFT"), (_cabi.SMC_CF_ROW_FFT, "cf_fft_mean row-FFT")):
    ms = timeit(lambda: _cabi.cf_fft_mean(mat, method))
    print(json.dumps({"what": name, "B": 65536, "N": 128, "ms": ms, "GBps": mat.numel() * 4 / ms / 1e6}), flush=True)
ms = timeit(lambda: torch.fft.fft(mat, dim=1).mean(dim=0))
print(json.dumps({"what": "torch.fft.fft(dim=1).mean(dim=0) (cuFFT, the reference's formulation)", "ms": ms, "GBps": mat.numel() * 4 / ms / 1e6}), flush=True)

# c3: Sobol batch of 1024 contracts -> targets -> CVNN (6 -> 32 modReLU -> N) Adam step, through the trainer API
import time

from spectralmc_b200.cvnn import make_cvnn
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import BlackScholesConfig, SimulationParams
from spectralmc_b200.gbm_trainer import GbmCVNNPricer, TrainingConfig
from spectralmc_b200.numerical import Precision
from spectralmc_b200.sobol_sampler import BoundSpec, build_domain_bounds
from spectralmc_b200.gbm import BlackScholes

bounds = build_domain_bounds(BlackScholes.Inputs, {k: BoundSpec(*b) for k, b in dict(
    X0=(0.001, 10_000.0), K=(0.001, 20_000.0), T=(0.0, 10.0), r=(-0.2, 0.2), d=(-0.2, 0.2), v=(0.0, 2.0)).items()}).unwrap()
for name, T, N, B, steps in (("c3 training step, trainer-test size (T=1,N=16,B=4096)", 1, 16, 4096, 10),
                             ("c3 training step, c2-size simulation (T=252,N=128,B=65536)", 252, 128, 65536, 2)):
    sp = SimulationParams(timesteps=T, network_size=N, batches_per_mc_run=B, threads_per_block=256, mc_seed=42, buffer_size=1, dtype=Precision.float32)
    cfg = BlackScholesConfig(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
    pricer = GbmCVNNPricer(cfg, bounds, make_cvnn(6, N, seed=42))
    pricer.train(TrainingConfig(num_batches=1, batch_size=1024)).unwrap()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    losses = pricer.train(TrainingConfig(num_batches=steps, batch_size=1024)).unwrap().losses
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(json.dumps({"what": name, "ms_per_training_step": dt * 1e3, "contracts_per_step": 1024, "cf_estimates_per_sec": 1024 / dt,
                      "path_steps_per_sec": 1024.0 * T * N * B / dt, "last_loss": losses[-1]}), flush=True)
