#!/usr/bin/env python
"""Secondary bench (BASELINE.json configs[2], trainer-test size): one GbmCVNNPricer training step
= Sobol batch of 1024 contracts -> [C, N] CF targets -> CVNN (6 -> 32 modReLU -> N) MSE/Adam step,
through the public trainer API, for the three routes of ``_torch_step``:
  torch        op-by-op torch autograd + torch.optim.Adam (what the reference does, gbm_trainer.py:819-835)
  torch-graph  the same torch step (capturable Adam) captured once and replayed as one CUDA graph — the route of
               networks the fused step does not cover (batch norms, residual blocks)
  fused      C-ABI smc_cvnn_train_step, ordinary launches
  graph      the same launches replayed as one CUDA graph (default)
Wall-clock over whole train() calls (host work, Sobol sampling and the final loss read-back included).
    python tools/bench_trainer.py [steps]
"""
import json
import sys
import time

import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spectralmc_b200.cvnn import make_cvnn
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import BlackScholes, BlackScholesConfig, SimulationParams
from spectralmc_b200.gbm_trainer import GbmCVNNPricer, TrainingConfig
from spectralmc_b200.numerical import Precision
from spectralmc_b200.sobol_sampler import BoundSpec, build_domain_bounds

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
bounds = build_domain_bounds(BlackScholes.Inputs, {k: BoundSpec(*b) for k, b in dict(
    X0=(0.001, 10_000.0), K=(0.001, 20_000.0), T=(0.0, 10.0), r=(-0.2, 0.2), d=(-0.2, 0.2), v=(0.0, 2.0)).items()}).unwrap()
for T, N, B in ((1, 16, 4096), (16, 128, 1024)):
    for route, kw in (("torch", dict(fused_step=False, cuda_graph=False)), ("torch-graph", dict(fused_step=False)),
                      ("fused", dict(cuda_graph=False)), ("graph", dict())):
        sp = SimulationParams(timesteps=T, network_size=N, batches_per_mc_run=B, threads_per_block=256, mc_seed=42, buffer_size=1,
                              dtype=Precision.float32)
        cfg = BlackScholesConfig(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
        pricer = GbmCVNNPricer(cfg, bounds, make_cvnn(6, N, seed=42), **kw)
        pricer.train(TrainingConfig(num_batches=3, batch_size=1024)).unwrap()
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(3):
            t0 = time.perf_counter()
            losses = pricer.train(TrainingConfig(num_batches=steps, batch_size=1024)).unwrap().losses
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / steps)
        # device-only time of the CVNN step (no simulation), CUDA events
        step_ms = None
        if route in ("fused", "graph"):
            f = pricer._fused
            real = torch.randn(1024, 6, device="cuda")
            tg = torch.randn(1024, N, dtype=torch.complex64, device="cuda")
            g = None
            if route == "graph":
                g = next(iter(pricer._graphs.values()))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(200):
                if g is not None:
                    g.graph.replay()
                else:
                    f.train_step(real, real, tg)
            b.record()
            b.synchronize()
            step_ms = a.elapsed_time(b) / 200
        print(json.dumps({"route": route, "T": T, "N": N, "B": B, "contracts_per_step": 1024, "ms_per_training_step": best * 1e3,
                          "cf_estimates_per_sec": 1024 / best, "cvnn_step_device_ms": step_ms, "last_loss": losses[-1]}), flush=True)
