#!/usr/bin/env python
"""A few eager (non-graph) fused trainer steps at BASELINE configs[2] trainer-test size, for
`ncu --metrics gpu__time_duration.sum` launch lists (profiles/r1_cvnn_step_launches.txt)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spectralmc_b200.cvnn import make_cvnn
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import BlackScholes, BlackScholesConfig, SimulationParams
from spectralmc_b200.gbm_trainer import GbmCVNNPricer, TrainingConfig
from spectralmc_b200.numerical import Precision
from spectralmc_b200.sobol_sampler import BoundSpec, build_domain_bounds

bounds = build_domain_bounds(BlackScholes.Inputs, {k: BoundSpec(*b) for k, b in dict(
    X0=(0.001, 10_000.0), K=(0.001, 20_000.0), T=(0.0, 10.0), r=(-0.2, 0.2), d=(-0.2, 0.2), v=(0.0, 2.0)).items()}).unwrap()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sp = SimulationParams(timesteps=1, network_size=N, batches_per_mc_run=4096, threads_per_block=256, mc_seed=42, buffer_size=1,
                      dtype=Precision.float32)
cfg = BlackScholesConfig(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
pricer = GbmCVNNPricer(cfg, bounds, make_cvnn(6, N, seed=42), cuda_graph=False)
print(pricer.train(TrainingConfig(num_batches=4, batch_size=1024)).unwrap().losses)
torch.cuda.synchronize()
