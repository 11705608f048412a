"""Host-side profile (cProfile) of GbmCVNNPricer.train at BASELINE configs[2] trainer-test size: shows whether the
training loop is bound by Python/launch overhead or by the device (Event.synchronize on the staging ring = waiting
for the GPU).  python tools/prof_train_host.py"""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd())
import torch
from spectralmc_b200.cvnn import make_cvnn
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import BlackScholes, BlackScholesConfig, SimulationParams
from spectralmc_b200.gbm_trainer import GbmCVNNPricer, TrainingConfig
from spectralmc_b200.numerical import Precision
from spectralmc_b200.sobol_sampler import BoundSpec, build_domain_bounds
bounds = build_domain_bounds(BlackScholes.Inputs, {k: BoundSpec(*b) for k, b in dict(
    X0=(0.001, 10_000.0), K=(0.001, 20_000.0), T=(0.0, 10.0), r=(-0.2, 0.2), d=(-0.2, 0.2), v=(0.0, 2.0)).items()}).unwrap()
sp = SimulationParams(timesteps=1, network_size=16, batches_per_mc_run=4096, threads_per_block=256, mc_seed=42, buffer_size=1, dtype=Precision.float32)
cfg = BlackScholesConfig(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
pricer = GbmCVNNPricer(cfg, bounds, make_cvnn(6, 16, seed=42))
pricer.train(TrainingConfig(num_batches=5, batch_size=1024)).unwrap()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
pricer.train(TrainingConfig(num_batches=300, batch_size=1024)).unwrap()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue())
