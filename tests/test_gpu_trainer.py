"""The trainer data path on the fused device path: the reference's lock-step / snapshot tests
(tests/test_gbm_trainer.py:122-162,182-263: two trainers stay bit-identical through
train / snapshot / restore) and the prediction smoke test (:302-320)."""

from __future__ import annotations

import math

import numpy as np
import pytest
import torch

from oracle import gbm as ogbm
from oracle import philox
from oracle.sobol import sobol_contracts
from spectralmc_b200.cvnn import make_cvnn
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import BlackScholes
from spectralmc_b200.gbm_trainer import GbmCVNNPricer, TrainingConfig, _split_inputs
from spectralmc_b200.numerical import Precision
from tests.helpers import expect_success, make_black_scholes_config, make_domain_bounds, make_simulation_params, rel_max

pytestmark = pytest.mark.gpu


def _pricer(precision=Precision.float32, *, seed=42, N=16, B=2**12, T=1, norm=ForwardNormalization.RAW):
    sp = make_simulation_params(timesteps=T, network_size=N, batches_per_mc_run=B, threads_per_block=256, mc_seed=seed,
                                buffer_size=1, dtype=precision)
    cfg = make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=norm)
    cvnn = make_cvnn(6, N, seed=seed, dtype=precision.to_torch())
    return GbmCVNNPricer(cfg, make_domain_bounds(), cvnn)


def _max_param_diff(a, b) -> float:
    return max(float((p - q).abs().max()) for p, q in zip(a._cvnn.parameters(), b._cvnn.parameters()))


@pytest.mark.parametrize("precision", [Precision.float32, Precision.float64])
def test_lockstep_training(precision) -> None:
    a, b = _pricer(precision), _pricer(precision)
    la = expect_success(a.train(TrainingConfig(num_batches=3, batch_size=8)))
    lb = expect_success(b.train(TrainingConfig(num_batches=3, batch_size=8)))
    assert la == lb and all(math.isfinite(x) for x in la)
    assert _max_param_diff(a, b) == 0.0
    sa, sb = a.snapshot(), b.snapshot()
    assert sa.sobol_skip == sb.sobol_skip == 24 and sa.global_step == 3
    assert sa.cfg.sim_params.skip == 24  # one normal matrix per contract priced


def test_snapshot_cycle_deterministic() -> None:
    """train 2 -> snapshot -> train 2 more  ==  restore(snapshot) -> train 2."""
    a = _pricer()
    expect_success(a.train(TrainingConfig(num_batches=2, batch_size=8)))
    snap = a.snapshot()
    expect_success(a.train(TrainingConfig(num_batches=2, batch_size=8)))
    b = GbmCVNNPricer.restore(snap, make_domain_bounds(), make_cvnn(6, 16, seed=1))
    expect_success(b.train(TrainingConfig(num_batches=2, batch_size=8)))
    assert _max_param_diff(a, b) == 0.0
    assert b.snapshot().global_step == 4 and b.snapshot().sobol_skip == 32


def test_targets_match_oracle_for_a_sobol_batch() -> None:
    """The [C, N] targets of one training step vs the oracle, contract by contract
    (Sobol contracts with mc_seed, matrix k for contract k)."""
    p = _pricer(Precision.float64, seed=42, N=16, B=64, T=3)
    rows = sobol_contracts(8, seed=42)
    got = expect_success(p.targets(rows)).cpu().numpy()
    for i, row in enumerate(rows):
        z = philox.normals_matrix(3, 16 * 64, np.float64, 42, i)
        ref, _ = ogbm.simulate_fft(ogbm.Contract(*row), z, 16, normalization=ogbm.RAW)
        assert rel_max(got[i], ref) <= 1e-12 or np.max(np.abs(ref)) == 0.0


def test_split_inputs_and_predict_price_smoke() -> None:
    p = _pricer()
    inputs = [BlackScholes.Inputs(X0=100, K=100, T=1.0, r=0.05, d=0.0, v=0.2), BlackScholes.Inputs(X0=50, K=60, T=0.5, r=0.0, d=0.0, v=0.4)]
    real, imag = _split_inputs(inputs, dtype=torch.float32, device=torch.device("cuda"))
    assert real.shape == (2, 6) and float(real[1, 1]) == 60.0 and float(imag.abs().max()) == 0.0
    prices = p.predict_price(inputs)
    assert len(prices) == 2 and all(math.isfinite(x) for x in prices)


def test_training_reduces_the_loss() -> None:
    p = _pricer(Precision.float32, N=16, B=2**10)
    losses = expect_success(p.train(TrainingConfig(num_batches=30, batch_size=32, learning_rate=1e-2)))
    assert np.mean(losses[-5:]) < np.mean(losses[:5])
