"""The trainer data path on the fused device path: the reference's lock-step / snapshot tests
(tests/test_gbm_trainer.py:122-162,182-263: two trainers stay bit-identical through
train / snapshot / restore) and the prediction smoke test (:302-320)."""

from __future__ import annotations

import copy
import math
import warnings

import numpy as np
import pytest
import torch

from oracle import gbm as ogbm
from oracle import philox
from oracle.sobol import sobol_contracts
from spectralmc_b200.cvnn import make_cvnn
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import BlackScholes
from spectralmc_b200.gbm_trainer import (
    GbmCVNNPricer,
    _split_inputs,
    build_gbm_cvnn_pricer_config,
    build_training_config,
)
from spectralmc_b200.numerical import Precision
from tests.helpers import expect_failure, expect_success, make_black_scholes_config, make_domain_bounds, make_simulation_params, rel_max

pytestmark = pytest.mark.gpu


LEARNING_RATE = 1.0e-2  # the reference's tests/test_gbm_trainer.py:70
PRECISIONS = (Precision.float32, Precision.float64)


def _pricer(precision=Precision.float32, *, seed=42, N=16, B=2**12, T=1, norm=ForwardNormalization.RAW, **kw):
    """The reference's ``_make_gbm_trainer`` (tests/test_gbm_trainer.py:122-162)."""
    sp = make_simulation_params(timesteps=T, network_size=N, batches_per_mc_run=B, threads_per_block=256, mc_seed=seed,
                                buffer_size=1, dtype=precision)
    cfg = make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=norm)
    net = make_cvnn(6, N, seed=seed, dtype=precision.to_torch())
    pricer_cfg = expect_success(build_gbm_cvnn_pricer_config(cfg=cfg, domain_bounds=make_domain_bounds(), cvnn=net))
    return expect_success(GbmCVNNPricer.create(pricer_cfg, **kw))


def _tc(num_batches, batch_size=8, learning_rate=LEARNING_RATE):
    return expect_success(build_training_config(num_batches=num_batches, batch_size=batch_size, learning_rate=learning_rate))


def _max_param_diff(a, b) -> float:
    return max(float((p.detach() - q.detach()).abs().max()) for p, q in zip(a._cvnn.parameters(), b._cvnn.parameters()))


def _clone_model(model):
    return copy.deepcopy(model)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_deterministic_construction(precision) -> None:
    assert _max_param_diff(_pricer(precision), _pricer(precision)) == 0.0


@pytest.mark.parametrize("precision", PRECISIONS)
def test_lockstep_training(precision) -> None:
    """tests/test_gbm_trainer.py:182-195: two trainers stay bit-identical through several train() calls."""
    a, b = _pricer(precision, seed=43), _pricer(precision, seed=43)
    for batches in (2, 3, 1):
        ra, rb = expect_success(a.train(_tc(batches))), expect_success(b.train(_tc(batches)))
        assert ra.losses == rb.losses and all(math.isfinite(x) for x in ra.losses)
        assert ra.total_batches == batches and ra.final_loss == ra.losses[-1] and math.isfinite(ra.final_grad_norm)
        assert _max_param_diff(a, b) == 0.0
    sa, sb = expect_success(a.snapshot()), expect_success(b.snapshot())
    assert sa.sobol_skip == sb.sobol_skip == 48 and sa.global_step == 6
    assert sa.cfg.sim_params.skip == 48  # one normal matrix per contract priced


@pytest.mark.parametrize("precision", PRECISIONS)
def test_snapshot_cycle_deterministic(precision) -> None:
    """tests/test_gbm_trainer.py:203-223: train 3 -> snapshot -> clone; both train 2 more."""
    trainer = _pricer(precision, seed=44)
    expect_success(trainer.train(_tc(3)))
    snap = expect_success(trainer.snapshot()).model_copy(update={"cvnn": _clone_model(trainer._cvnn)})
    clone = expect_success(GbmCVNNPricer.create(snap))
    expect_success(trainer.train(_tc(2)))
    expect_success(clone.train(_tc(2)))
    assert _max_param_diff(trainer, clone) == 0.0
    done = expect_success(clone.snapshot())
    assert done.global_step == 5 and done.sobol_skip == 40 and done.cfg.sim_params.skip == 40


@pytest.mark.parametrize("precision", PRECISIONS)
def test_snapshot_restart_without_optimizer(precision) -> None:
    """tests/test_gbm_trainer.py:231-263."""
    def restarted_without_optimizer():
        t = _pricer(precision, seed=45)
        expect_success(t.train(_tc(3)))
        snap = expect_success(t.snapshot()).model_copy(update={"optimizer_state": None, "cvnn": _clone_model(t._cvnn)})
        return expect_success(GbmCVNNPricer.create(snap))

    a, b = restarted_without_optimizer(), restarted_without_optimizer()
    expect_success(a.train(_tc(2)))
    expect_success(b.train(_tc(2)))
    assert _max_param_diff(a, b) == 0.0
    assert float(expect_success(a.snapshot()).optimizer_state["state"][0]["step"]) == 2.0  # Adam restarted from zero


@pytest.mark.parametrize("precision", PRECISIONS)
def test_snapshot_optimizer_state_layout(precision) -> None:
    """tests/test_gbm_trainer.py:271-296: the snapshot carries Adam's state for every parameter, on the CPU."""
    trainer = _pricer(precision, seed=50)
    expect_success(trainer.train(_tc(4)))
    opt = expect_success(trainer.snapshot()).optimizer_state
    assert opt is not None and len(opt["param_groups"]) == 1 and len(opt["state"]) == len(list(trainer._cvnn.parameters()))
    for entry in opt["state"].values():
        assert float(entry["step"]) == 4.0 and entry["exp_avg"].device.type == "cpu" and entry["exp_avg_sq"].device.type == "cpu"
    reference_adam = torch.optim.Adam(_clone_model(trainer._cvnn).parameters(), lr=LEARNING_RATE)
    reference_adam.load_state_dict(opt)  # the layout IS torch.optim.Adam's


@pytest.mark.parametrize("precision", PRECISIONS)
def test_predict_price_smoke(precision) -> None:
    """tests/test_gbm_trainer.py:304-320."""
    trainer = _pricer(precision, seed=60)
    expect_success(trainer.train(_tc(1, batch_size=4)))
    contracts = [BlackScholes.Inputs(X0=100.0, K=100.0, T=1.0, r=0.05, d=0.02, v=0.20),
                 BlackScholes.Inputs(X0=120.0, K=110.0, T=0.5, r=0.03, d=0.01, v=0.25)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)  # an untrained network's DC bin need not be real
        results = expect_success(trainer.predict_price(contracts))
    assert len(results) == len(contracts)
    for res in results:
        for val in res.model_dump(mode="python").values():
            assert isinstance(val, float) and math.isfinite(val)
    assert expect_success(trainer.predict_price([])) == []


def test_step_logger_receives_every_step() -> None:
    """The ``logger`` callback of train() (reference :1567-1583)."""
    seen = []
    trainer = _pricer(seed=7)
    result = expect_success(trainer.train(_tc(3), logger=seen.append))
    assert [m.step for m in seen] == [1, 2, 3] and [m.loss for m in seen] == list(result.losses)
    assert all(m.lr == LEARNING_RATE and math.isfinite(m.grad_norm) and m.grad_norm > 0 and m.model is trainer._cvnn for m in seen)
    assert seen[-1].grad_norm == result.final_grad_norm


def test_training_config_and_device_validation() -> None:
    from spectralmc_b200.errors import DeviceDTypeError, DeviceNotCUDA, InvalidTrainingConfig

    for bad in (dict(num_batches=0, batch_size=8, learning_rate=0.01), dict(num_batches=1, batch_size=0, learning_rate=0.01),
                dict(num_batches=1, batch_size=8, learning_rate=1.0), dict(num_batches=1, batch_size=8, learning_rate=0.0)):
        assert isinstance(expect_failure(build_training_config(**bad)), InvalidTrainingConfig)
    sp = make_simulation_params(timesteps=1, network_size=16, batches_per_mc_run=64, mc_seed=3, dtype=Precision.float32)
    cfg = make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
    on_cpu = expect_success(build_gbm_cvnn_pricer_config(cfg=cfg, domain_bounds=make_domain_bounds(), cvnn=make_cvnn(6, 16, device="cpu")))
    assert isinstance(expect_failure(GbmCVNNPricer.create(on_cpu)), DeviceNotCUDA)
    wrong = expect_success(build_gbm_cvnn_pricer_config(cfg=cfg, domain_bounds=make_domain_bounds(), cvnn=make_cvnn(6, 16, dtype=torch.float64)))
    assert isinstance(expect_failure(GbmCVNNPricer.create(wrong)), DeviceDTypeError)


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
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        prices = expect_success(p.predict_price(inputs))
    assert len(prices) == 2 and all(math.isfinite(x.put_price) for x in prices)


def test_training_reduces_the_loss() -> None:
    p = _pricer(Precision.float32, N=16, B=2**10)
    losses = expect_success(p.train(_tc(30, batch_size=32))).losses
    assert np.mean(losses[-5:]) < np.mean(losses[:5])


def test_full_stack_normalize_train_snapshot_reload_predict() -> None:
    """The reference's end-to-end test without its S3 store (tests/test_e2e/test_full_stack_cvnn_pricer.py:41-125):
    16 steps x 128 x 4 batches with the default NORMALIZE, train 4 x 8, snapshot, rebuild, predict."""
    trainer = _pricer(Precision.float32, seed=11, N=128, B=4, T=16, norm=ForwardNormalization.NORMALIZE)
    result = expect_success(trainer.train(_tc(4, batch_size=8)))
    assert result.total_batches == 4 and all(math.isfinite(x) for x in result.losses)
    snap = result.updated_config
    assert snap.global_step == 4 and snap.sobol_skip == 32 and snap.cfg.sim_params.skip == 32
    assert snap.cfg.normalization is ForwardNormalization.NORMALIZE
    reloaded = expect_success(GbmCVNNPricer.create(snap.model_copy(update={"cvnn": _clone_model(trainer._cvnn)})))
    contracts = [BlackScholes.Inputs(X0=100.0, K=100.0, T=1.0, r=0.05, d=0.02, v=0.20)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        a = expect_success(trainer.predict_price(contracts))[0]
        b = expect_success(reloaded.predict_price(contracts))[0]
    assert a == b and math.isfinite(a.put_price)
    # and the reloaded trainer continues exactly where the original would
    expect_success(trainer.train(_tc(1, batch_size=8)))
    expect_success(reloaded.train(_tc(1, batch_size=8)))
    assert _max_param_diff(trainer, reloaded) == 0.0


def test_wide_steps_take_the_graphed_torch_route_on_the_same_buffers(monkeypatch) -> None:
    """fused_step=None routes per batch size: below FUSED_STEP_MAX_MACS the C-ABI step, above it torch autograd +
    smc_adam_step (profiles/r2_cvnn_widths.md) — on the SAME flat parameter / Adam buffers, so a run may mix batch
    sizes.  With the threshold forced to zero every step takes the torch route and must track the C-ABI route."""
    import spectralmc_b200.gbm_trainer as gt

    a = _pricer(Precision.float32, seed=5, N=16, B=2**8)
    b = _pricer(Precision.float32, seed=5, N=16, B=2**8)
    monkeypatch.setattr(gt, "FUSED_STEP_MAX_MACS", 0)
    la = expect_success(a.train(_tc(6, batch_size=16))).losses
    assert a._torch_graphs and not a._graphs  # every step went through the torch graph
    monkeypatch.setattr(gt, "FUSED_STEP_MAX_MACS", 10**18)
    lb = expect_success(b.train(_tc(6, batch_size=16))).losses
    assert b._graphs and not b._torch_graphs
    assert np.allclose(la, lb, rtol=2e-4)
    assert _max_param_diff(a, b) <= 2e-4
    # mixing the routes on one trainer: the optimiser state lives in one place
    monkeypatch.setattr(gt, "FUSED_STEP_MAX_MACS", 0)
    expect_success(b.train(_tc(2, batch_size=16)))
    expect_success(a.train(_tc(2, batch_size=16)))
    assert _max_param_diff(a, b) <= 5e-4
    snap = expect_success(a.snapshot())
    assert snap.global_step == 8 and snap.optimizer_state is not None
