"""Host-side logic of the drop-in API (no device work): config validation, Result ADTs, the
Sobol sampler — mirroring the reference's CPU-only assertions
(tests/test_async_normals.py:153-170, tests/test_sobol_sampler.py:107-187)."""

from __future__ import annotations

import numpy as np
import pytest

from spectralmc_b200.async_normals import BufferConfig, ConcurrentNormGeneratorConfig
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.errors import GPUMemoryLimitExceeded, InvalidSimulationParams
from spectralmc_b200.gbm import BlackScholes, build_simulation_params
from spectralmc_b200.numerical import Precision
from spectralmc_b200.result import Failure, Success, collect_results, fold_results
from spectralmc_b200.sobol_sampler import SobolConfig, SobolSampler, build_bound_spec, build_domain_bounds
from tests.conftest import load_golden
from tests.helpers import expect_failure, expect_success, make_black_scholes_config, make_domain_bounds, make_simulation_params


def test_norm_config_validation() -> None:
    bad_rows = ConcurrentNormGeneratorConfig.create(rows=0, cols=2, seed=1, dtype=Precision.float32, skips=0)
    assert "InvalidShape" in str(expect_failure(bad_rows))
    bad_seed = ConcurrentNormGeneratorConfig.create(rows=2, cols=2, seed=0, dtype=Precision.float32, skips=0)
    assert "SeedOutOfRange" in str(expect_failure(bad_seed))
    bad_skip = ConcurrentNormGeneratorConfig.create(rows=2, cols=2, seed=1, dtype=Precision.float32, skips=-1)
    assert "SeedOutOfRange" in str(expect_failure(bad_skip))
    assert isinstance(BufferConfig.create(0, 2, 2), Failure)
    assert isinstance(BufferConfig.create(5, 2, 2), Failure)  # more buffers than elements
    assert expect_success(BufferConfig.create(2, 2, 2)).size == 2


def test_simulation_params_schema() -> None:
    sp = make_simulation_params(timesteps=12, network_size=16, batches_per_mc_run=64, dtype=Precision.float64)
    assert sp.total_paths() == 1024 and sp.total_blocks() == 4
    with pytest.raises(Exception):
        sp.timesteps = 3  # frozen
    bad = build_simulation_params(timesteps=0, network_size=16, batches_per_mc_run=1, threads_per_block=256,
                                  mc_seed=1, buffer_size=1, dtype=Precision.float32)
    assert isinstance(expect_failure(bad), InvalidSimulationParams)
    bad_tpb = build_simulation_params(timesteps=1, network_size=16, batches_per_mc_run=1, threads_per_block=48,
                                      mc_seed=1, buffer_size=1, dtype=Precision.float32)
    assert isinstance(expect_failure(bad_tpb), InvalidSimulationParams)
    big = build_simulation_params(timesteps=1, network_size=1024, batches_per_mc_run=1_000_000, threads_per_block=256,
                                  mc_seed=1, buffer_size=1, dtype=Precision.float32)
    err = expect_failure(big)
    assert isinstance(err, GPUMemoryLimitExceeded) and err.max_paths == 1_000_000_000
    big64 = build_simulation_params(timesteps=1, network_size=1024, batches_per_mc_run=600_000, threads_per_block=256,
                                    mc_seed=1, buffer_size=1, dtype=Precision.float64)
    assert expect_failure(big64).max_paths == 500_000_000


def test_config_defaults_and_inputs_validation() -> None:
    cfg = make_black_scholes_config()
    assert cfg.path_scheme is PathScheme.LOG_EULER and cfg.normalization is ForwardNormalization.NORMALIZE
    assert list(BlackScholes.Inputs.model_fields) == ["X0", "K", "T", "r", "d", "v"]  # CVNN input order
    BlackScholes.Inputs(X0=1.0, K=1.0, T=0.0, r=0.0, d=0.0, v=0.0)  # T = 0, v = 0 are legal
    for bad in (dict(X0=0.0), dict(K=-1.0), dict(T=-0.1), dict(v=-0.1)):
        with pytest.raises(Exception):
            BlackScholes.Inputs(**{**dict(X0=1.0, K=1.0, T=1.0, r=0.0, d=0.0, v=0.1), **bad})


def test_result_helpers() -> None:
    assert collect_results([Success(1), Success(2)]) == Success([1, 2])
    assert collect_results([Success(1), Failure("a"), Failure("b")]) == Failure("a")
    assert fold_results([1, 2, 3], lambda acc, x: Success(acc + x), 0) == Success(6)
    assert fold_results([1, -2, 3], lambda acc, x: Success(acc + x) if x > 0 else Failure("neg"), 0) == Failure("neg")
    assert Success(2).map(lambda v: v + 1).unwrap() == 3 and Failure("e").unwrap_or(7) == 7
    with pytest.raises(RuntimeError):
        Failure("e").unwrap()


# ---- Sobol -----------------------------------------------------------------------------------
def _sampler(seed: int, skip: int = 0) -> SobolSampler:
    return expect_success(SobolSampler.create(BlackScholes.Inputs, make_domain_bounds(), config=SobolConfig(seed=seed, skip=skip)))


def test_sobol_contracts_bit_exact_with_reference() -> None:
    """Golden rows came from the reference's own SobolSampler (make_golden.py)."""
    g = load_golden("sobol_contracts")
    for key, ref in g.items():
        seed, skip, n = (int(t[len(p):]) for t, p in zip(key.split("_"), ("seed", "skip", "n")))
        pts = expect_success(_sampler(seed, skip).sample(n))
        got = np.array([[p.X0, p.K, p.T, p.r, p.d, p.v] for p in pts])
        assert np.array_equal(got, ref), key
        assert np.array_equal(expect_success(_sampler(seed, skip).sample_array(n)), ref), key


def test_sobol_oracle_matches_golden() -> None:
    from oracle.sobol import sobol_contracts

    g = load_golden("sobol_contracts")
    assert np.array_equal(sobol_contracts(16, 42), g["seed42_skip0_n16"])
    assert np.array_equal(sobol_contracts(8, 42, skip=5), g["seed42_skip5_n8"])


def test_sobol_skip_repro() -> None:
    """skip = n is the same as burning n points (reference tests/test_sobol_sampler.py:131-150)."""
    a = _sampler(42)
    expect_success(a.sample(5))
    rest = expect_success(a.sample(8))
    skipped = expect_success(_sampler(42, skip=5).sample(8))
    assert [tuple(p.model_dump().values()) for p in rest] == [tuple(p.model_dump().values()) for p in skipped]


def test_sobol_bounds_and_validation() -> None:
    pts = expect_success(_sampler(3).sample_array(256))
    b = make_domain_bounds()
    for i, f in enumerate(b.fields):
        assert np.all(pts[:, i] >= b[f].lower - 1e-12) and np.all(pts[:, i] <= b[f].upper + 1e-12)
    assert "NegativeSamples" in str(expect_failure(_sampler(3).sample(-1)))
    assert expect_success(_sampler(3).sample(0)) == []
    assert "BoundSpecInvalid" in str(expect_failure(build_bound_spec(1.0, 1.0)))
    spec = {"X0": expect_success(build_bound_spec(0.0, 1.0))}
    assert "DimensionMismatch" in str(expect_failure(build_domain_bounds(BlackScholes.Inputs, spec)))
    # bounds that violate the model (X0 must be > 0) surface as SamplerValidationFailed
    neg = expect_success(SobolSampler.create(BlackScholes.Inputs, make_domain_bounds(x0=(-5.0, -1.0)), config=SobolConfig(seed=1)))
    assert "SamplerValidationFailed" in str(expect_failure(neg.sample(4)))
    assert "SamplerValidationFailed" in str(expect_failure(neg.sample_array(4)))


def test_training_config_validation_mirrors_the_reference() -> None:
    """gbm_trainer.py:261-298 — positive counts, learning rate in (0, 1); no device needed."""
    from spectralmc_b200.errors import InvalidTrainingConfig
    from spectralmc_b200.gbm_trainer import TrainingConfig, build_training_config

    ok = build_training_config(num_batches=3, batch_size=8, learning_rate=1e-2)
    assert isinstance(ok, Success) and ok.value == TrainingConfig(num_batches=3, batch_size=8, learning_rate=1e-2)
    for bad, word in ((dict(num_batches=0, batch_size=8, learning_rate=0.01), "num_batches"),
                      (dict(num_batches=1, batch_size=-1, learning_rate=0.01), "batch_size"),
                      (dict(num_batches=1, batch_size=8, learning_rate=1.0), "learning_rate"),
                      (dict(num_batches=1, batch_size=8, learning_rate=0.0), "learning_rate")):
        got = build_training_config(**bad)
        assert isinstance(got, Failure) and isinstance(got.error, InvalidTrainingConfig) and word in got.error.message


def test_pricer_create_rejects_a_cpu_model_without_touching_the_device() -> None:
    """gbm_trainer.py:600-640 — DeviceNotCUDA / DeviceDTypeError are decided from the module's tensors."""
    import torch

    from spectralmc_b200.cvnn import make_cvnn
    from spectralmc_b200.errors import DeviceDTypeError, DeviceNotCUDA
    from spectralmc_b200.gbm_trainer import GbmCVNNPricer, build_gbm_cvnn_pricer_config
    from tests.helpers import make_black_scholes_config, make_domain_bounds, make_simulation_params

    sp = make_simulation_params(timesteps=1, network_size=16, batches_per_mc_run=64, mc_seed=3)
    cfg = make_black_scholes_config(sim_params=sp)
    pc = build_gbm_cvnn_pricer_config(cfg=cfg, domain_bounds=make_domain_bounds(), cvnn=make_cvnn(6, 16, device="cpu"))
    assert isinstance(pc, Success) and pc.value.global_step == 0 and pc.value.optimizer_state is None
    got = GbmCVNNPricer.create(pc.value)
    assert isinstance(got, Failure) and isinstance(got.error, DeviceNotCUDA)
    mixed = make_cvnn(6, 16, device="cpu")
    mixed.layers[1].double()
    got = GbmCVNNPricer.create(pc.value.model_copy(update={"cvnn": mixed}))
    assert isinstance(got, Failure) and isinstance(got.error, DeviceDTypeError)
    assert isinstance(build_gbm_cvnn_pricer_config(cfg=cfg, domain_bounds=make_domain_bounds(), cvnn=mixed, bogus=1), Failure)


def test_registry_mirrors_the_reference_names() -> None:
    """effects/registry.py:166-254,502-553 — tensor / kernel registration under the reference's class names."""
    import torch

    from spectralmc_b200.interpreter import MonteCarloInterpreter, MonteCarloOperators, SharedRegistry, TensorRegistry

    assert SharedRegistry is TensorRegistry and MonteCarloInterpreter is MonteCarloOperators
    reg = SharedRegistry()
    assert reg.has_kernel("SimulateBlackScholes") and not reg.has_tensor("z")
    assert isinstance(reg.register_tensor("z", torch.zeros(2, 3)), Success) and reg.has_tensor("z")
    assert isinstance(reg.get_torch_tensor("z"), Success) and isinstance(reg.get_tensor("missing"), Failure)
    assert isinstance(reg.register_tensor("", torch.zeros(1)), Failure)
    assert isinstance(reg.register_kernel("mine", print), Success) and reg.get_kernel("mine").value is print
    assert isinstance(reg.register_kernel("", print), Failure) and isinstance(reg.get_kernel("other"), Failure)
    reg.clear_tensors()
    assert not reg.has_tensor("z") and reg.has_kernel("mine")


def test_import_level_drop_in_alias() -> None:
    """``from spectralmc.gbm import ...`` resolves to this package after ``install_as_spectralmc()``; a real
    installation would not be shadowed silently; out-of-scope modules stay unresolved."""
    import importlib
    import sys

    import pytest

    import spectralmc_b200.compat as compat

    assert "spectralmc" not in sys.modules
    names = compat.install_as_spectralmc()
    try:
        assert "spectralmc.gbm" in names and "spectralmc.effects.montecarlo" in names
        gbm = importlib.import_module("spectralmc.gbm")
        from spectralmc.async_normals import ConcurrentNormGenerator  # noqa: F401
        from spectralmc.effects.montecarlo import ForwardNormalization as FN
        from spectralmc.models.numerical import Precision as P
        from spectralmc.sobol_sampler import SobolSampler  # noqa: F401

        import spectralmc_b200

        assert gbm.BlackScholes is spectralmc_b200.BlackScholes and gbm.SimulateBlackScholes is spectralmc_b200.SimulateBlackScholes
        assert FN is spectralmc_b200.ForwardNormalization and P is spectralmc_b200.Precision
        with pytest.raises(ImportError):
            importlib.import_module("spectralmc.storage")  # out of scope: loudly absent
        compat.install_as_spectralmc()  # idempotent over its own aliases
    finally:
        compat.uninstall()
    assert "spectralmc" not in sys.modules and "spectralmc.gbm" not in sys.modules
