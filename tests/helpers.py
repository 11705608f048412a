"""Factories mirroring the reference's tests/helpers (factories.py:108-249, constants.py:17-34)."""

from __future__ import annotations

import numpy as np

from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import (
    BlackScholes,
    BlackScholesConfig,
    SimulationParams,
    ThreadsPerBlock,
    build_black_scholes_config,
    build_simulation_params,
)
from spectralmc_b200.numerical import Precision
from spectralmc_b200.result import Failure, Success
from spectralmc_b200.sobol_sampler import DomainBounds, build_bound_spec, build_domain_bounds


def expect_success(result):
    assert isinstance(result, Success), f"expected Success, got {result}"
    return result.value


def expect_failure(result):
    assert isinstance(result, Failure), f"expected Failure, got {result}"
    return result.error


def make_domain_bounds(
    *,
    x0=(0.001, 10_000.0),
    k=(0.001, 20_000.0),
    t=(0.0, 10.0),
    r=(-0.20, 0.20),
    d=(-0.20, 0.20),
    v=(0.0, 2.0),
) -> DomainBounds:
    spec = {n: expect_success(build_bound_spec(*b)) for n, b in dict(X0=x0, K=k, T=t, r=r, d=d, v=v).items()}
    return expect_success(build_domain_bounds(BlackScholes.Inputs, spec))


def make_simulation_params(
    timesteps: int = 100,
    network_size: int = 1024,
    batches_per_mc_run: int = 8,
    threads_per_block: ThreadsPerBlock = 256,
    mc_seed: int = 42,
    buffer_size: int = 1,
    skip: int = 0,
    dtype: Precision = Precision.float32,
) -> SimulationParams:
    return expect_success(
        build_simulation_params(
            timesteps=timesteps,
            network_size=network_size,
            batches_per_mc_run=batches_per_mc_run,
            threads_per_block=threads_per_block,
            mc_seed=mc_seed,
            buffer_size=buffer_size,
            skip=skip,
            dtype=dtype,
        )
    )


def make_black_scholes_config(
    sim_params: SimulationParams | None = None,
    path_scheme: PathScheme = PathScheme.LOG_EULER,
    normalization: ForwardNormalization = ForwardNormalization.NORMALIZE,
) -> BlackScholesConfig:
    return expect_success(
        build_black_scholes_config(
            sim_params=sim_params or make_simulation_params(), path_scheme=path_scheme, normalization=normalization
        )
    )


def rel_max(a, b) -> float:
    """Norm-wise relative error max|a-b| / max|b| (SURVEY.md §8d parity metric)."""
    a, b = np.asarray(a), np.asarray(b)
    denom = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b))) / (denom if denom > 0 else 1.0)


def rel_elem(a, b, floor: float = 0.0) -> float:
    """Element-wise relative error max |a-b| / max(|b|, floor)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), max(floor, 1e-300))))
