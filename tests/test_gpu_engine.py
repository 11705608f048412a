"""The drop-in engine API (spectralmc_b200.gbm.BlackScholes) — the reference's engine tests
(tests/test_gbm.py:142-156 snapshot determinism; :94-99 price_to_host) plus parity of the
materialised engine path against the oracle on the same Philox normals."""

from __future__ import annotations

import math

import numpy as np
import pytest
import torch

from oracle import gbm as ogbm
from oracle import philox
from spectralmc_b200.analytic import bs_price_quantlib
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import BlackScholes
from spectralmc_b200.numerical import Precision
from spectralmc_b200.result import Success
from tests.helpers import expect_success, make_black_scholes_config, make_simulation_params, rel_elem, rel_max

pytestmark = pytest.mark.gpu

INP = BlackScholes.Inputs(X0=100, K=100, T=1.0, r=0.05, d=0.0, v=0.2)


def _make_engine(precision, *, skip=0, T=1, N=256, B=2**12, norm=ForwardNormalization.RAW, scheme=PathScheme.LOG_EULER, buffer=1):
    sp = make_simulation_params(timesteps=T, network_size=N, batches_per_mc_run=B, threads_per_block=256, mc_seed=7,
                                buffer_size=buffer, skip=skip, dtype=precision)
    return BlackScholes(make_black_scholes_config(sim_params=sp, path_scheme=scheme, normalization=norm))


def _collect(engine, inp, n):
    return [expect_success(engine.price_to_host(inp)) for _ in range(n)]


@pytest.mark.parametrize("precision", [Precision.float64, Precision.float32])
def test_snapshot_determinism(precision) -> None:
    """snapshot -> restore reproduces the next prices (reference rel_tol 1e-6; here bit-equal)."""
    engine = _make_engine(precision)
    _ = _collect(engine, INP, 8)
    snap = expect_success(engine.snapshot())
    assert snap.sim_params.skip == 8
    expected = _collect(engine, INP, 8)
    reproduced = _collect(BlackScholes(snap), INP, 8)
    for e, r in zip(expected, reproduced, strict=True):
        assert e.put_price == r.put_price and e.call_price == r.call_price
    # a different buffer size does not change the stream
    other = _collect(_make_engine(precision, skip=8, buffer=3), INP, 8)
    assert [o.put_price for o in other] == [e.put_price for e in expected]


@pytest.mark.parametrize("precision", [Precision.float64, Precision.float32])
@pytest.mark.parametrize("norm", [ForwardNormalization.RAW, ForwardNormalization.NORMALIZE])
@pytest.mark.parametrize("scheme", [PathScheme.LOG_EULER, PathScheme.SIMPLE_EULER])
def test_materialised_engine_matches_oracle(precision, norm, scheme) -> None:
    """_simulate / price / get_host_price with full SimResults, vs the oracle fed the same matrix."""
    T, N, B = 12, 16, 64
    engine = _make_engine(precision, skip=4, T=T, N=N, B=B, norm=norm, scheme=scheme)
    sr = expect_success(engine._simulate(INP))
    pr = expect_success(engine.price(inputs=INP, sr_result=Success(sr)))
    host = engine.get_host_price(pr)
    z = philox.normals_matrix(T, N * B, precision.to_numpy(), 7, 4)
    c = ogbm.Contract(100, 100, 1.0, 0.05, 0.0, 0.2)
    osr = ogbm.simulate(c, z, scheme=scheme.value, normalization=norm.value)
    opr = ogbm.price(c, osr)
    tol = 1e-12 if precision is Precision.float64 else 1e-5
    assert sr.sims.shape == (T, N * B) and sr.sims.dtype == precision.to_torch()
    for name in ("times", "forwards", "df"):
        assert rel_elem(getattr(sr, name).cpu().numpy(), getattr(osr, name)) <= (0 if precision is Precision.float32 else 1e-15), name
    assert rel_elem(sr.sims.cpu().numpy(), osr.sims) <= tol
    assert rel_max(pr.put_price.cpu().numpy(), opr.put_price) <= tol
    assert rel_max(pr.call_price.cpu().numpy(), opr.call_price) <= tol
    assert pr.underlying.data_ptr() == sr.sims[-1].data_ptr()  # a view, as in the reference (gbm.py:472)
    ohost = ogbm.host_price(opr)
    for k, v in ohost.items():
        assert abs(getattr(host, k) - v) <= tol * max(abs(v), 1.0), k
    assert expect_success(engine.snapshot()).sim_params.skip == 5


@pytest.mark.parametrize("precision", [Precision.float64, Precision.float32])
def test_fused_and_materialised_engine_paths_share_one_stream(precision) -> None:
    """cf_targets (fused) and price (materialised) consume the same matrices in the same order."""
    T, N, B = 8, 32, 128
    a = _make_engine(precision, T=T, N=N, B=B)
    b = _make_engine(precision, T=T, N=N, B=B)
    fused = expect_success(a.cf_targets([INP, INP, INP]))
    mats = []
    for _ in range(3):
        pr = expect_success(b.price(inputs=INP))
        from spectralmc_b200 import _cabi

        mats.append(_cabi.cf_fft_mean(pr.put_price.view(B, N)))
    tol = 1e-12 if precision is Precision.float64 else 1e-5
    for i in range(3):
        assert rel_max(fused[i].cpu().numpy(), mats[i].cpu().numpy()) <= tol
    # interleaving: a materialised call after a fused batch picks up at matrix 3
    pr = expect_success(a.price(inputs=INP))
    nxt = expect_success(b.price(inputs=INP))
    assert torch.equal(pr.put_price, nxt.put_price)
    assert expect_success(a.snapshot()).sim_params.skip == 4


def test_price_to_host_against_black76() -> None:
    engine = _make_engine(Precision.float64, T=1, N=256, B=2**13)
    vals = np.array([h.put_price for h in _collect(engine, INP, 16)])
    ref = bs_price_quantlib(INP).put_price
    assert abs(vals.mean() - ref) <= 4 * vals.std(ddof=1) / math.sqrt(16)
    h = _collect(engine, INP, 1)[0]
    assert h.put_convexity == pytest.approx(h.put_price - h.put_price_intrinsic)
    assert h.call_price_intrinsic == pytest.approx(math.exp(-0.05) * (100 * math.exp(0.05) - 100), rel=1e-12)


def test_dlpack_handoff() -> None:
    """Targets cross into torch/any DLPack consumer zero-copy (reference gbm_trainer.py:1556)."""
    engine = _make_engine(Precision.float32, T=4, N=16, B=64)
    cf = expect_success(engine.cf_targets([INP, INP]))
    again = torch.from_dlpack(cf)
    assert again.data_ptr() == cf.data_ptr() and again.dtype == torch.complex64 and again.shape == (2, 16)
