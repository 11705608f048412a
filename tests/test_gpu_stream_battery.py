"""Statistical battery for the float32 normal stream (21-bit radius / angle uniforms, Philox4x32-10), at sample
sizes no HBM matrix reaches: the stream is this package's own specification (oracle/philox.py — the reference's
CuPy XORWOW bits are third-party and unpinned), so trust has to be earned by measurement.

* the device audit kernels are first pinned on the oracle (exact field histograms at a small size);
* chi-square of all 2^21 cells of the radius and of the angle fields over 2^33 draws each;
* tail mass of the normals beyond 4 / 5 / 5.5 / 6 sigma (refined entries included) against the tail mass the
  SPECIFICATION implies (computed exactly from the 21-bit grid), which in turn is compared with the normal law;
* moments at 1.7e10 draws;
* serial correlation along a path at lags 1..6 (within and across the 6-row blocks) and across adjacent columns;
* a pricing check where the tail matters: deep out-of-the-money / in-the-money puts at 2^31 paths, float32 stream
  vs float64 stream vs Black-76, within 4 standard errors.
* the float64 stream (two pairs per block, 43-bit radius + 21-bit angle field per pair) audited from its output at 1.3e8 normals.
Budget: well under a minute on a B200.
"""

from __future__ import annotations

import math

import numpy as np
import pytest
import torch
from scipy import integrate, stats

from oracle import philox
from oracle.black76 import black76
from spectralmc_b200 import _cabi

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda", 0)


@pytest.mark.parametrize("stream_version", [0, 1])
def test_audit_kernels_follow_the_specification(stream_version) -> None:
    """Small enough for the oracle: every field count and the power sums must match it exactly / to rounding."""
    cols, groups, seed, k = 64, 40, 42, 3
    got = _cabi.diag_stream_fields(seed, k, cols * groups, cols, DEV, stream_version=stream_version)
    j = np.arange(cols, dtype=np.uint32)[None, :]
    q = np.arange(groups, dtype=np.uint32)[:, None]
    x = philox.philox4x32_10((j, q, k, 0), (seed, 0), philox.STREAM_ROUNDS[stream_version])
    radius, angle = philox.f32_fields(*x)
    want_r = np.bincount(np.concatenate([r.ravel() for r in radius]).astype(np.int64), minlength=1 << 21)
    want_a = np.bincount(np.concatenate([a.ravel() for a in angle]).astype(np.int64), minlength=1 << 21)
    assert np.array_equal(got["radius_hist"].cpu().numpy(), want_r)
    assert np.array_equal(got["angle_hist"].cpu().numpy(), want_a)
    z = philox.normals_matrix(6 * groups, cols, np.float32, seed, k, stream_version=stream_version).astype(np.float64)
    sums = got["power_sums"].cpu().numpy()
    for p in range(4):
        assert abs(sums[p] - np.sum(z ** (p + 1))) <= 2e-4 * np.sum(np.abs(z) ** (p + 1))
    assert got["tails"].cpu().tolist() == [int(np.sum(np.abs(z) > t)) for t in (4.0, 5.0, 5.5, 6.0)]
    lags = _cabi.diag_stream_lags(seed, k, cols, 6 * groups, DEV, stream_version=stream_version).cpu().numpy()
    for lag in range(1, 7):
        assert abs(lags[lag - 1] - np.sum(z[lag:] * z[:-lag])) <= 1e-3 * (1 + abs(np.sum(z[lag:] * z[:-lag])))
    pairs = np.arange(cols - 1)
    pairs = pairs[pairs % 32 != 31]
    assert abs(lags[6] - np.sum(z[:, pairs] * z[:, pairs + 1])) <= 1e-3 * (1 + abs(np.sum(z[:, pairs] * z[:, pairs + 1])))


def _spec_tail_probability(t: float) -> float:
    """P(|z| > t) the float32 stream's SPECIFICATION implies: radius uniform on the 21-bit midpoint grid (field 0
    refined to a uniform on (0, 2^-21)), angle uniform, |z| = r |cos theta|  =>  P = E[(2 / pi) acos(t / r); r > t]."""
    F = np.arange(1, 1 << 21, dtype=np.float64)
    r = np.sqrt(-2.0 * np.log((F + 0.5) * 2.0**-21))
    grid = np.sum(np.where(r > t, np.arccos(np.minimum(1.0, t / np.maximum(r, 1e-300))), 0.0)) * (2.0 / math.pi) * 2.0**-21

    def integrand(u):
        rr = math.sqrt(-2.0 * math.log(u))
        return (2.0 / math.pi) * math.acos(min(1.0, t / rr)) if rr > t else 0.0

    refined, _ = integrate.quad(integrand, 0.0, 2.0**-21, limit=200, points=[math.exp(-t * t / 2)] if math.exp(-t * t / 2) < 2.0**-21 else None)
    return float(grid + refined)


# every statistic below runs on the default stream (Philox4x32-10) AND on the opt-in one (Philox4x32-7)
@pytest.fixture(scope="module", params=[0, 1], ids=["philox10", "philox7"])
def big_audit(request):
    n_blocks = (1 << 33) // 3 + 1  # 2^33 draws of each field type, 1.7e10 normals
    out = _cabi.diag_stream_fields(20260318, 5, n_blocks, 1 << 22, DEV, stream_version=request.param)
    torch.cuda.synchronize()
    return n_blocks, {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("which", ["radius_hist", "angle_hist"])
def test_chi_square_of_all_field_cells(big_audit, which) -> None:
    n_blocks, out = big_audit
    counts = out[which].astype(np.float64)
    total = counts.sum()
    assert total == 3 * n_blocks
    expect = total / (1 << 21)
    chi2 = float(np.sum((counts - expect) ** 2) / expect)
    dof = (1 << 21) - 1
    z = (chi2 - dof) / math.sqrt(2 * dof)
    assert abs(z) < 5.0, (which, chi2, z)
    assert counts.min() > 0.9 * expect and counts.max() < 1.1 * expect  # ~4096 per cell, sd 64


def test_tail_mass_and_moments(big_audit) -> None:
    n_blocks, out = big_audit
    n = 6.0 * n_blocks
    normal_law = [2.0 * stats.norm.sf(t) for t in (4.0, 5.0, 5.5, 6.0)]
    for t, got, ideal in zip((4.0, 5.0, 5.5, 6.0), out["tails"], normal_law):
        spec = _spec_tail_probability(t)
        sd = math.sqrt(n * spec)
        assert abs(got - n * spec) < 5.0 * sd + 1.0, (t, int(got), n * spec)
        # and the specification itself carries the normal law's tail mass (the grid is refined where it would matter)
        assert abs(spec / ideal - 1.0) < (0.01 if t < 5.4 else 1e-3), (t, spec, ideal)  # 5 sigma: +0.3 % from the 21-bit grid
    s1, s2, s3, s4 = (float(v) / n for v in out["power_sums"])
    assert abs(s1) < 5 / math.sqrt(n) + 1e-6
    assert abs(s2 - 1.0) < 5 * math.sqrt(2 / n) + 2e-6  # + the MUFU evaluation error of sqrt / lg2 / sin / cos
    assert abs(s3) < 5 * math.sqrt(15 / n) + 1e-5
    assert abs(s4 - 3.0) < 5 * math.sqrt(96 / n) + 2e-5


@pytest.mark.parametrize("stream_version", [0, 1])
def test_no_serial_or_cross_column_correlation(stream_version) -> None:
    cols, rows = 1 << 17, 1536
    sums = _cabi.diag_stream_lags(99, 1, cols, rows, DEV, stream_version=stream_version).cpu().numpy()
    for lag in range(1, 7):
        count = cols * (rows - lag)
        assert abs(sums[lag - 1] / count) < 5 / math.sqrt(count), (lag, sums[lag - 1] / count)
    count = (cols // 32 * 31) * rows
    assert abs(sums[6] / count) < 5 / math.sqrt(count)


def _put_moments(F: float, K: float, sigma: float, df: float) -> tuple[float, float]:
    """Mean and standard deviation of df * max(K - S, 0) for lognormal S with forward F and total volatility sigma."""
    mu = math.log(F) - 0.5 * sigma * sigma

    def m(power):
        f = lambda s: (K - s) ** power * stats.lognorm.pdf(s, sigma, scale=math.exp(mu))  # noqa: E731
        return integrate.quad(f, 0.0, K, limit=400, points=[min(K, math.exp(mu - 6 * sigma)), min(K, math.exp(mu))])[0]

    m1, m2 = m(1), m(2)
    return df * m1, df * math.sqrt(max(m2 - m1 * m1, 0.0))


@pytest.mark.parametrize("stream_version", [0, 1])
@pytest.mark.parametrize("strike_ratio", [0.5, 2.0])
def test_deep_out_of_the_money_prices(strike_ratio, stream_version) -> None:
    """2^31 paths, v = 0.2, T = 1: the put with K / X0 = 0.5 lives entirely in the left tail of the terminal
    distribution (3.6 sigma and beyond), the one with K / X0 = 2 in the bulk; both streams must price both within
    4 standard errors of Black-76."""
    X0, K, Tm, r, d, v = 100.0, 100.0 * strike_ratio, 1.0, 0.05, 0.0, 0.2
    N, B, T = 128, 1 << 24, 6
    contracts = torch.tensor([[X0, K, Tm, r, d, v]], dtype=torch.float64, device=DEV)
    F, df = X0 * math.exp((r - d) * Tm), math.exp(-r * Tm)
    analytic = black76(X0, K, Tm, r, d, v)["put_price"]
    mean, sd = _put_moments(F, K, v * math.sqrt(Tm), df)
    assert abs(mean - analytic) <= 1e-6 * max(analytic, 1e-12) + 1e-12
    se = sd / math.sqrt(N * B)
    prices = {}
    for dtype in (torch.float32, torch.float64):
        args = _cabi.make_fused_args(contracts, 1, T, N, B, dtype, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, 31, 0, stream_version=stream_version)
        cf = _cabi.cf_fused(args, DEV, dtype)
        prices[dtype] = float(cf[0, 0].real) / N  # DC bin = N * mean put price
        # float32 path arithmetic adds a relative ~1e-7 bias to every payoff: far below one standard error here
        assert abs(prices[dtype] - analytic) < 4.0 * se + 2e-7 * analytic, (dtype, prices[dtype], analytic, se)
    assert abs(prices[torch.float32] - prices[torch.float64]) < 6.0 * se + 2e-7 * analytic


@pytest.mark.parametrize("stream_version", [0, 1])
def test_float64_stream_at_scale(stream_version) -> None:
    """The float64 stream (two pairs per block: 43-bit radius and 21-bit angle field per pair) audited from its OUTPUT,
    1.3e8 normals: each pair is taken apart again — u1 = exp(-r^2 / 2) and u2 = atan2(z_odd, z_even) / 2 pi + 1/2 must
    be uniform (chi-square over 4096 cells each) and independent of each other — plus moments, tail counts against the
    normal law and correlations along a path (within a pair, across the two pairs of a block, across blocks)."""
    rows, cols = 64, 1 << 21
    z = torch.empty((rows, cols), dtype=torch.float64, device=DEV)
    _cabi.philox_normals(z, 20260318, 9, stream_version=stream_version)
    n = z.numel()
    for p, (mean, var) in enumerate([(0.0, 1.0), (1.0, 2.0), (0.0, 15.0), (3.0, 96.0)], start=1):
        m = float((z**p).mean())
        assert abs(m - mean) < 5 * math.sqrt(var / n), (p, m)
    a = z.abs()
    for t in (3.0, 4.0, 5.0):
        expect = n * 2.0 * stats.norm.sf(t)
        got = int((a > t).sum())
        assert abs(got - expect) < 5 * math.sqrt(expect) + 1, (t, got, expect)
    even, odd = z[0::2], z[1::2]  # rows 2p, 2p + 1 of a column are one Box-Muller pair
    u1 = torch.exp(-0.5 * (even * even + odd * odd))
    u2 = torch.atan2(odd, even) / (2 * math.pi) + 0.5
    cells = 4096
    pairs = u1.numel()
    for name, u in (("radius", u1), ("angle", u2)):
        counts = torch.bincount((u * cells).long().clamp_(0, cells - 1).ravel(), minlength=cells).double()
        chi2 = float(((counts - pairs / cells) ** 2).sum() / (pairs / cells))
        assert abs(chi2 - (cells - 1)) < 5 * math.sqrt(2 * (cells - 1)), (name, chi2)
    joint = torch.bincount(((u1 * 64).long().clamp_(0, 63) * 64 + (u2 * 64).long().clamp_(0, 63)).ravel(), minlength=4096).double()
    chi2 = float(((joint - pairs / 4096) ** 2).sum() / (pairs / 4096))
    assert abs(chi2 - 4095) < 5 * math.sqrt(2 * 4095), ("joint", chi2)
    for lag in (1, 2, 3, 4, 5):  # 1: within a pair and across pairs, 2-3: across the pairs of a block, 4-5: across blocks
        c = float((z[lag:] * z[:-lag]).mean())
        assert abs(c) < 5 / math.sqrt((rows - lag) * cols), (lag, c)
    c = float((z[:, 1:] * z[:, :-1]).mean())
    assert abs(c) < 5 / math.sqrt(rows * (cols - 1))
