"""K1 parity: smc_philox_normals (through the C ABI) against the stream specification in
oracle/philox.py, plus the reference's own generator tests (tests/test_async_normals.py:68-150)
run against the drop-in ``spectralmc_b200.async_normals``.

Tolerances
  float64: |dz| <= 1e-13 (libdevice log/sqrt/sincospi vs NumPy, both ~1 ulp).
  float32: the device uses MUFU.LG2/SQRT/SIN/COS.  Documented worst cases: lg2.approx absolute
           error 2^-22, sin/cos.approx absolute error 2^-20.9 on (-pi, pi).  Propagated through
           z = r*trig(theta), r = sqrt(-2 ln2 * lg2(u)):  |dz| <= 6e-7*r + 2^-22*ln2/r + 4 ulp,
           rounded up to   tol = 2e-6 * (1 + |z|) + 4e-7 / r.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import philox
from spectralmc_b200 import _cabi, async_normals
from spectralmc_b200.async_normals import BufferConfig
from spectralmc_b200.numerical import Precision
from tests.helpers import expect_success

pytestmark = pytest.mark.gpu

SHAPES = [(12, 1024), (1, 16), (4, 4), (6, 8), (7, 132), (13, 1001), (5, 3), (252, 4096),
          # short float32 layout (rows <= 3: adjacent columns share a block), ragged column counts
          (1, 65536), (1, 1001), (2, 1000), (2, 7), (3, 4097), (3, 1)]


def _device_matrix(rows, cols, dtype, seed, k):
    out = torch.empty((rows, cols), dtype=dtype, device="cuda")
    _cabi.philox_normals(out, seed, k)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("rows,cols", SHAPES)
def test_float64_matches_specification(rows, cols) -> None:
    got = _device_matrix(rows, cols, torch.float64, seed=42, k=3)
    ref = philox.normals_matrix(rows, cols, np.float64, 42, 3)
    assert np.max(np.abs(got - ref)) <= 1e-13


@pytest.mark.parametrize("rows,cols", SHAPES)
def test_float32_matches_specification(rows, cols) -> None:
    got = _device_matrix(rows, cols, torch.float32, seed=42, k=3).astype(np.float64)
    ref, rad = philox.normals_matrix(rows, cols, np.float32, 42, 3, return_radius=True)
    tol = 2e-6 * (1 + np.abs(ref)) + 6e-7 * rad + 4e-7 / np.maximum(rad, 1e-6)
    err = np.abs(got - ref.astype(np.float64))
    assert np.all(err <= tol), float(np.max(err / tol))
    # the bulk is far tighter than the bound
    assert np.quantile(err, 0.99) <= 3e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_seed_and_index_select_the_matrix(dtype) -> None:
    np_dtype = np.float32 if dtype == torch.float32 else np.float64
    for seed, k in ((1, 0), (2**40 + 17, 0), (7, 2**33 + 5)):
        got = _device_matrix(6, 64, dtype, seed, k)
        ref = philox.normals_matrix(6, 64, np_dtype, seed, k)
        assert np.max(np.abs(got - ref)) <= (1e-4 if dtype == torch.float32 else 1e-13)
    a, b = _device_matrix(6, 64, dtype, 9, 0), _device_matrix(6, 64, dtype, 9, 1)
    assert not np.allclose(a, b)


def test_refined_tail_entries_match() -> None:
    """Blocks whose 21-bit radius field is zero (probability 2^-21) take the radius from the
    refinement block; locate them with the oracle and compare those entries explicitly."""
    cols = 1 << 19
    j = np.arange(cols, dtype=np.uint32)[None, :]
    q = np.arange(4, dtype=np.uint32)[:, None]
    radius, _ = philox.f32_fields(*philox.philox4x32_10((j, q, 0, 0), (7, 0)))
    hits = [(int(a), int(b), p) for p in range(3) for a, b in zip(*np.nonzero(radius[p] == 0))]
    assert hits
    got = _device_matrix(24, cols, torch.float32, seed=7, k=0).astype(np.float64)
    ref, rad = philox.normals_matrix(24, cols, np.float32, 7, 0, return_radius=True)
    assert np.all(np.isfinite(got))
    for qq, col, p in hits:
        for row in (6 * qq + 2 * p, 6 * qq + 2 * p + 1):
            assert rad[row, col] > 5.46
            assert abs(got[row, col] - ref[row, col]) <= 2e-6 * (1 + abs(ref[row, col])) + 6e-7 * rad[row, col] + 4e-7 / rad[row, col]


def test_unaligned_output_takes_the_scalar_path() -> None:
    base = torch.empty(8 * 64 + 1, dtype=torch.float32, device="cuda")
    view = base[1:].view(8, 64)  # 4-byte aligned only
    _cabi.philox_normals(view, 5, 0)
    torch.cuda.synchronize()
    ref = philox.normals_matrix(8, 64, np.float32, 5, 0)
    assert np.max(np.abs(view.cpu().numpy() - ref)) <= 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_moments_at_scale(dtype) -> None:
    """16.8M normals: mean, variance, kurtosis and tail mass of the device stream."""
    out = torch.empty((64, 1 << 18), dtype=dtype, device="cuda")
    _cabi.philox_normals(out, 123, 0)
    z = out.double()
    n = z.numel()
    assert abs(z.mean().item()) < 5 / n**0.5
    assert abs(z.var().item() - 1) < 5 * (2 / n) ** 0.5
    assert abs((z**4).mean().item() - 3) < 5 * (96 / n) ** 0.5
    tail = (z.abs() > 4).double().mean().item()
    assert abs(tail - 6.334e-5) < 5 * (6.334e-5 / n) ** 0.5
    assert torch.isfinite(z).all()


# ---- the reference's generator tests against the drop-in module ------------------------------
@pytest.fixture(params=[Precision.float32, Precision.float64])
def precision(request):
    return request.param


def _collect(gen_result, n):
    gen = expect_success(gen_result)
    return [expect_success(gen.get_matrix()) for _ in range(n)]


def test_private_norm_generator(precision) -> None:
    rows, cols = 4, 6
    gen = expect_success(async_normals._NormGenerator.create(rows, cols, dtype=precision.to_torch()))
    expect_success(gen.enqueue(123))
    first = expect_success(gen.get_matrix(456))
    second = expect_success(gen.get_matrix(789))
    assert first.shape == (rows, cols) and first.dtype == precision.to_torch()
    assert not torch.allclose(first, second)
    before = gen.get_time_spent_synchronizing()
    _ = expect_success(gen.get_matrix(111))
    assert gen.get_time_spent_synchronizing() >= before
    torch.cuda.synchronize()
    assert gen.is_ready() is True
    assert "QueueBusy" in str(gen.enqueue(5))
    assert "SeedOutOfRange" in str(async_normals._NormGenerator(2, 2, dtype=torch.float32).enqueue(0))


def test_checkpoint_reproducibility(precision) -> None:
    """Matrix k depends on (seed, k) only: identical after restore and for another buffer size."""
    rows, cols, buffer = 3, 5, 3
    cfg0 = async_normals.ConcurrentNormGeneratorConfig(rows=rows, cols=cols, seed=42, dtype=precision, skips=0)
    gen0 = async_normals.ConcurrentNormGenerator.create(BufferConfig.create(buffer, rows, cols), cfg0)
    initial = _collect(gen0, 10)
    snap = expect_success(gen0).snapshot()
    assert snap.skips == 10 and len(initial) == 10
    expected = _collect(gen0, 6)
    same = _collect(async_normals.ConcurrentNormGenerator.create(BufferConfig.create(buffer, rows, cols), snap), 6)
    diff = _collect(async_normals.ConcurrentNormGenerator.create(BufferConfig.create(buffer + 2, rows, cols), snap), 6)
    for e, s, d in zip(expected, same, diff, strict=True):
        assert torch.equal(e, s) and torch.equal(e, d)
    ref = philox.normals_matrix(rows, cols, precision.to_numpy(), 42, 12)
    assert np.max(np.abs(expected[2].cpu().numpy() - ref)) <= 1e-4


def test_diagnostics(precision) -> None:
    cfg = async_normals.ConcurrentNormGeneratorConfig(rows=2, cols=2, seed=7, dtype=precision, skips=0)
    gen = expect_success(async_normals.ConcurrentNormGenerator.create(BufferConfig.create(2, 2, 2), cfg))
    t0 = gen.get_time_spent_synchronizing()
    expect_success(gen.get_matrix())
    assert gen.get_time_spent_synchronizing() >= t0
    torch.cuda.synchronize()
    idle_before = gen.get_idle_time()
    expect_success(gen.get_matrix())
    torch.cuda.synchronize()
    assert gen.get_idle_time() >= idle_before
    assert gen.dtype == precision.to_torch()


def test_block_structure_leaves_no_statistical_trace() -> None:
    """The six normals of a float32 block come from disjoint bit fields of one Philox block; check
    on 25M draws that each row position has unit moments and that rows inside a block — in
    particular the (cos, sin) partners and rows sharing a Philox word — are uncorrelated, also in
    their squares (radius dependence shows up there)."""
    rows, cols = 12, 1 << 21
    out = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
    _cabi.philox_normals(out, 2024, 5)
    z = out.double()
    n = cols
    for i in range(6):
        r = z[i]
        assert abs(r.mean().item()) < 5 / n**0.5, i
        assert abs(r.var().item() - 1) < 5 * (2 / n) ** 0.5, i
        assert abs((r**3).mean().item()) < 5 * (15 / n) ** 0.5, i
        assert abs((r**4).mean().item() - 3) < 5 * (96 / n) ** 0.5, i
    zc = z[:6] - z[:6].mean(dim=1, keepdim=True)
    corr = (zc @ zc.T / n).cpu().numpy()
    sq = z[:6] ** 2 - 1.0
    corr_sq = (sq @ sq.T / n / 2.0).cpu().numpy()  # var(z^2) = 2
    off = ~np.eye(6, dtype=bool)
    assert np.max(np.abs(corr[off])) < 5 / n**0.5
    # partners of one pair share the radius: z0^2 + z1^2 = r^2 is chi2(2) => corr(z0^2, z1^2) = 0 for exact Box-Muller
    assert np.max(np.abs(corr_sq[off])) < 6 / n**0.5
    # consecutive blocks (rows 5 and 6) and consecutive columns are uncorrelated
    assert abs((z[5] * z[6]).mean().item()) < 5 / n**0.5
    assert abs((z[0, :-1] * z[0, 1:]).mean().item()) < 5 / n**0.5


def test_uniformity_of_the_sum_over_a_path() -> None:
    """Sum of T=252 normals of a path / sqrt(T) is N(0,1): KS test over 1M paths (what the
    log-Euler terminal price actually consumes)."""
    from scipy import stats

    T, P = 252, 1 << 20
    out = torch.empty((T, P), dtype=torch.float32, device="cuda")
    _cabi.philox_normals(out, 99, 0)
    s = (out.double().sum(dim=0) / T**0.5).cpu().numpy()
    assert stats.kstest(s, "norm").pvalue > 1e-3
    assert abs(s.std() - 1) < 5 / (2 * P) ** 0.5


@pytest.mark.parametrize("rows,cols", [(12, 1024), (7, 132), (1, 1001), (3, 4097), (252, 512)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_opt_in_philox7_stream_matches_its_specification(rows, cols, dtype) -> None:
    """stream_version 1 (Philox4x32-7, an explicit opt-in): the generator follows oracle/philox.py's statement of it,
    in both layouts and both precisions, and differs from the default stream."""
    np_dtype = np.float32 if dtype == torch.float32 else np.float64
    out = torch.empty((rows, cols), dtype=dtype, device="cuda")
    _cabi.philox_normals(out, 42, 3, _cabi.SMC_STREAM_PHILOX7)
    got = out.cpu().numpy().astype(np.float64)
    ref, rad = philox.normals_matrix(rows, cols, np_dtype, 42, 3, return_radius=True, stream_version=1)
    if dtype == torch.float64:
        assert np.max(np.abs(got - ref)) <= 1e-13
    else:
        tol = 2e-6 * (1 + np.abs(ref)) + 6e-7 * rad + 4e-7 / np.maximum(rad, 1e-6)
        assert np.all(np.abs(got - ref.astype(np.float64)) <= tol)
    assert not np.allclose(got, _device_matrix(rows, cols, dtype, 42, 3), atol=1e-3)
