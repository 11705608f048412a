"""The fused batch path (smc_cf_fused and the two-phase NORMALIZE API) against the oracle run on
the oracle's own statement of the same Philox normals, against analytic Black-76 (statistically,
as the reference's tests/test_gbm.py:103-139 does against QuantLib), and against itself through
size-independent properties (sharding, determinism, linearity).

Tolerances: the fused float64 path sees the same normals as the oracle to ~1e-15, so CF vectors
agree to 1e-12 norm-wise.  The fused float32 path draws its normals with MUFU (|dz| ~ 1e-6) and
sums them in float32; against the float64-arithmetic oracle that is ~1e-6 norm-wise on the CF —
the stated 1e-5 bound holds with margin for ordinary contracts and is asserted as such.
"""

from __future__ import annotations

import math

import numpy as np
import pytest
import torch

from oracle import gbm as ogbm
from oracle import philox
from oracle.black76 import black76
from oracle.sobol import sobol_contracts
from spectralmc_b200 import _cabi
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import BlackScholes, build_simulation_params
from spectralmc_b200.numerical import Precision
from spectralmc_b200.result import Failure
from tests.helpers import expect_success, make_black_scholes_config, make_simulation_params, rel_max

pytestmark = pytest.mark.gpu

CANON = (100.0, 100.0, 1.0, 0.05, 0.0, 0.2)
ODD = (37.5, 41.0, 2.5, -0.01, 0.03, 0.65)


def _oracle_cf(rows, T, N, B, np_dtype, seed, first_index, scheme, norm, stream_version=0):
    out = []
    for i, row in enumerate(rows):
        z = philox.normals_matrix(T, N * B, np_dtype, seed, first_index + i, stream_version=stream_version)
        cf, _ = ogbm.simulate_fft(ogbm.Contract(*row), z, N, scheme=scheme, normalization=norm)
        out.append(np.asarray(cf, dtype=np.complex128))
    return np.stack(out)


def _fused(rows, T, N, B, dtype, seed, first_index, scheme, norm, **shard):
    contracts = torch.tensor(np.asarray(rows, dtype=np.float64), device="cuda")
    args = _cabi.make_fused_args(contracts, len(rows), T, N, B, dtype, scheme, norm, seed, first_index, **shard)
    out = _cabi.cf_fused(args, contracts.device, dtype)
    torch.cuda.synchronize()
    return out.cpu().numpy()


CASES = [
    # T, N, B  — c1 shape, ragged shapes, N > 256, T not a multiple of 4 / 2
    (12, 16, 64), (1, 16, 32), (7, 12, 11), (5, 1, 40), (3, 300, 5), (252, 128, 16), (9, 64, 33), (2, 1024, 3),
    # short paths (T <= 3) in the grouped layout: several passes per tile, sub-lane folds (N = 48: five lanes per
    # column), N > 256, and shapes the grouped tile does not cover (256 G not a multiple of N: general form)
    (1, 128, 100), (2, 64, 70), (3, 32, 50), (1, 48, 37), (1, 512, 7), (1, 20, 33), (2, 3, 500), (4, 16, 40),
]


@pytest.mark.parametrize("T,N,B", CASES)
@pytest.mark.parametrize("scheme", ["log_euler", "simple_euler"])
@pytest.mark.parametrize("norm", ["raw_paths", "normalize_forwards"])
@pytest.mark.parametrize("prec", ["float64", "float32"])
def test_fused_matches_oracle(T, N, B, scheme, norm, prec) -> None:
    dtype = torch.float64 if prec == "float64" else torch.float32
    rows = [CANON, ODD, (90.0, 100.0, 0.0, 0.03, 0.01, 0.3), (90.0, 100.0, 2.0, 0.03, 0.01, 0.0)]
    ref = _oracle_cf(rows, T, N, B, np.dtype(prec), 42, 5, scheme, norm)
    got = _fused(rows, T, N, B, dtype, 42, 5,
                 _cabi.SMC_LOG_EULER if scheme == "log_euler" else _cabi.SMC_SIMPLE_EULER,
                 _cabi.SMC_NORMALIZE if norm == "normalize_forwards" else _cabi.SMC_RAW)
    tol = 1e-12 if prec == "float64" else 1e-5
    for c in range(len(rows)):
        assert rel_max(got[c], ref[c]) <= tol, (c, rel_max(got[c], ref[c]))
    assert got.dtype == (np.complex128 if prec == "float64" else np.complex64)


@pytest.mark.parametrize("T,N,B", [(1, 16, 4096), (2, 16, 999), (3, 128, 257), (1, 256, 300), (5, 16, 100)])
@pytest.mark.parametrize("scheme", [_cabi.SMC_LOG_EULER, _cabi.SMC_SIMPLE_EULER, _cabi.SMC_LOG_EULER_STEPWISE])
def test_short_paths_fused_equals_materialised(T, N, B, scheme) -> None:
    """T <= 3 uses the short layout of the stream (adjacent columns share a block): the fused kernel, the
    materialised generator + stepper and the oracle all read the same normals — at the reference's own test
    size (T = 1, N = 16, B = 4096; tests/test_gbm_trainer.py:127-136) and ragged neighbours; T = 5 is the
    general layout."""
    dtype = torch.float32
    fused = _fused([CANON, ODD], T, N, B, dtype, 77, 3, scheme, _cabi.SMC_RAW)
    for i, row in enumerate([CANON, ODD]):
        z = torch.empty((T, N * B), dtype=dtype, device="cuda")
        _cabi.philox_normals(z, 77, 3 + i)
        ref_z = philox.normals_matrix(T, N * B, np.float32, 77, 3 + i)
        assert float(np.max(np.abs(z.cpu().numpy() - ref_z))) <= 4e-6 * (1 + float(np.max(np.abs(ref_z))))
        X0, K, Tm, r, d, v = row
        k2 = _cabi.SMC_SIMPLE_EULER if scheme == _cabi.SMC_SIMPLE_EULER else _cabi.SMC_LOG_EULER
        term = _cabi.gbm_terminal_from_normals(z, Tm / T, X0, r, d, v, k2)
        put, _ = _cabi.payoff(term, K, math.exp(-r * Tm))
        mat = _cabi.cf_fft_mean(put.view(B, N)).cpu().numpy()
        assert rel_max(fused[i], mat) <= 1e-5, (i, rel_max(fused[i], mat))


@pytest.mark.parametrize("cuts", [[0, 5, 100], [0, 33, 34, 100], [0, 99, 100]])
def test_short_paths_shard_at_any_row(cuts) -> None:
    """Shard boundaries fall inside a group of columns sharing a block (N = 16, G = 6): both shards draw the
    block and each keeps its own lanes — the partial CFs still add up to the unsharded result."""
    T, N, B = 1, 16, 100
    whole = _fused([CANON, ODD], T, N, B, torch.float32, 5, 0, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW)
    total = sum(_fused([CANON, ODD], T, N, B, torch.float32, 5, 0, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, batch_begin=lo, batch_end=hi)
                for lo, hi in zip(cuts[:-1], cuts[1:]))
    assert rel_max(total, whole) <= 2e-6


@pytest.mark.parametrize("T,N,B", [(12, 16, 64), (252, 128, 16), (1, 16, 100), (3, 32, 50), (7, 12, 11)])
@pytest.mark.parametrize("norm", ["raw_paths", "normalize_forwards"])
@pytest.mark.parametrize("prec", ["float64", "float32"])
def test_opt_in_philox7_stream_through_the_fused_path(T, N, B, norm, prec) -> None:
    """stream_version 1: the fused kernels (a second build of the same translation unit with seven Philox rounds)
    against the oracle's statement of that stream; and the default stream is untouched by its presence."""
    dtype = torch.float64 if prec == "float64" else torch.float32
    rows = [CANON, ODD]
    contracts = torch.tensor(np.asarray(rows, dtype=np.float64), device="cuda")
    code = _cabi.SMC_NORMALIZE if norm == "normalize_forwards" else _cabi.SMC_RAW
    args = _cabi.make_fused_args(contracts, 2, T, N, B, dtype, _cabi.SMC_LOG_EULER, code, 42, 5, stream_version=_cabi.SMC_STREAM_PHILOX7)
    got = _cabi.cf_fused(args, contracts.device, dtype).cpu().numpy()
    ref = _oracle_cf(rows, T, N, B, np.dtype(prec), 42, 5, "log_euler", norm, stream_version=1)
    tol = 1e-12 if prec == "float64" else 1e-5
    for c in range(2):
        assert rel_max(got[c], ref[c]) <= tol, (c, rel_max(got[c], ref[c]))
    default = _fused(rows, T, N, B, dtype, 42, 5, _cabi.SMC_LOG_EULER, code)
    assert rel_max(default, _oracle_cf(rows, T, N, B, np.dtype(prec), 42, 5, "log_euler", norm)) <= tol
    assert rel_max(got, default) > 1e-4  # a different sample set


def test_opt_in_stream_through_the_engine_api() -> None:
    """SimulationParams(stream_version=1) selects it end to end: fused targets, the materialised generator, snapshots."""
    sp7 = make_simulation_params(timesteps=6, network_size=16, batches_per_mc_run=64, mc_seed=9, dtype=Precision.float32).model_copy(
        update={"stream_version": 1})
    cfg = make_black_scholes_config(sim_params=sp7, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
    engine = BlackScholes(cfg)
    contract = BlackScholes.Inputs(X0=CANON[0], K=CANON[1], T=CANON[2], r=CANON[3], d=CANON[4], v=CANON[5])
    fused = expect_success(engine.cf_targets([contract]))[0].cpu().numpy()
    ref = _oracle_cf([CANON], 6, 16, 64, np.dtype("float32"), 9, 0, "log_euler", "raw_paths", stream_version=1)[0]
    assert rel_max(fused, ref) <= 1e-5
    sims = expect_success(engine._simulate(contract)).sims.cpu().numpy()  # matrix 1 of the stream, materialised
    z = philox.normals_matrix(6, 16 * 64, np.float32, 9, 1, stream_version=1)
    ogbm.simulate_paths_inplace(z, 6, CANON[2] / 6, CANON[0], CANON[3], CANON[4], CANON[5], True)
    assert rel_max(sims, z) <= 2e-5
    assert expect_success(engine.snapshot()).sim_params.stream_version == 1


@pytest.mark.parametrize("prec", ["float64", "float32"])
@pytest.mark.parametrize("norm", [_cabi.SMC_RAW, _cabi.SMC_NORMALIZE])
def test_fine_tail_tile_plan_covers_every_row_exactly_once(prec, norm) -> None:
    """A single-contract launch of this size ends on a fine tail (main tiles of 14 rows, then tiles of 12 rows for the
    last rows; smc_cf_fused_plan), its four batch shards do not: the shards' partial CFs must add up to the whole —
    a row simulated twice or not at all would show at the 1e-5 level."""
    dtype = torch.float64 if prec == "float64" else torch.float32
    T, N, B = 12, 128, 40000
    contracts = torch.tensor(np.asarray([ODD]), device="cuda")
    whole_args = _cabi.make_fused_args(contracts, 1, T, N, B, dtype, _cabi.SMC_LOG_EULER, norm, 11, 2)
    plan = _cabi.cf_fused_plan(whole_args)
    assert plan["main_tiles"] < plan["tiles"] and plan["tail_tile_rows"] < plan["tile_rows"], plan
    whole = _cabi.cf_fused(whole_args, contracts.device, dtype).cpu().numpy()
    cuts = [0, 10000, 20001, 29999, 40000]
    shards = [_cabi.make_fused_args(contracts, 1, T, N, B, dtype, _cabi.SMC_LOG_EULER, norm, 11, 2, batch_begin=lo, batch_end=hi)
              for lo, hi in zip(cuts[:-1], cuts[1:])]
    for a in shards:
        p = _cabi.cf_fused_plan(a)
        assert p["main_tiles"] == p["tiles"], p  # no fine tail in the shards
    if norm == _cabi.SMC_RAW:
        total = sum(_cabi.cf_fused(a, contracts.device, dtype).cpu().numpy() for a in shards)
    else:
        staged = [_cabi.fused_terminal(a, contracts.device, dtype) for a in shards]
        tsum = sum(s for _, s in staged)
        total = sum(_cabi.cf_from_terminal(a, t, tsum, dtype).cpu().numpy() for a, (t, _) in zip(shards, staged))
    assert rel_max(total, whole) <= (1e-13 if prec == "float64" else 2e-6)


@pytest.mark.parametrize("prec", ["float64", "float32"])
def test_stepwise_exponential_variant_agrees(prec) -> None:
    """prod_j exp(x_j) == exp(sum_j x_j): the per-step-exp kernel and the log-sum kernel agree."""
    dtype = torch.float64 if prec == "float64" else torch.float32
    a = _fused([CANON, ODD], 50, 32, 64, dtype, 9, 0, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW)
    b = _fused([CANON, ODD], 50, 32, 64, dtype, 9, 0, _cabi.SMC_LOG_EULER_STEPWISE, _cabi.SMC_RAW)
    assert rel_max(a, b) <= (1e-12 if prec == "float64" else 1e-5)


@pytest.mark.parametrize("prec", ["float64", "float32"])
def test_fused_equals_materialised_pipeline(prec) -> None:
    """Same counters => same normals: K1 -> K2 -> K6 -> K7/K8 on HBM equals the fused kernel."""
    dtype = torch.float64 if prec == "float64" else torch.float32
    T, N, B = 37, 64, 128
    fused = _fused([ODD], T, N, B, dtype, 77, 3, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW)[0]
    z = torch.empty((T, N * B), dtype=dtype, device="cuda")
    _cabi.philox_normals(z, 77, 3)
    X0, K, Tm, r, d, v = ODD
    term = _cabi.gbm_terminal_from_normals(z, Tm / T, X0, r, d, v, _cabi.SMC_LOG_EULER)
    put, _ = _cabi.payoff(term, K, math.exp(-r * Tm))
    mat = _cabi.cf_fft_mean(put.view(B, N)).cpu().numpy()
    assert rel_max(fused, mat) <= (1e-12 if prec == "float64" else 1e-5)


@pytest.mark.parametrize("prec", ["float64", "float32"])
@pytest.mark.parametrize("norm", [_cabi.SMC_RAW, _cabi.SMC_NORMALIZE])
def test_batch_sharding_sums_to_the_single_device_result(prec, norm) -> None:
    """Ranks simulate disjoint batch rows with GLOBAL path counters; partial CFs add up
    (SURVEY.md §8e).  NORMALIZE goes through the two-phase API with the summed terminal sums."""
    dtype = torch.float64 if prec == "float64" else torch.float32
    T, N, B = 10, 32, 96
    rows = [CANON, ODD, (5.0, 4.0, 0.7, 0.1, 0.0, 1.1)]
    whole = _fused(rows, T, N, B, dtype, 11, 2, _cabi.SMC_LOG_EULER, norm)
    contracts = torch.tensor(np.asarray(rows), device="cuda")
    cuts = [0, 17, 64, 96]
    shards = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        shards.append(_cabi.make_fused_args(contracts, len(rows), T, N, B, dtype, _cabi.SMC_LOG_EULER, norm, 11, 2,
                                            batch_begin=lo, batch_end=hi))
    if norm == _cabi.SMC_RAW:
        total = sum(_cabi.cf_fused(a, contracts.device, dtype).cpu().numpy() for a in shards)
    else:
        staged = [_cabi.fused_terminal(a, contracts.device, dtype) for a in shards]
        tsum = sum(s for _, s in staged)  # the allreduce
        total = sum(_cabi.cf_from_terminal(a, t, tsum, dtype).cpu().numpy() for a, (t, _) in zip(shards, staged))
        with pytest.raises(_cabi.SmcError, match="NORMALIZE"):
            _cabi.cf_fused(shards[0], contracts.device, dtype)
    assert rel_max(total, whole) <= (1e-13 if prec == "float64" else 2e-6)


def test_bit_reproducible_and_skip_semantics() -> None:
    """Fixed-order reductions: identical bits run to run; contract c of a batch starting at matrix k
    equals a single-contract call at matrix k + c (each contract consumes one matrix, gbm.py:405)."""
    rows = [CANON, ODD, CANON]
    a = _fused(rows, 12, 16, 512, torch.float32, 42, 10, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW)
    b = _fused(rows, 12, 16, 512, torch.float32, 42, 10, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW)
    assert np.array_equal(a, b)
    single = _fused([CANON], 12, 16, 512, torch.float32, 42, 12, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW)
    assert np.array_equal(a[2], single[0])
    assert not np.array_equal(a[0], a[2])  # same contract, different matrix


@pytest.mark.parametrize("pin_in,pin_out", [(True, True), (False, False), (True, False), (False, True)])
def test_host_buffer_entry_point(pin_in, pin_out) -> None:
    """Pinned buffers are read / written through their device alias (no staging copy), pageable ones
    through the staged copies; every combination gives the device-resident result bit for bit."""
    rows = torch.tensor([CANON, ODD], dtype=torch.float64)
    out = torch.full((2, 16), float("nan"), dtype=torch.complex64)
    rows = rows.pin_memory() if pin_in else rows
    out = out.pin_memory() if pin_out else out
    args = _cabi.make_fused_args(None, 2, 12, 16, 64, torch.float32, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, 42, 0)
    ws = torch.empty(_cabi.cf_fused_host_workspace_bytes(args), dtype=torch.uint8, device="cuda")
    _cabi.cf_fused_host(args, rows, out, ws)
    dev = _fused([CANON, ODD], 12, 16, 64, torch.float32, 42, 0, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW)
    assert np.array_equal(out.numpy(), dev)


def test_host_buffer_entry_point_large_batch_takes_the_staged_copies() -> None:
    """Above 64 KiB the pinned buffers go through cudaMemcpyAsync (DMA beats kernel stores over PCIe)."""
    C, N = 600, 32  # 600 * 32 * 8 B = 150 KiB of targets
    rows_np = sobol_contracts(C, seed=9)
    rows = torch.tensor(rows_np, dtype=torch.float64).pin_memory()
    out = torch.full((C, N), float("nan"), dtype=torch.complex64).pin_memory()
    args = _cabi.make_fused_args(None, C, 3, N, 8, torch.float32, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, 42, 0)
    ws = torch.empty(_cabi.cf_fused_host_workspace_bytes(args), dtype=torch.uint8, device="cuda")
    _cabi.cf_fused_host(args, rows, out, ws)
    dev = _fused(rows_np, 3, N, 8, torch.float32, 42, 0, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW)
    assert np.array_equal(out.numpy(), dev)


def test_error_paths() -> None:
    contracts = torch.tensor([CANON], dtype=torch.float64, device="cuda")
    args = _cabi.make_fused_args(contracts, 1, 12, 16, 64, torch.float32, 0, 1, 42, 0)
    out = torch.empty((1, 16), dtype=torch.complex64, device="cuda")
    tiny = torch.empty(8, dtype=torch.uint8, device="cuda")
    rc = _cabi.LIB.smc_cf_fused(_cabi.byref(args), out.data_ptr(), tiny.data_ptr(), 8, None)
    assert rc == 3 and b"workspace" in _cabi.LIB.smc_last_error()
    args.batch_end = 65
    rc = _cabi.LIB.smc_cf_fused(_cabi.byref(args), out.data_ptr(), tiny.data_ptr(), 8, None)
    assert rc == 1 and b"batch range" in _cabi.LIB.smc_last_error()
    with pytest.raises(_cabi.SmcError, match="ROW_FFT") as info:
        _cabi.cf_fft_mean(torch.zeros((4, 7), device="cuda"), _cabi.SMC_CF_ROW_FFT)  # N not a power of two
    assert info.value.code == 4


# ---- statistical validation against analytic Black-Scholes ------------------------------------
def _engine(precision, *, T=1, N=256, B=2**12, scheme=PathScheme.LOG_EULER, norm=ForwardNormalization.RAW, skip=0, seed=7):
    sp = make_simulation_params(timesteps=T, network_size=N, batches_per_mc_run=B, threads_per_block=256,
                                mc_seed=seed, buffer_size=1, skip=skip, dtype=precision)
    return BlackScholes(make_black_scholes_config(sim_params=sp, path_scheme=scheme, normalization=norm))


@pytest.mark.parametrize("precision", [Precision.float32, Precision.float64])
@pytest.mark.parametrize("T", [1, 16])
def test_fused_prices_match_black76(precision, T) -> None:
    """64 Sobol contracts (seed 31) x 16 MC repetitions; DC bin / N is the put price.  Same
    acceptance as the reference (tests/test_gbm.py:61-63,135-139): <= 5 % of contracts beyond 3
    standard errors, RMSPE <= 0.15 where the analytic price is >= 1."""
    rows = sobol_contracts(64, seed=31)
    engine = _engine(precision, T=T)
    N = 256
    reps = []
    for _ in range(16):
        cf = expect_success(engine.cf_targets(torch.tensor(rows, device="cuda")))
        reps.append(cf[:, 0].real.double().cpu().numpy() / N)
    vals = np.stack(reps)  # [16, 64]
    mean, sd = vals.mean(axis=0), vals.std(axis=0, ddof=1)
    analytic = np.array([black76(*r)["put_price"] for r in rows])
    se = sd / math.sqrt(vals.shape[0])
    eps = 1e-4 if precision is Precision.float32 else 1e-8
    z = np.where(se > 0, np.abs(mean - analytic) / np.where(se > 0, se, 1), np.where(np.abs(mean - analytic) <= eps * np.maximum(analytic, 1), 0.0, 4.0))
    big = analytic >= 1.0
    rmspe = float(np.sqrt(np.mean(((mean[big] - analytic[big]) / analytic[big]) ** 2)))
    assert float(np.mean(z > 3.0)) <= 0.05, np.sort(z)[-5:]
    assert rmspe <= 0.15
    assert expect_success(engine.snapshot()).sim_params.skip == 16 * 64


def test_full_size_headline_config_statistics() -> None:
    """BASELINE config c2 at full size (fp32, T=252, N=128, B=65 536; 2.1e9 path-steps): the DC
    bin prices the canonical put within 4 standard errors of Black-76; off-DC bins are sampling
    noise of the predicted size; the imaginary DC part is exactly zero."""
    engine = _engine(Precision.float32, T=252, N=128, B=65536, seed=7)
    cf = expect_success(engine.simulate_fft(BlackScholes.Inputs(X0=100, K=100, T=1.0, r=0.05, d=0.0, v=0.2))).cpu().numpy()
    ref = black76(*CANON)["put_price"]
    put_sd = 7.2  # std of the discounted put payoff for this contract (analytic: ~7.2)
    se = put_sd / math.sqrt(128 * 65536)
    assert abs(cf[0].real / 128 - ref) <= 4 * se + 2e-5 * ref
    assert cf[0].imag == 0.0
    noise = put_sd * math.sqrt(128 / 65536)  # |CF_k| scale for k != 0
    assert np.max(np.abs(cf[1:])) <= 6 * noise
    assert np.sqrt(np.mean(np.abs(cf[1:]) ** 2)) > 0.3 * noise
    # Hermitian symmetry of the DFT of a real vector
    assert np.max(np.abs(cf[1:] - np.conj(cf[1:][::-1]))) <= 1e-5 * abs(cf[0])


@pytest.mark.parametrize("C,T,N,B", [(4096, 12, 128, 8), (70000, 1, 16, 16)])
def test_many_contracts_in_one_call(C, T, N, B) -> None:
    """c5-like contract counts (and > 65 535, which spills into gridDim.z): sampled contracts of the
    batch equal single-contract calls at the matching matrix index."""
    rows = sobol_contracts(C, seed=3)
    big = _fused(rows, T, N, B, torch.float32, 11, 0, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW)
    assert np.all(np.isfinite(big))
    for c in (0, 1, C // 2, 65535 if C > 65536 else C // 3, C - 1):
        one = _fused(rows[c : c + 1], T, N, B, torch.float32, 11, c, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW)
        assert np.array_equal(big[c], one[0]) or rel_max(big[c], one[0]) <= 1e-6, c
    norm = _fused(rows[:300], T, N, B, torch.float32, 11, 0, _cabi.SMC_LOG_EULER, _cabi.SMC_NORMALIZE)
    assert np.all(np.isfinite(norm))


def test_paths_that_need_the_refinement_block() -> None:
    """The fused float32 loop checks for zero radius fields once per PATH and re-simulates those
    paths with the refinement applied.  Locate such paths with the oracle (seed 7, matrix 0 has
    several within 2^19 columns and 4 row groups) and compare every stored terminal price."""
    T, N, B = 24, 128, 4096
    P = N * B
    j = np.arange(P, dtype=np.uint32)[None, :]
    q = np.arange(T // 6, dtype=np.uint32)[:, None]
    radius, _ = philox.f32_fields(*philox.philox4x32_10((j, q, 0, 0), (7, 0)))
    cols = sorted({int(b) for p in range(3) for b in np.nonzero(radius[p] == 0)[1]})
    assert cols, "no path needs the refinement block in this range"
    contracts = torch.tensor([CANON], dtype=torch.float64, device="cuda")
    args = _cabi.make_fused_args(contracts, 1, T, N, B, torch.float32, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, 7, 0)
    terminal, tsum = _cabi.fused_terminal(args, contracts.device, torch.float32)
    got = terminal[0].cpu().numpy().astype(np.float64)
    z = philox.normals_matrix(T, P, np.float32, 7, 0)
    ref = z.copy()
    ogbm.simulate_paths_inplace(ref, T, 1.0 / T, 100.0, 0.05, 0.0, 0.2, True)
    ref_t = ref[-1].astype(np.float64)
    assert np.max(np.abs(got - ref_t) / ref_t) <= 1e-5
    # the refined paths differ visibly from what an unrefined radius (u = 2^-22, r = 5.52) would give
    for c in cols:
        assert abs(got[c] - ref_t[c]) / ref_t[c] <= 1e-5, c
    for scheme in (_cabi.SMC_SIMPLE_EULER, _cabi.SMC_LOG_EULER_STEPWISE):
        a2 = _cabi.make_fused_args(contracts, 1, T, N, B, torch.float32, scheme, _cabi.SMC_RAW, 7, 0)
        t2, _ = _cabi.fused_terminal(a2, contracts.device, torch.float32)
        ref2 = z.copy()
        ogbm.simulate_paths_inplace(ref2, T, 1.0 / T, 100.0, 0.05, 0.0, 0.2, scheme != _cabi.SMC_SIMPLE_EULER)
        g2 = t2[0].cpu().numpy().astype(np.float64)
        assert np.max(np.abs(g2[cols] - ref2[-1][cols]) / ref2[-1][cols]) <= 2e-5


def test_empty_contract_batch() -> None:
    engine = _engine(Precision.float32, T=4, N=16, B=8)
    out = expect_success(engine.cf_targets([]))
    assert out.shape == (0, 16) and out.dtype == torch.complex64
    assert expect_success(engine.snapshot()).sim_params.skip == 0


def test_random_shapes_against_oracle() -> None:
    """40 random (C, T, N, B, scheme, normalisation, dtype) draws, including N that are not powers of
    two, N > 256, T not a multiple of 6, single rows — each against the oracle on the same counters."""
    rng = np.random.default_rng(12345)
    for trial in range(40):
        C = int(rng.integers(1, 4))
        T = int(rng.integers(1, 20))
        N = int(rng.choice([1, 2, 3, 7, 16, 31, 32, 100, 128, 257, 300, 512]))
        B = int(rng.integers(1, 40))
        prec = "float32" if rng.random() < 0.5 else "float64"
        scheme = "log_euler" if rng.random() < 0.6 else "simple_euler"
        norm = "raw_paths" if rng.random() < 0.5 else "normalize_forwards"
        rows = [(float(rng.uniform(1, 200)), float(rng.uniform(1, 200)), float(rng.uniform(0.05, 3)), float(rng.uniform(-0.1, 0.1)),
                 float(rng.uniform(-0.1, 0.1)), float(rng.uniform(0.05, 0.8))) for _ in range(C)]
        seed, first = int(rng.integers(1, 2**40)), int(rng.integers(0, 1000))
        dtype = torch.float64 if prec == "float64" else torch.float32
        ref = _oracle_cf(rows, T, N, B, np.dtype(prec), seed, first, scheme, norm)
        got = _fused(rows, T, N, B, dtype, seed, first,
                     _cabi.SMC_LOG_EULER if scheme == "log_euler" else _cabi.SMC_SIMPLE_EULER,
                     _cabi.SMC_NORMALIZE if norm == "normalize_forwards" else _cabi.SMC_RAW)
        tol = 1e-12 if prec == "float64" else 1e-5
        for c in range(C):
            scale = max(float(np.max(np.abs(ref[c]))), 1e-30)
            assert float(np.max(np.abs(got[c] - ref[c]))) / scale <= tol or scale < 1e-9, (trial, C, T, N, B, prec, scheme, norm)


def test_baseline_config_c4_against_black76() -> None:
    """BASELINE.json configs[3] at full size: float64, 512 Sobol contracts (seed 31, as
    tests/test_gbm.py:110 of the reference), 365 timesteps, network_size 256, B = 4096 — 1.96e11
    path-steps per repetition.  8 repetitions; acceptance as the reference's (<= 5 % of contracts
    beyond 3 standard errors of the analytic price, RMSPE <= 0.15 where the price is >= 1)."""
    rows = sobol_contracts(512, seed=31)
    engine = _engine(Precision.float64, T=365, N=256, B=4096, seed=7)
    dev_rows = torch.tensor(rows, device="cuda")
    vals = np.stack([expect_success(engine.cf_targets(dev_rows))[:, 0].real.cpu().numpy() / 256 for _ in range(8)])
    mean, sd = vals.mean(axis=0), vals.std(axis=0, ddof=1)
    analytic = np.array([black76(*r)["put_price"] for r in rows])
    se = sd / math.sqrt(vals.shape[0])
    z = np.where(se > 0, np.abs(mean - analytic) / np.where(se > 0, se, 1.0), np.where(np.abs(mean - analytic) <= 1e-8 * np.maximum(analytic, 1), 0.0, 4.0))
    big = analytic >= 1.0
    assert float(np.mean(z > 3.0)) <= 0.05, np.sort(z)[-8:]
    assert float(np.sqrt(np.mean(((mean[big] - analytic[big]) / analytic[big]) ** 2))) <= 0.15  # mask first: analytic is 0 for T = 0 rows
    assert expect_success(engine.snapshot()).sim_params.skip == 8 * 512


@pytest.mark.parametrize("prec", ["float64", "float32"])
def test_top_of_the_32_bit_path_counter(prec) -> None:
    """Maximum sizes: the ABI admits batches_total * N <= 2^32 - 1 global paths.  A rank that owns the
    LAST batch rows of such a job draws counters just below 2^32; its partial CF equals the oracle's
    on the same column slice (no wrap-around, no sign trouble in the 64-bit index arithmetic)."""
    dtype, np_dtype = (torch.float64, np.float64) if prec == "float64" else (torch.float32, np.float32)
    T, N = 7, 64
    B_total = 2**26 - 1  # 64 * (2^26 - 1) = 2^32 - 64 paths
    rows_local = 8
    b0 = B_total - rows_local
    got = _fused([CANON, ODD], T, N, B_total, dtype, 42, 3, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, batch_begin=b0, batch_end=B_total)
    for i, row in enumerate((CANON, ODD)):
        z = philox.normals_matrix(T, N * B_total, np_dtype, 42, 3 + i, col_begin=b0 * N, col_end=B_total * N)
        cf, _ = ogbm.simulate_fft(ogbm.Contract(*row), z, N, scheme="log_euler", normalization=ogbm.RAW)
        ref = np.asarray(cf, dtype=np.complex128) * (rows_local / B_total)  # partial: (1 / B_total) * sum over local rows
        assert rel_max(got[i], ref) <= (1e-12 if prec == "float64" else 1e-5)
    # one more row would pass 2^32 - 1 paths: rejected with a message, nothing launched
    contracts = torch.tensor([CANON], dtype=torch.float64, device="cuda")
    args = _cabi.make_fused_args(contracts, 1, T, N, 2**26, dtype, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, 42, 0, batch_begin=0, batch_end=8)
    with pytest.raises(_cabi.SmcError) as err:
        _cabi.cf_fused(args, contracts.device, dtype)
    assert err.value.code == 1 and "32-bit path counter" in err.value.message


def test_engine_at_the_reference_memory_guard() -> None:
    """The reference's soft cap (gbm.py:123-137): 1e9 float32 paths per contract.  The fused path never
    materialises them, so a full-guard contract is one ordinary call: N = 1000 (table DFT), B = 10^6, T = 1."""
    engine = _engine(Precision.float32, T=1, N=1000, B=1_000_000, seed=5)
    cf = expect_success(engine.simulate_fft(BlackScholes.Inputs(X0=100, K=100, T=1.0, r=0.05, d=0.0, v=0.2))).cpu().numpy()
    ref = black76(*CANON)["put_price"]
    se = 7.2 / math.sqrt(1e9)
    assert cf.shape == (1000,) and abs(cf[0].real / 1000 - ref) <= 5 * se + 2e-5 * ref and cf[0].imag == 0.0
    assert np.max(np.abs(cf[1:] - np.conj(cf[1:][::-1]))) <= 1e-5 * abs(cf[0])
    too_big = build_simulation_params(timesteps=1, network_size=1001, batches_per_mc_run=1_000_000, threads_per_block=256,
                                      mc_seed=5, buffer_size=1, dtype=Precision.float32)
    assert isinstance(too_big, Failure)


def _zero_radius_hits(seed: int, k: int, cols: int, row_groups) -> list[tuple[int, int, int]]:
    """(row-group word, column, pair) of every float32 Box-Muller pair of matrix ``k`` with a zero radius field (the oracle's Philox)."""
    j = np.arange(cols, dtype=np.uint32)[None, :]
    q = np.asarray(row_groups, dtype=np.uint32)[:, None]
    x = philox.philox4x32_10((j, q, k & 0xFFFFFFFF, k >> 32), (seed & 0xFFFFFFFF, seed >> 32))
    radius, _ = philox.f32_fields(*x)
    return [(int(q[a, 0]), int(b), pair) for pair in range(3) for a, b in zip(*np.nonzero(radius[pair] == 0))]


@pytest.mark.parametrize("layout", ["general", "short"])
def test_rare_refinement_path_of_the_fused_kernels(layout) -> None:
    """A zero radius field (2^-21 per pair) sends the fused kernels down an out-of-line path — the whole path re-simulated
    with the refinement (long paths), the group's block redone (short tile).  At oracle-sized shapes that path is almost never
    taken, so this test LOOKS for such pairs with the oracle's Philox and runs the one batch row that contains each hit:
    every path of the row is its own CF column, so a wrong refinement shows (the test also proves that it would)."""
    seed, k = 2026, 11
    if layout == "general":
        T, N = 12, 128  # two row groups: one normal in twelve carries enough of the path for a changed radius to show
        paths = sorted({col for _, col, _ in _zero_radius_hits(seed, k, 1 << 20, range(T // 6))})
    else:  # T = 1: six adjacent paths share a block, pair p serves paths 6 g + 2 p and 6 g + 2 p + 1
        T, N = 1, 16
        paths = sorted({6 * g + 2 * pair for _, g, pair in _zero_radius_hits(seed, k, 1 << 22, [philox.F32_SHORT_BIT])})
    assert paths, "no zero radius field found: enlarge the search"
    row_contract = (100.0, 150.0, 1.0, 0.05, 0.0, 0.2)  # deep in the money: the payoff is linear in the terminal price
    contract = ogbm.Contract(*row_contract)
    checked = 0
    for path in paths[:6]:
        b = path // N
        B_total = b + 2
        got = _fused([row_contract], T, N, B_total, torch.float32, seed, k, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, batch_begin=b, batch_end=b + 1)[0]
        z, rad = philox.normals_matrix(T, N * B_total, np.float32, seed, k, col_begin=b * N, col_end=(b + 1) * N, return_radius=True)
        refined = rad > 5.39  # sqrt(-2 ln u) for u < 2^-21: only refined draws reach it
        assert refined.any(), (layout, path)
        # what a kernel that skipped the refinement would produce (radius of the unrefined uniform 2^-22)
        z_coarse = np.where(refined, z * (math.sqrt(-2.0 * math.log(2.0**-22)) / np.maximum(rad, 1e-30)), z).astype(np.float32)
        ref, _ = ogbm.simulate_fft(contract, z.copy(), N, scheme="log_euler", normalization="raw_paths")  # (the oracle steps in place)
        coarse, _ = ogbm.simulate_fft(contract, z_coarse, N, scheme="log_euler", normalization="raw_paths")
        ref = np.asarray(ref, dtype=np.complex128) / B_total  # one row of B_total: the shard's share of the batch mean
        coarse = np.asarray(coarse, dtype=np.complex128) / B_total
        tol = 1e-5
        if rel_max(coarse, ref) > 20 * tol:  # this hit is visible (its radius moved enough): the kernel must follow the specification
            assert rel_max(got, ref) <= tol, (layout, path, rel_max(got, ref), rel_max(coarse, ref))
            checked += 1
    assert checked, "no hit moved the result enough to be checked"
