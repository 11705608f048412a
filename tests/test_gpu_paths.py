"""K2 / K4-K8 / K11 parity on the reference's own golden vectors (tests/golden/*.npz): the
reference's normal draws are injected into the CUDA kernels through the C ABI and every stage is
compared with what the reference produced.

Tolerances (BASELINE.json north_star): 1e-12 relative for float64, 1e-5 for float32;
element-wise for path values and payoffs, norm-wise (max|d|/max|ref|) for CF vectors.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import gbm as ogbm
from oracle import philox
from spectralmc_b200 import _cabi
from spectralmc_b200.gbm import SimulateBlackScholes
from tests.helpers import rel_elem, rel_max

pytestmark = pytest.mark.gpu


def _tol(dtype) -> float:
    return 1e-5 if np.dtype(dtype) == np.float32 else 1e-12


def _scheme(g) -> int:
    return _cabi.SMC_LOG_EULER if str(g["scheme"]) == "log_euler" else _cabi.SMC_SIMPLE_EULER


def _kernel_only_reference(g) -> np.ndarray:
    """Raw kernel output for the golden normals (the golden `sims` are post-normalisation)."""
    X0, K, T, r, d, v = (float(x) for x in g["contract"])
    io = g["normals"].copy()
    ogbm.simulate_paths_inplace(io, io.shape[0], T / io.shape[0], X0, r, d, v, str(g["scheme"]) == "log_euler")
    return io


def test_path_kernel_inplace(golden) -> None:
    g = golden
    X0, K, T, r, d, v = (float(x) for x in g["contract"])
    rows = int(g["timesteps"])
    io = torch.from_numpy(g["normals"].copy()).cuda()
    SimulateBlackScholes[1, int(g["threads_per_block"]), torch.cuda.current_stream()](
        io, rows, T / rows, X0, r, d, v, str(g["scheme"]) == "log_euler"
    )
    got = io.cpu().numpy()
    if str(g["normalization"]) == "raw_paths":
        assert rel_elem(got, g["sims"]) <= _tol(got.dtype), g["name"]  # reference output directly
    assert rel_elem(got, _kernel_only_reference(g)) <= _tol(got.dtype), g["name"]


def test_launch_object_accepts_what_the_reference_passes(golden) -> None:
    """The reference launches with ``cuda.as_cuda_array(sims)`` and a Numba stream (gbm.py:413-426,
    effects/interpreter.py:634-654): a Numba device array (``__cuda_array_interface__``) and a ``numba.cuda.stream()``
    must go through the same launch object, zero-copy and in place."""
    numba_cuda = pytest.importorskip("numba.cuda")
    g = golden
    X0, K, T, r, d, v = (float(x) for x in g["contract"])
    rows = int(g["timesteps"])
    torch.cuda.synchronize()
    io = numba_cuda.to_device(g["normals"].copy())
    stream = numba_cuda.stream()
    SimulateBlackScholes[1, int(g["threads_per_block"]), stream](io, rows, T / rows, X0, r, d, v, str(g["scheme"]) == "log_euler")
    stream.synchronize()
    got = io.copy_to_host()
    if str(g["normalization"]) == "raw_paths":
        assert rel_elem(got, g["sims"]) <= _tol(got.dtype), g["name"]
    assert rel_elem(got, _kernel_only_reference(g)) <= _tol(got.dtype), g["name"]


def test_launch_object_accepts_dlpack_and_raw_stream_handles() -> None:
    z = torch.randn(5, 64, device="cuda", dtype=torch.float64)
    want = z.clone()
    SimulateBlackScholes[1, 256](want, 5, 0.2, 100.0, 0.05, 0.0, 0.2, True)

    class OnlyDLPack:  # a foreign array that exports DLPack and nothing else
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, **kw):
            return self._t.__dlpack__(**kw)

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    a, b = z.clone(), z.clone()
    SimulateBlackScholes[1, 256, torch.cuda.current_stream().cuda_stream](OnlyDLPack(a), 5, 0.2, 100.0, 0.05, 0.0, 0.2, True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    SimulateBlackScholes[1, 256, side](b, 5, 0.2, 100.0, 0.05, 0.0, 0.2, True)
    side.synchronize()
    torch.cuda.synchronize()
    assert torch.equal(a, want) and torch.equal(b, want)
    with pytest.raises(ValueError):  # a host array (NumPy exports DLPack too) is refused loudly: there is no CPU path
        SimulateBlackScholes[1, 256](z.cpu().numpy(), 5, 0.2, 100.0, 0.05, 0.0, 0.2, True)
    with pytest.raises(TypeError):
        SimulateBlackScholes[1, 256]([[1.0, 2.0]], 1, 0.2, 100.0, 0.05, 0.0, 0.2, True)
    with pytest.raises(ValueError):
        SimulateBlackScholes[1, 256](z.t(), 64, 0.2, 100.0, 0.05, 0.0, 0.2, True)  # not C-contiguous


def test_terminal_only_kernel(golden) -> None:
    g = golden
    X0, K, T, r, d, v = (float(x) for x in g["contract"])
    rows = int(g["timesteps"])
    z = torch.from_numpy(g["normals"].copy()).cuda()
    term = _cabi.gbm_terminal_from_normals(z, T / rows, X0, r, d, v, _scheme(g)).cpu().numpy()
    assert torch.equal(z.cpu(), torch.from_numpy(g["normals"]))  # input untouched
    assert rel_elem(term, _kernel_only_reference(g)[-1]) <= _tol(term.dtype)


def test_normalise_payoff_cf_and_host_price(golden) -> None:
    g = golden
    X0, K, T, r, d, v = (float(x) for x in g["contract"])
    rows, n, b = int(g["timesteps"]), int(g["network_size"]), int(g["batches"])
    tol = _tol(g["normals"].dtype)
    io = torch.from_numpy(g["normals"].copy()).cuda()
    _cabi.gbm_paths_inplace(io, T / rows, X0, r, d, v, _scheme(g), int(g["threads_per_block"]))
    if str(g["normalization"]) == "normalize_forwards":
        _cabi.normalize_rows(io, torch.from_numpy(g["forwards"]).cuda())
    assert rel_elem(io.cpu().numpy(), g["sims"]) <= tol
    put, call = _cabi.payoff(io[-1], K, float(g["df"][-1]))
    # payoffs that are exactly zero in the reference must be zero here unless the path sits on the strike
    assert rel_max(put.cpu().numpy(), g["put_price"]) <= tol
    assert rel_max(call.cpu().numpy(), g["call_price"]) <= tol
    cf = _cabi.cf_fft_mean(put.view(b, n)).cpu().numpy()
    assert cf.dtype == g["cf"].dtype
    assert rel_max(cf, g["cf"]) <= tol
    means = _cabi.means3(io[-1].contiguous(), put, call).cpu().numpy()
    assert rel_elem(means, g["host"][[2, 5, 6]], floor=1e-6 * abs(g["host"][2])) <= tol


def test_cf_of_the_reference_payoffs(golden) -> None:
    """K7+K8 alone: the reference's own put_price matrix in, the reference's CF out."""
    g = golden
    n, b = int(g["network_size"]), int(g["batches"])
    mat = torch.from_numpy(g["put_price"].copy()).cuda().view(b, n)
    cf = _cabi.cf_fft_mean(mat).cpu().numpy()
    assert rel_max(cf, g["cf"]) <= (2e-6 if mat.dtype == torch.float32 else 1e-13)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("tpb", [32, 64, 128, 256, 512, 1024])
def test_threads_per_block_and_ragged_widths(dtype, tpb) -> None:
    """Every legal CTA size; widths that defeat the vector path; idle tail threads (gbm.py:242)."""
    for cols in (1, 3, 130, 1000, 4096 + 4):
        z = philox.normals_matrix(9, cols, dtype, 17, 0)
        ref = z.copy()
        ogbm.simulate_paths_inplace(ref, 9, 0.3 / 9, 50.0, 0.02, 0.01, 0.4, True)
        io = torch.from_numpy(z).cuda()
        _cabi.gbm_paths_inplace(io, 0.3 / 9, 50.0, 0.02, 0.01, 0.4, _cabi.SMC_LOG_EULER, tpb)
        assert rel_elem(io.cpu().numpy(), ref) <= _tol(dtype), (cols, tpb)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_long_horizon_extreme_contract(dtype) -> None:
    """365 steps, 200 % vol, 10 years: the float32 kernel must hold 1e-5 where a float32 running
    product would not (SURVEY.md §7 hard part 1)."""
    z = philox.normals_matrix(365, 2048, dtype, 3, 0)
    for scheme, flag in ((_cabi.SMC_LOG_EULER, True), (_cabi.SMC_SIMPLE_EULER, False)):
        ref = z.copy()
        ogbm.simulate_paths_inplace(ref, 365, 10.0 / 365, 9000.0, 0.15, -0.1, 2.0, flag)
        io = torch.from_numpy(z.copy()).cuda()
        _cabi.gbm_paths_inplace(io, 10.0 / 365, 9000.0, 0.15, -0.1, 2.0, scheme, 256)
        got = io.cpu().numpy()
        finite = np.isfinite(ref) & (np.abs(ref) > np.finfo(dtype).tiny * 1e3)
        assert rel_elem(got[finite], ref[finite]) <= _tol(dtype), scheme
        assert np.array_equal(np.isinf(got), np.isinf(ref))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("b,n", [(1, 1), (3, 2), (11, 12), (64, 16), (1000, 100), (257, 128), (40, 1024), (3, 5000), (2, 8192)])
def test_cf_fft_mean_shapes(dtype, b, n) -> None:
    """Powers of two (radix-2), arbitrary N (table DFT), N = 1, B = 1."""
    rng = np.random.default_rng(b * 1000 + n)
    mat = rng.random((b, n)).astype(np.float32 if dtype == torch.float32 else np.float64)
    ref = np.fft.fft(mat.astype(np.float64).mean(axis=0))
    cf = _cabi.cf_fft_mean(torch.from_numpy(mat).cuda()).cpu().numpy()
    assert rel_max(cf, ref) <= (2e-6 if dtype == torch.float32 else 1e-13)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("b,n", [(1, 32), (7, 32), (9, 64), (1000, 128), (65, 256), (33, 512), (4096, 128)])
def test_cf_row_fft_method(dtype, b, n) -> None:
    """SMC_CF_ROW_FFT: a register/shuffle FFT per row fused with the batch mean equals
    numpy.mean(numpy.fft.fft(mat, axis=1), axis=0) and the mean-then-FFT method."""
    rng = np.random.default_rng(b + n)
    mat = (rng.random((b, n)) * 10).astype(np.float32 if dtype == torch.float32 else np.float64)
    ref = np.mean(np.fft.fft(mat.astype(np.float64), axis=1), axis=0)
    dev = torch.from_numpy(mat).cuda()
    row = _cabi.cf_fft_mean(dev, _cabi.SMC_CF_ROW_FFT).cpu().numpy()
    lin = _cabi.cf_fft_mean(dev, _cabi.SMC_CF_MEAN_THEN_FFT).cpu().numpy()
    tol = 2e-6 if dtype == torch.float32 else 1e-13
    assert rel_max(row, ref) <= tol and rel_max(row, lin) <= tol


def test_cf_row_fft_on_reference_payoffs() -> None:
    from tests.conftest import load_golden

    g = load_golden("c1_float32_log_euler_raw_paths")  # N = 16 is below the method's range
    with pytest.raises(_cabi.SmcError, match="ROW_FFT"):
        _cabi.cf_fft_mean(torch.from_numpy(g["put_price"].copy()).cuda().view(64, 16), _cabi.SMC_CF_ROW_FFT)
    # same payoffs viewed as 32 rows of 32: the estimate of THAT matrix
    mat = g["put_price"].reshape(32, 32)
    ref = np.mean(np.fft.fft(mat, axis=1), axis=0)
    got = _cabi.cf_fft_mean(torch.from_numpy(mat.copy()).cuda(), _cabi.SMC_CF_ROW_FFT).cpu().numpy()
    assert rel_max(got, ref) <= 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("b,n", [(1, 32), (9, 64), (300, 128), (17, 512), (5, 1), (7, 12), (33, 100), (3, 1024), (2, 8192)])
def test_fft_rows(dtype, b, n) -> None:
    """smc_fft_rows == numpy.fft.fft(mat, axis=1) (shuffle FFT for 32..512 powers of two, table DFT otherwise)."""
    mat = (np.random.default_rng(b * 7 + n).random((b, n)) * 3).astype(np.float32 if dtype == torch.float32 else np.float64)
    got = _cabi.fft_rows(torch.from_numpy(mat).cuda()).cpu().numpy()
    ref = np.fft.fft(mat.astype(np.float64), axis=1)
    assert got.shape == (b, n) and rel_max(got, ref) <= (2e-6 if dtype == torch.float32 else 1e-13)
    with pytest.raises(_cabi.SmcError, match="8192"):
        _cabi.fft_rows(torch.zeros((1, 9000), device="cuda"))


def test_cf_large_network_size_fallback() -> None:
    mat = np.random.default_rng(0).random((2, 9000))
    cf = _cabi.cf_fft_mean(torch.from_numpy(mat).cuda()).cpu().numpy()
    assert rel_max(cf, np.fft.fft(mat.mean(axis=0))) <= 1e-12


def test_full_size_properties() -> None:
    """BASELINE config c2 shape (T=252, N=128) at B=8192 (1M paths, 1 GB of normals):
    size-independent properties — terminal-only == last row of in-place, CF DC bin == N * mean,
    linearity of the CF in the payoff matrix."""
    T, N, B = 252, 128, 8192
    z = torch.empty((T, N * B), dtype=torch.float32, device="cuda")
    _cabi.philox_normals(z, 7, 0)
    term = _cabi.gbm_terminal_from_normals(z, 1.0 / T, 100.0, 0.05, 0.0, 0.2, _cabi.SMC_LOG_EULER)
    _cabi.gbm_paths_inplace(z, 1.0 / T, 100.0, 0.05, 0.0, 0.2, _cabi.SMC_LOG_EULER, 256)
    assert torch.allclose(term, z[-1], rtol=1e-6, atol=0)
    put, call = _cabi.payoff(term, 100.0, float(np.exp(-0.05)))
    cf = _cabi.cf_fft_mean(put.view(B, N))
    assert abs(cf[0].real.item() - N * put.double().mean().item()) <= 1e-6 * N * put.double().mean().item()
    cf2 = _cabi.cf_fft_mean((2.0 * put + call).view(B, N))
    cfc = _cabi.cf_fft_mean(call.view(B, N))
    assert rel_max((2 * cf + cfc).cpu().numpy(), cf2.cpu().numpy()) <= 1e-5
    # put-call parity in expectation: mean(call) - mean(put) = df * (mean(X_T) - K)
    lhs = call.double().mean().item() - put.double().mean().item()
    rhs = float(np.exp(-0.05)) * (term.double().mean().item() - 100.0)
    assert abs(lhs - rhs) <= 1e-5 * abs(rhs) + 1e-6


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_headline_shape_parity_subrun(dtype) -> None:
    """BASELINE config c2's shape (T=252, N=128) at B=512 with injected NumPy normals (the kind of
    draws the golden fixtures hold): every stored path value and the CF against the oracle."""
    T, N, B = 252, 128, 512
    z = np.random.default_rng(2024).standard_normal((T, N * B)).astype(dtype)
    c = ogbm.Contract(100.0, 100.0, 1.0, 0.05, 0.0, 0.2)
    for norm in (ogbm.RAW, ogbm.NORMALIZE):
        sr = ogbm.simulate(c, z.copy(), normalization=norm)
        pr = ogbm.price(c, sr)
        ref_cf = ogbm.cf_estimate(pr.put_price, B, N)
        io = torch.from_numpy(z.copy()).cuda()
        _cabi.gbm_paths_inplace(io, 1.0 / T, 100.0, 0.05, 0.0, 0.2, _cabi.SMC_LOG_EULER, 256)
        if norm == ogbm.NORMALIZE:
            _cabi.normalize_rows(io, torch.from_numpy(sr.forwards).cuda())
        assert rel_elem(io.cpu().numpy(), sr.sims) <= _tol(dtype)
        put, _ = _cabi.payoff(io[-1], 100.0, float(sr.df[-1]))
        cf = _cabi.cf_fft_mean(put.view(B, N)).cpu().numpy()
        assert rel_max(cf, ref_cf) <= _tol(dtype)
        if N >= 32:
            cf_row = _cabi.cf_fft_mean(put.view(B, N), _cabi.SMC_CF_ROW_FFT).cpu().numpy()
            assert rel_max(cf_row, ref_cf) <= _tol(dtype)
