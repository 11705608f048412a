"""The oracle's C restatement (oracle/gbm_oracle.c through oracle/cport.py) — what `bench.py` times as `cpu_baseline` and as the
`--impl reference` arm — pinned on the reference-run golden vectors and on the NumPy oracle, both precisions.  CPU only."""

from __future__ import annotations

import glob
import os

import numpy as np
import pytest

from oracle import cport, philox
from oracle import gbm as ogbm
from tests.helpers import rel_max

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*_paths.npz")) +
                glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*_forwards.npz")))


@pytest.mark.parametrize("rows", [1, 2, 3, 4, 5, 7, 9, 12])
def test_float64_stream_of_the_c_port_is_the_specification(rows) -> None:
    """Two pairs per block, 43-bit radius and 21-bit angle field per pair; ragged last blocks; libm on both sides."""
    want = philox.normals_matrix(rows, 257, np.float64, 11, 4)
    got = cport.normals(rows, 257, np.float64, 11, 4)
    assert got.shape == want.shape and np.max(np.abs(got - want)) <= 1e-14


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_c_path_kernel_reproduces_the_reference_run(path) -> None:
    """K2 of the C port on the normals the reference drew == the reference's own simulated paths (before NORMALIZE rescales them)."""
    g = np.load(path, allow_pickle=True)
    if str(g["normalization"]).lower().startswith("normalize"):
        pytest.skip("golden sims are rescaled after the path kernel; the kernel itself is pinned by the raw cases")
    X0, _, Tm, r, d, v = (float(x) for x in g["contract"])
    io = np.ascontiguousarray(g["normals"]).copy()
    cport.paths_inplace(io, Tm / int(g["timesteps"]), X0, r, d, v, str(g["scheme"]).lower().startswith("log"))
    tol = 1e-12 if io.dtype == np.float64 else 2e-6
    assert rel_max(io, g["sims"]) <= tol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("scheme", ["log_euler", "simple_euler"])
@pytest.mark.parametrize("norm", ["raw_paths", "normalize_forwards"])
@pytest.mark.parametrize("T,N,B", [(12, 16, 8), (5, 8, 9), (1, 16, 12), (3, 4, 7)])
def test_c_simulate_fft_is_the_numpy_oracle(dtype, scheme, norm, T, N, B) -> None:
    """The whole streaming pipeline of the C port (normals -> paths -> payoff -> CF) == the NumPy oracle on its own statement of
    the same stream, including the short layout (T <= 3) and the float64 two-pairs-per-block layout; threaded == single-threaded."""
    contract = (100.0, 105.0, 0.75, 0.04, 0.01, 0.3)
    z = philox.normals_matrix(T, N * B, dtype, 9, 2)
    want, _ = ogbm.simulate_fft(ogbm.Contract(*contract), z.copy(), N, scheme=scheme, normalization=norm)
    got1, price1 = cport.simulate_fft(contract, T, N, B, dtype, scheme == "log_euler", norm == "normalize_forwards", 9, 2, threads=1)
    got3, price3 = cport.simulate_fft(contract, T, N, B, dtype, scheme == "log_euler", norm == "normalize_forwards", 9, 2, threads=3)
    tol = 1e-12 if dtype == np.float64 else 2e-6
    assert rel_max(got1, np.asarray(want, dtype=np.complex128)) <= tol
    assert rel_max(got3, got1) <= 1e-13 and abs(price3 - price1) <= 1e-12 * max(1.0, abs(price1))
    assert abs(price1 - float(got1[0].real) / N) <= 1e-12 * max(1.0, abs(price1))  # DC bin = N * mean put price
