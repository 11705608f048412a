"""Oracle (test-only): ctypes access to the C restatement ``oracle/gbm_oracle.c``.

Used by tests as a second, independent checker of ``oracle/gbm.py`` / ``oracle/philox.py`` and by
``bench.py`` as the multi-core CPU baseline.  Never imported by the product.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_int64, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libgbm_oracle.so")


def load(build: bool = True) -> ctypes.CDLL:
    if not os.path.exists(_LIB) and build:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = ctypes.CDLL(_LIB)
    lib.oracle_philox4x32_10.argtypes = [POINTER(c_uint32), POINTER(c_uint32), POINTER(c_uint32)]
    lib.oracle_normals.argtypes = [c_void_p, c_int64, c_int64, c_int, c_uint64, c_uint64]
    lib.oracle_paths_inplace.argtypes = [c_void_p, c_int64, c_int64, c_int] + [c_double] * 5 + [c_int]
    lib.oracle_terminal_range.argtypes = [POINTER(c_double), c_int64, c_int, c_int, c_uint64, c_uint64, c_int64, c_int64, c_void_p]
    lib.oracle_terminal_range.restype = c_double
    lib.oracle_cf_from_terminal.argtypes = [POINTER(c_double), c_int64, c_int64, c_int, c_int, c_void_p, c_double, c_void_p]
    lib.oracle_cf_from_terminal.restype = c_double
    return lib


def philox(ctr, key):
    lib = load()
    c = (c_uint32 * 4)(*ctr)
    k = (c_uint32 * 2)(*key)
    o = (c_uint32 * 4)()
    lib.oracle_philox4x32_10(c, k, o)
    return tuple(int(x) for x in o)


def normals(rows: int, cols: int, dtype, seed: int, matrix_index: int) -> np.ndarray:
    lib = load()
    out = np.empty((rows, cols), dtype=dtype)
    lib.oracle_normals(out.ctypes.data, rows, cols, 0 if out.dtype == np.float32 else 1, seed, matrix_index)
    return out


def paths_inplace(io: np.ndarray, dt, X0, r, d, v, log_flag: bool) -> None:
    lib = load()
    assert io.flags.c_contiguous
    lib.oracle_paths_inplace(io.ctypes.data, io.shape[0], io.shape[1], 0 if io.dtype == np.float32 else 1, dt, X0, r, d, v, int(log_flag))


def simulate_fft(contract, T: int, N: int, B: int, dtype, log_flag: bool, normalize: bool, seed: int, matrix_index: int,
                 threads: int | None = None):
    """One contract end to end on `threads` host threads (ctypes releases the GIL in the C calls).
    Returns (cf[N] complex128, mean put price)."""
    from concurrent.futures import ThreadPoolExecutor

    lib = load()
    threads = threads or os.cpu_count() or 1
    code = 0 if np.dtype(dtype) == np.float32 else 1
    c = (c_double * 6)(*[float(x) for x in contract])
    P = N * B
    terminal = np.empty(P, dtype=np.float64)
    cuts = np.linspace(0, P, min(threads, P) + 1).astype(np.int64)

    def work(i):
        lo, hi = int(cuts[i]), int(cuts[i + 1])
        return lib.oracle_terminal_range(c, T, code, int(log_flag), seed, matrix_index, lo, hi, terminal[lo:].ctypes.data)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        tsum = sum(ex.map(work, range(len(cuts) - 1)))
    out = np.empty(2 * N, dtype=np.float64)
    mean_put = lib.oracle_cf_from_terminal(c, N, B, code, int(normalize), terminal.ctypes.data, tsum, out.ctypes.data)
    return out.view(np.complex128), float(mean_put)
