"""Oracle (test-only): specification of the counter-based normal stream.

The reference draws each normal matrix with
``cupy.random.default_rng(seed).standard_normal((rows, cols), dtype)``
(/root/reference/src/spectralmc/async_normals.py:214-215).  CuPy's XORWOW bits
are third-party and unpinned by any reference test, so the B200 path defines
its own stream; this file is its normative CPU statement.

Stream definition
-----------------
``normal(seed, k, i, j)`` = element ``(i, j)`` of the ``k``-th matrix ever served
(``k`` = the reference's ``skips`` counter, async_normals.py:394,400-408).

* key      = (seed & 0xffffffff, seed >> 32)
* float32: one Philox4x32-10 block per (column j, row-quad q = i // 4):
  counter = (j, q, k & 0xffffffff, k >> 32); words (x0, x1) give rows 4q, 4q+1 and
  (x2, x3) give rows 4q+2, 4q+3 through one Box–Muller pair each.
* float64: one block per (column j, row-pair q = i // 2): counter as above with the
  top bit of word 3 set (``0x80000000 | k >> 32``) so the two precisions never share a
  block; (x0, x1) -> 52-bit radius uniform, (x2, x3) -> 52-bit angle uniform, one pair.
* uniforms (exactly representable, open interval):
    f32: m = x >> 9;  u = (m + 0.5) * 2**-23, except the radius uniform when m == 0,
         which is refined with the 9 discarded low bits: u = ((x & 0x1ff) + 0.5) * 2**-32
    f64: m = ((hi & 0xfffff) << 32) | lo;  u = (m + 0.5) * 2**-52
* Box–Muller: r = sqrt(-2 ln u1), theta = 2 pi (u2 - 0.5); even row = r cos(theta),
  odd row = r sin(theta).

Philox4x32-10 follows Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"
(SC'11) — multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 /
0xBB67AE85 — and is checked against the Random123 known-answer vectors in
``tests/test_oracle_philox.py``.
"""

from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
F64_STREAM_BIT = 0x80000000


def philox4x32_10(ctr, key):
    """Philox4x32-10 on arrays of counters.

    ``ctr``: 4 broadcastable uint32 arrays (c0, c1, c2, c3); ``key``: 2 ints.
    Returns 4 uint32 arrays.
    """
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint32) for c in ctr])
    c0, c1, c2, c3 = (c.astype(np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0  # 64-bit product, fits (both < 2**32)
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = (
            hi1 ^ c1 ^ np.uint64(k0),
            lo1,
            hi0 ^ c3 ^ np.uint64(k1),
            lo0,
        )
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _key(seed: int) -> tuple[int, int]:
    return seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF


def uniform_f32_radius(x: np.ndarray) -> np.ndarray:
    """Radius uniform for float32 (float64 array holding exactly-representable f32 values)."""
    x = x.astype(np.uint64)
    m = x >> np.uint64(9)
    coarse = (m.astype(np.float64) + 0.5) * 2.0**-23
    fine = ((x & np.uint64(0x1FF)).astype(np.float64) + 0.5) * 2.0**-32
    return np.where(m == 0, fine, coarse)


def uniform_f32_angle(x: np.ndarray) -> np.ndarray:
    m = x.astype(np.uint64) >> np.uint64(9)
    return (m.astype(np.float64) + 0.5) * 2.0**-23


def uniform_f64(hi: np.ndarray, lo: np.ndarray) -> np.ndarray:
    m = ((hi.astype(np.uint64) & np.uint64(0xFFFFF)) << np.uint64(32)) | lo.astype(np.uint64)
    return (m.astype(np.float64) + 0.5) * 2.0**-52


def _box_muller(u1: np.ndarray, u2: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    r = np.sqrt(-2.0 * np.log(u1))
    theta = 2.0 * np.pi * (u2 - 0.5)
    return r * np.cos(theta), r * np.sin(theta), r


def normals_matrix(
    rows: int,
    cols: int,
    dtype,
    seed: int,
    matrix_index: int,
    *,
    col_begin: int = 0,
    col_end: int | None = None,
    return_radius: bool = False,
):
    """The ``matrix_index``-th ``(rows, cols)`` standard-normal matrix of stream ``seed``.

    ``col_begin/col_end`` select a column slice (the stream is a pure function of the
    global column index, so a slice equals the same slice of the full matrix).
    With ``return_radius`` also returns the Box–Muller radius per element (used by the
    tests to form error bounds for the MUFU-based float32 device path).
    """
    dtype = np.dtype(dtype)
    col_end = cols if col_end is None else col_end
    j = np.arange(col_begin, col_end, dtype=np.uint64).astype(np.uint32)[None, :]
    k_lo, k_hi = matrix_index & 0xFFFFFFFF, (matrix_index >> 32) & 0x7FFFFFFF
    key = _key(seed)
    if dtype == np.float32:
        nq = (rows + 3) // 4
        q = np.arange(nq, dtype=np.uint32)[:, None]
        x0, x1, x2, x3 = philox4x32_10((j, q, k_lo, k_hi), key)
        za, zb, ra = _box_muller(uniform_f32_radius(x0), uniform_f32_angle(x1))
        zc, zd, rc = _box_muller(uniform_f32_radius(x2), uniform_f32_angle(x3))
        z = np.stack([za, zb, zc, zd], axis=1).reshape(4 * nq, -1)[:rows]
        rad = np.stack([ra, ra, rc, rc], axis=1).reshape(4 * nq, -1)[:rows]
        z = z.astype(np.float32)
    elif dtype == np.float64:
        nq = (rows + 1) // 2
        q = np.arange(nq, dtype=np.uint32)[:, None]
        x0, x1, x2, x3 = philox4x32_10((j, q, k_lo, F64_STREAM_BIT | k_hi), key)
        za, zb, ra = _box_muller(uniform_f64(x0, x1), uniform_f64(x2, x3))
        z = np.stack([za, zb], axis=1).reshape(2 * nq, -1)[:rows]
        rad = np.stack([ra, ra], axis=1).reshape(2 * nq, -1)[:rows]
    else:
        raise ValueError(f"unsupported dtype {dtype}")
    z = np.ascontiguousarray(z)
    return (z, np.ascontiguousarray(rad)) if return_radius else z
