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
* float32: one Philox4x32-10 block per (column j, row-group q = i // 6):
  counter = (j, q, k & 0xffffffff, k >> 32); its 128 bits feed three Box–Muller pairs
  (rows 6q+2p, 6q+2p+1 for pair p = 0, 1, 2) with 21-bit uniforms:
    radius field  R_p = x_p >> 11
    angle field   A_0 = (x0 & 0x7ff) << 10 | x3 >> 22
                  A_1 = (x1 & 0x7ff) << 10 | (x3 >> 12) & 0x3ff
                  A_2 = (x2 & 0x7ff) << 10 | (x3 >> 2) & 0x3ff
  A zero radius field (probability 2**-21) is refined from word p of a SECOND block with
  counter (j, q | 0x80000000, k lo, k hi): u = ((y_p >> 9) + 0.5) * 2**-44.
* float32, SHORT matrices (rows <= 3; the reference's own tests run ONE timestep,
  /root/reference/tests/test_gbm_trainer.py:127-136, tests/test_gbm.py:49-56): a block's six normals
  would serve a single row of one column, so G = 6 // rows ADJACENT COLUMNS share one block instead:
  counter = (j // G, 0x40000000, k lo, k hi) — the row-group word carries the short-layout bit, so the two
  layouts never share a block — and element (i, j) is normal number (j % G) * rows + i of that block
  (normals 2p, 2p+1 = even / odd value of pair p, as above; refinement counter word 0xC0000000).
  rows = 1: six columns per block; rows = 2: three; rows = 3: two.  rows = 4, 5 keep the general layout.
* float64: one block per (column j, row-group q = i // 4): counter as above with the
  top bit of word 3 set (``0x80000000 | k >> 32``) so the two precisions never share a
  block; its 128 bits feed TWO Box–Muller pairs, 64 bits each — (x0, x1) -> rows 4q, 4q+1 and
  (x2, x3) -> rows 4q+2, 4q+3 — with, for the words (w0, w1) of a pair,
    radius field  R = w0 << 11 | w1 >> 21      (43 bits: the radius reaches 7.7 sigma, no refinement needed)
    angle field   A = w1 & 0x1fffff            (21 bits, as in the float32 stream)
  (since round 2; one pair per block with two 52-bit uniforms before: the integer work of a block bounded the
  float64 kernels — profiles/r2_pipe_overlap_f64_microbench.txt — and 64 random bits per pair are what the
  float32 stream's statistical battery already vouches for at 42).  All ARITHMETIC stays float64.
* uniforms (exactly representable, open interval):
    f32: u = (F + 0.5) * 2**-21 for a 21-bit field F (radius: F = R_p unless R_p == 0, see above)
    f64: u1 = (R + 0.5) * 2**-43,  u2 = (A + 0.5) * 2**-21
* Box–Muller: r = sqrt(-2 ln u1), theta = 2 pi (u2 - 0.5); even row = r cos(theta),
  odd row = r sin(theta).

* ``stream_version`` (include/spectralmc_b200.h): 0 = everything above on Philox4x32-10, the default and the stream every
  published number is measured on; 1 = the identical construction on Philox4x32-7 (the fewest rounds that pass BigCrush
  in Salmon et al.; an explicit opt-in, a different sample set, profiles/r2_philox7_evaluation.md).

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


STREAM_ROUNDS = {0: 10, 1: 7}  # smc_stream_version -> Philox rounds (include/spectralmc_b200.h)


def philox4x32_10(ctr, key, rounds: int = 10):
    """Philox4x32-``rounds`` (10 unless stated) on arrays of counters.

    ``ctr``: 4 broadcastable uint32 arrays (c0, c1, c2, c3); ``key``: 2 ints.
    Returns 4 uint32 arrays.
    """
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint32) for c in ctr])
    c0, c1, c2, c3 = (c.astype(np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(rounds):
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


F32_REFINE_BIT = 0x80000000
F32_SHORT_BIT = 0x40000000  # row-group word of the short-matrix layout (rows <= 3)


def short_group(rows: int, dtype) -> int:
    """Adjacent columns sharing one block: 6 // rows for float32 matrices of at most 3 rows, else 1."""
    return 6 // rows if np.dtype(dtype) == np.float32 and 1 <= rows <= 3 else 1



def f32_fields(x0, x1, x2, x3):
    """21-bit radius and angle fields of the three pairs of a float32 block."""
    x0, x1, x2, x3 = (x.astype(np.uint64) for x in (x0, x1, x2, x3))
    s = np.uint64
    radius = [x0 >> s(11), x1 >> s(11), x2 >> s(11)]
    angle = [
        ((x0 & s(0x7FF)) << s(10)) | (x3 >> s(22)),
        ((x1 & s(0x7FF)) << s(10)) | ((x3 >> s(12)) & s(0x3FF)),
        ((x2 & s(0x7FF)) << s(10)) | ((x3 >> s(2)) & s(0x3FF)),
    ]
    return radius, angle


def uniform_21(field: np.ndarray) -> np.ndarray:
    return (field.astype(np.float64) + 0.5) * 2.0**-21


def uniform_refined(y: np.ndarray) -> np.ndarray:
    return ((y.astype(np.uint64) >> np.uint64(9)).astype(np.float64) + 0.5) * 2.0**-44


def f64_fields(w0: np.ndarray, w1: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """43-bit radius and 21-bit angle fields of one float64 pair (two words of a block)."""
    w0, w1 = w0.astype(np.uint64), w1.astype(np.uint64)
    return (w0 << np.uint64(11)) | (w1 >> np.uint64(21)), w1 & np.uint64(0x1FFFFF)


def uniform_43(field: np.ndarray) -> np.ndarray:
    return (field.astype(np.float64) + 0.5) * 2.0**-43  # exact: 44 significant bits


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
    stream_version: int = 0,
):
    """The ``matrix_index``-th ``(rows, cols)`` standard-normal matrix of stream ``seed``.

    ``col_begin/col_end`` select a column slice (the stream is a pure function of the
    global column index, so a slice equals the same slice of the full matrix).
    With ``return_radius`` also returns the Box–Muller radius per element (used by the
    tests to form error bounds for the MUFU-based float32 device path).
    ``stream_version`` 1 is the opt-in stream: the identical construction on Philox4x32-7.
    """
    dtype = np.dtype(dtype)
    rounds = STREAM_ROUNDS[stream_version]
    col_end = cols if col_end is None else col_end
    j = np.arange(col_begin, col_end, dtype=np.uint64).astype(np.uint32)[None, :]
    k_lo, k_hi = matrix_index & 0xFFFFFFFF, (matrix_index >> 32) & 0x7FFFFFFF
    key = _key(seed)
    if dtype == np.float32:
        G = short_group(rows, dtype)
        if G > 1:  # short layout: block index = column // G, one "row group" carrying the short-layout bit
            g0 = col_begin // G
            j = np.arange(g0, (col_end + G - 1) // G, dtype=np.uint64).astype(np.uint32)[None, :]
            nq = 1
            q = np.full((1, 1), F32_SHORT_BIT, dtype=np.uint32)
        else:
            nq = (rows + 5) // 6
            q = np.arange(nq, dtype=np.uint32)[:, None]
        x = philox4x32_10((j, q, k_lo, k_hi), key, rounds)
        radius, angle = f32_fields(*x)
        u1 = [uniform_21(r) for r in radius]
        need = (radius[0] == 0) | (radius[1] == 0) | (radius[2] == 0)
        if need.any():  # the rare refinement block
            qq, jj = np.nonzero(need)
            y = philox4x32_10((j[0, jj], q[qq, 0] | np.uint32(F32_REFINE_BIT), k_lo, k_hi), key, rounds)
            for p in range(3):
                hit = radius[p][qq, jj] == 0
                u1[p][qq[hit], jj[hit]] = uniform_refined(y[p][hit])
        zs, rs = [], []
        for p in range(3):
            ze, zo, r = _box_muller(u1[p], uniform_21(angle[p]))
            zs += [ze, zo]
            rs += [r, r]
        z = np.stack(zs, axis=1).reshape(6 * nq, -1)
        rad = np.stack(rs, axis=1).reshape(6 * nq, -1)
        if G > 1:  # [6, groups] -> element (i, col) = normal (col % G) * rows + i of group col // G
            cols_here = np.arange(col_begin, col_end)
            grp, lane = cols_here // G - g0, cols_here % G
            pick = lane[None, :] * rows + np.arange(rows)[:, None]
            z, rad = z[pick, grp[None, :]], rad[pick, grp[None, :]]
        else:
            z, rad = z[:rows], rad[:rows]
        z = z.astype(np.float32)
    elif dtype == np.float64:
        nq = (rows + 3) // 4
        q = np.arange(nq, dtype=np.uint32)[:, None]
        x0, x1, x2, x3 = philox4x32_10((j, q, k_lo, F64_STREAM_BIT | k_hi), key, rounds)
        zs, rs = [], []
        for w0, w1 in ((x0, x1), (x2, x3)):  # two pairs per block: rows 4q, 4q+1 and 4q+2, 4q+3
            radius, angle = f64_fields(w0, w1)
            ze, zo, r = _box_muller(uniform_43(radius), uniform_21(angle))
            zs += [ze, zo]
            rs += [r, r]
        z = np.stack(zs, axis=1).reshape(4 * nq, -1)[:rows]
        rad = np.stack(rs, axis=1).reshape(4 * nq, -1)[:rows]
    else:
        raise ValueError(f"unsupported dtype {dtype}")
    z = np.ascontiguousarray(z)
    return (z, np.ascontiguousarray(rad)) if return_radius else z
