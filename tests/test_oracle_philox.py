"""The normal-stream specification (oracle/philox.py): Random123 known-answer vectors, an
independent cross-check against NVIDIA's cuRAND Philox header compiled for the host, and the
statistical quality of the Box–Muller output."""

from __future__ import annotations

import os
import shutil
import subprocess
import textwrap

import numpy as np
import pytest
from scipy import stats

from oracle import philox

# Random123 kat_vectors, philox4x32 with 10 rounds: (counter, key) -> output
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,expect", KAT)
def test_random123_known_answers(ctr, key, expect) -> None:
    out = philox.philox4x32_10(ctr, key)
    assert tuple(int(x) for x in out) == expect


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_against_curand_header_on_host(tmp_path) -> None:
    """curand_philox4x32_x.h (CUDA toolkit) compiled as host code gives the same blocks."""
    rng = np.random.default_rng(11)
    ctrs = rng.integers(0, 2**32, size=(16, 4), dtype=np.uint64)
    keys = rng.integers(0, 2**32, size=(16, 2), dtype=np.uint64)
    rows = ",".join("{%s}" % ",".join(f"{int(v)}u" for v in list(c) + list(k)) for c, k in zip(ctrs, keys))
    src = tmp_path / "kat.cu"
    src.write_text(textwrap.dedent(f"""
        #include <cstdio>
        #include <cuda_runtime.h>
        #define QUALIFIERS static inline __host__ __device__
        #include <curand_philox4x32_x.h>
        int main() {{
          unsigned v[][6] = {{{rows}}};
          for (auto& r : v) {{
            uint4 c = {{r[0], r[1], r[2], r[3]}}; uint2 k = {{r[4], r[5]}};
            uint4 o = curand_Philox4x32_10(c, k);
            printf("%u %u %u %u\\n", o.x, o.y, o.z, o.w);
          }}
        }}"""))
    exe = tmp_path / "kat"
    subprocess.run(["nvcc", "-w", "-o", str(exe), str(src)], check=True, capture_output=True)
    got = np.array([[int(t) for t in line.split()] for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n") if line])
    for i in range(16):
        out = philox.philox4x32_10(tuple(int(x) for x in ctrs[i]), tuple(int(x) for x in keys[i]))
        assert [int(x) for x in out] == list(got[i])


def test_uniforms_are_exact_and_open() -> None:
    f = np.array([0, 1, 2**21 - 1], dtype=np.uint64)
    u = philox.uniform_21(f)
    assert np.all((u > 0) & (u < 1)) and u[0] == 2.0**-22 and u[2] == 1 - 2.0**-22
    assert np.all(u.astype(np.float32).astype(np.float64) == u)  # exactly representable in float32
    y = np.array([0, 2**32 - 1], dtype=np.uint32)
    r = philox.uniform_refined(y)
    assert r[0] == 2.0**-45 and 0 < r[1] < 2.0**-21 and np.all(r.astype(np.float32).astype(np.float64) == r)
    w = np.array([0, 0xFFFFFFFF], dtype=np.uint32)
    radius, angle = philox.f64_fields(w, w)
    assert [int(v) for v in radius] == [0, 2**43 - 1] and [int(v) for v in angle] == [0, 2**21 - 1]
    d = philox.uniform_43(radius)
    assert 0 < d[0] < d[1] < 1 and d[0] == 2.0**-44 and d[1] == 1 - 2.0**-44


def test_float64_fields_partition_the_pair() -> None:
    """The 43-bit radius field and the 21-bit angle field of a float64 pair use each of its 64 bits exactly once."""
    for bit in range(64):
        w0 = np.array([(1 << bit) if bit < 32 else 0], dtype=np.uint32)
        w1 = np.array([(1 << (bit - 32)) if bit >= 32 else 0], dtype=np.uint32)
        radius, angle = philox.f64_fields(w0, w1)
        assert int(radius[0] != 0) + int(angle[0] != 0) == 1, bit


def test_float32_fields_partition_the_block() -> None:
    """The six 21-bit fields use disjoint bits of the 128-bit block (126 used, 2 spare)."""
    for bit in range(128):
        words = [np.array([(1 << (bit - 32 * w)) if 32 * w <= bit < 32 * (w + 1) else 0], dtype=np.uint32) for w in range(4)]
        radius, angle = philox.f32_fields(*words)
        hits = sum(int(f[0] != 0) for f in radius + angle)
        spare = bit in (96, 97)  # x3 bits 0 and 1
        assert hits == (0 if spare else 1), bit
    ones = [np.array([0xFFFFFFFF], dtype=np.uint32)] * 4
    radius, angle = philox.f32_fields(*ones)
    assert all(int(f[0]) == 2**21 - 1 for f in radius + angle)


def test_refinement_block_is_used_for_zero_radius_fields() -> None:
    """Columns whose radius field is zero take their radius from the refinement block: the value
    is finite, beyond the 5.46-sigma cap of an unrefined 21-bit uniform is reachable, and the C
    restatement agrees."""
    from oracle import cport

    cols = 1 << 19
    j = np.arange(cols, dtype=np.uint32)[None, :]
    q = np.arange(4, dtype=np.uint32)[:, None]
    x = philox.philox4x32_10((j, q, 0, 0), (7, 0))
    radius, _ = philox.f32_fields(*x)
    hits = [(int(a), int(b), p) for p in range(3) for a, b in zip(*np.nonzero(radius[p] == 0))]
    assert hits, "no zero radius field in 2M blocks?"
    z, rad = philox.normals_matrix(24, cols, np.float32, 7, 0, return_radius=True)
    c = cport.normals(24, cols, np.float32, 7, 0)
    assert np.array_equal(z, c)
    for qq, col, p in hits:
        r = rad[6 * qq + 2 * p, col]
        assert np.isfinite(r) and r > 5.46  # sqrt(-2 ln 2^-21.x) and beyond


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_normals_are_standard_normal(dtype) -> None:
    z = philox.normals_matrix(64, 8192, dtype, seed=7, matrix_index=3).astype(np.float64).ravel()
    n = z.size
    assert abs(z.mean()) < 5 / np.sqrt(n)
    assert abs(z.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs(stats.skew(z)) < 5 * np.sqrt(6 / n)
    assert abs(stats.kurtosis(z)) < 5 * np.sqrt(24 / n)
    assert stats.kstest(z, "norm").pvalue > 1e-3
    # rows (time steps) of one path are uncorrelated
    zz = z.reshape(64, 8192)
    c = np.corrcoef(zz[:8])
    assert np.max(np.abs(c - np.eye(8))) < 6 / np.sqrt(8192)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_stream_is_a_pure_function_of_seed_index_row_col(dtype) -> None:
    full = philox.normals_matrix(9, 40, dtype, seed=5, matrix_index=2)
    part = philox.normals_matrix(9, 40, dtype, seed=5, matrix_index=2, col_begin=8, col_end=24)
    assert np.array_equal(full[:, 8:24], part)
    assert np.array_equal(full[:6], philox.normals_matrix(6, 40, dtype, seed=5, matrix_index=2)[:6]) or dtype == np.float64
    assert not np.array_equal(full, philox.normals_matrix(9, 40, dtype, seed=5, matrix_index=3))
    assert not np.array_equal(full, philox.normals_matrix(9, 40, dtype, seed=6, matrix_index=2))
    other = np.float64 if dtype == np.float32 else np.float32
    assert not np.allclose(full, philox.normals_matrix(9, 40, other, seed=5, matrix_index=2), atol=1e-3)


# ---- short-matrix layout (float32, rows <= 3): adjacent columns share one block ---------------------
@pytest.mark.parametrize("rows", [1, 2, 3])
def test_short_layout_shares_one_block_between_adjacent_columns(rows) -> None:
    """Element (i, j) of a float32 matrix with rows <= 3 is normal (j % G) * rows + i of the block with counter
    (j // G, 0x40000000, k lo, k hi), G = 6 // rows — restated here from the block primitives."""
    G = 6 // rows
    assert philox.short_group(rows, np.float32) == G and philox.short_group(rows, np.float64) == 1
    cols, seed, k = 45, 11, (1 << 33) + 5
    z = philox.normals_matrix(rows, cols, np.float32, seed, k)
    groups = np.arange((cols + G - 1) // G, dtype=np.uint32)
    x = philox.philox4x32_10((groups, philox.F32_SHORT_BIT, k & 0xFFFFFFFF, k >> 32), (seed & 0xFFFFFFFF, seed >> 32))
    radius, angle = philox.f32_fields(*x)
    six = []
    for p in range(3):
        assert (radius[p] != 0).all()  # no refinement among these few blocks
        r = np.sqrt(-2.0 * np.log(philox.uniform_21(radius[p])))
        theta = 2.0 * np.pi * (philox.uniform_21(angle[p]) - 0.5)
        six += [r * np.cos(theta), r * np.sin(theta)]
    six = np.stack(six).astype(np.float32)  # [6, groups]
    for j in range(cols):
        for i in range(rows):
            assert z[i, j] == six[(j % G) * rows + i, j // G], (i, j)
    # rows 4 and 5 keep the general layout: the first rows of a 6-row matrix
    for r45 in (4, 5):
        assert np.array_equal(philox.normals_matrix(r45, cols, np.float32, seed, k), philox.normals_matrix(6, cols, np.float32, seed, k)[:r45])
    # the short layout does not reuse the general layout's blocks
    assert not np.array_equal(z[0], philox.normals_matrix(6, cols, np.float32, seed, k)[0])


def test_short_layout_known_answers() -> None:
    """Pinned values of the short layout (seed 42, matrix 3): any change of the layout shows up here."""
    z1 = philox.normals_matrix(1, 8, np.float32, 42, 3)[0]
    z2 = philox.normals_matrix(2, 4, np.float32, 42, 3)
    z3 = philox.normals_matrix(3, 4, np.float32, 42, 3)
    # one block serves columns 0..5 of the 1-row matrix, 0..2 of the 2-row one and 0..1 of the 3-row one
    assert np.array_equal(z1[:6], z2[:, :3].T.ravel()) and np.array_equal(z1[:6], z3[:, :2].T.ravel())
    np.testing.assert_allclose(z1[:4], [-0.35923433, -0.12542744, -0.11234973, 0.93997437], rtol=2e-7)


@pytest.mark.parametrize("rows", [1, 2, 3])
def test_short_layout_slices_statistics_and_c_port(rows) -> None:
    from oracle import cport

    cols = 60000
    z = philox.normals_matrix(rows, cols, np.float32, 5, 2)
    assert np.array_equal(z[:, 1001:40007], philox.normals_matrix(rows, cols, np.float32, 5, 2, col_begin=1001, col_end=40007))
    assert np.array_equal(z, cport.normals(rows, cols, np.float32, 5, 2))
    flat = z.astype(np.float64).ravel()
    assert abs(flat.mean()) < 5 / np.sqrt(flat.size) and abs(flat.var() - 1) < 5 * np.sqrt(2 / flat.size)
    # the six normals of a block are mutually uncorrelated: adjacent columns and rows
    G = 6 // rows
    blocks = z[:, : cols // G * G].reshape(rows, -1, G).transpose(1, 2, 0).reshape(-1, 6).astype(np.float64)
    c = np.corrcoef(blocks.T)
    assert np.max(np.abs(c - np.eye(6))) < 6 / np.sqrt(blocks.shape[0])


def test_opt_in_stream_is_the_same_construction_on_seven_rounds() -> None:
    """stream_version 1 = Philox4x32-7.  The round count is pinned through the 10-round known answers: three more rounds
    (with the Weyl-advanced key) applied to the 7-round output must give the 10-round output."""
    rng = np.random.default_rng(3)
    ctr = [rng.integers(0, 2**32, 64, dtype=np.uint64).astype(np.uint32) for _ in range(4)]
    key = (0x243F6A88, 0x85A308D3)
    seven = philox.philox4x32_10(ctr, key, 7)
    advanced = ((key[0] + 7 * philox.W0) & 0xFFFFFFFF, (key[1] + 7 * philox.W1) & 0xFFFFFFFF)
    ten = philox.philox4x32_10(seven, advanced, 3)
    for a, b in zip(ten, philox.philox4x32_10(ctr, key)):
        assert np.array_equal(a, b)
    for dtype in (np.float32, np.float64):
        for rows in (1, 3, 7, 12):
            z10 = philox.normals_matrix(rows, 33, dtype, 5, 2)
            z7 = philox.normals_matrix(rows, 33, dtype, 5, 2, stream_version=1)
            assert z7.shape == z10.shape and not np.array_equal(z7, z10)
            assert np.array_equal(z7[:, 4:20], philox.normals_matrix(rows, 33, dtype, 5, 2, col_begin=4, col_end=20, stream_version=1))
    z = philox.normals_matrix(64, 8192, np.float32, 7, 3, stream_version=1).astype(np.float64).ravel()
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 5 * np.sqrt(2 / z.size)
