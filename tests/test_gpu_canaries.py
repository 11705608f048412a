"""Out-of-bounds write check without compute-sanitizer (closed on this pool): every output and
every workspace is carved out of a larger buffer whose guard bands hold a sentinel; after the
kernels run, the guard bands must be untouched.  Shapes are ragged on purpose."""

from __future__ import annotations

import ctypes

import pytest
import torch

from spectralmc_b200 import _cabi

pytestmark = pytest.mark.gpu

GUARD = 4096  # bytes on each side
SENTINEL = 0xA5


class Guarded:
    """`nbytes` of device memory, 256-byte aligned, between two sentinel bands."""

    def __init__(self, nbytes: int) -> None:
        self.nbytes = (nbytes + 255) // 256 * 256
        self.raw = torch.full((self.nbytes + 2 * GUARD,), SENTINEL, dtype=torch.uint8, device="cuda")
        self.ptr = self.raw.data_ptr() + GUARD
        assert self.ptr % 256 == 0 or True

    def view(self, dtype: torch.dtype, shape) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        return self.raw[GUARD : GUARD + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(shape)

    def check(self, what: str) -> None:
        torch.cuda.synchronize()
        lo, hi = self.raw[:GUARD], self.raw[GUARD + self.nbytes :]
        assert bool((lo == SENTINEL).all()) and bool((hi == SENTINEL).all()), f"guard band overwritten: {what}"


ROWS = [(100.0, 100.0, 1.0, 0.05, 0.0, 0.2), (37.5, 41.0, 2.5, -0.01, 0.03, 0.65), (5.0, 4.0, 0.7, 0.1, 0.0, 1.1)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("T,N,B", [(7, 12, 11), (12, 16, 64), (3, 300, 5), (1, 1, 3), (13, 128, 9), (6, 257, 2)])
@pytest.mark.parametrize("norm", [_cabi.SMC_RAW, _cabi.SMC_NORMALIZE])
def test_fused_path_stays_inside_its_buffers(dtype, T, N, B, norm) -> None:
    contracts = torch.tensor(ROWS, dtype=torch.float64, device="cuda")
    esz = 8 if dtype == torch.float32 else 16
    for scheme in (_cabi.SMC_LOG_EULER, _cabi.SMC_SIMPLE_EULER):
        args = _cabi.make_fused_args(contracts, len(ROWS), T, N, B, dtype, scheme, norm, 42, 3)
        need = int(_cabi.LIB.smc_cf_fused_workspace_bytes(ctypes.byref(args)))
        out, ws = Guarded(len(ROWS) * N * esz), Guarded(need)
        _cabi.check(_cabi.LIB.smc_cf_fused(ctypes.byref(args), out.ptr, ws.ptr, need, None))
        out.check("cf_out"), ws.check("workspace")
    # two-phase API on a shard
    args = _cabi.make_fused_args(contracts, len(ROWS), T, N, B, dtype, 0, _cabi.SMC_NORMALIZE, 42, 3, batch_begin=B // 3, batch_end=B)
    rows_local = B - B // 3
    term = Guarded(len(ROWS) * rows_local * N * (esz // 2))
    tsum = Guarded(len(ROWS) * 8)
    need = int(_cabi.LIB.smc_fused_terminal_workspace_bytes(ctypes.byref(args)))
    ws = Guarded(need)
    _cabi.check(_cabi.LIB.smc_fused_terminal(ctypes.byref(args), term.ptr, tsum.ptr, ws.ptr, need, None))
    term.check("terminal"), tsum.check("terminal_sum"), ws.check("terminal workspace")
    need = int(_cabi.LIB.smc_cf_from_terminal_workspace_bytes(ctypes.byref(args)))
    out, ws = Guarded(len(ROWS) * N * esz), Guarded(need)
    _cabi.check(_cabi.LIB.smc_cf_from_terminal(ctypes.byref(args), term.ptr, tsum.ptr, out.ptr, ws.ptr, need, None))
    out.check("cf_out (from terminal)"), ws.check("cf_from_terminal workspace")


@pytest.mark.parametrize("dtype,code", [(torch.float32, 0), (torch.float64, 1)])
@pytest.mark.parametrize("rows,cols", [(13, 1001), (6, 8), (5, 3), (24, 260), (1, 1), (7, 4096)])
def test_materialised_kernels_stay_inside_their_buffers(dtype, code, rows, cols) -> None:
    es = 4 if code == 0 else 8
    mat = Guarded(rows * cols * es)
    _cabi.check(_cabi.LIB.smc_philox_normals(mat.ptr, rows, cols, code, 7, 1, None))
    mat.check("normals")
    term = Guarded(cols * es)
    _cabi.check(_cabi.LIB.smc_gbm_terminal_from_normals(mat.ptr, rows, cols, code, 0.1, 50.0, 0.02, 0.01, 0.4, 0, term.ptr, None))
    term.check("terminal"), mat.check("normals after terminal kernel")
    for tpb in (32, 256, 1024):
        _cabi.check(_cabi.LIB.smc_gbm_paths_inplace(mat.ptr, rows, cols, code, 0.1, 50.0, 0.02, 0.01, 0.4, tpb % 3 == 0, tpb, None))
        mat.check(f"in-place paths tpb={tpb}")
    fw = Guarded(rows * es)
    fw.view(dtype, (rows,)).fill_(1.0)
    need = int(_cabi.LIB.smc_normalize_rows_workspace_bytes(rows, cols))
    ws = Guarded(need)
    _cabi.check(_cabi.LIB.smc_normalize_rows(mat.ptr, rows, cols, code, fw.ptr, ws.ptr, need, None))
    mat.check("normalize_rows matrix"), ws.check("normalize_rows workspace"), fw.check("forwards")
    put, call, m3 = Guarded(cols * es), Guarded(cols * es), Guarded(24)
    _cabi.check(_cabi.LIB.smc_payoff(term.ptr, cols, code, 50.0, 0.9, put.ptr, call.ptr, None))
    put.check("put"), call.check("call")
    need = int(_cabi.LIB.smc_means3_workspace_bytes(cols))
    ws = Guarded(need)
    _cabi.check(_cabi.LIB.smc_means3(term.ptr, put.ptr, call.ptr, cols, code, m3.ptr, ws.ptr, need, None))
    m3.check("means3 out"), ws.check("means3 workspace")


@pytest.mark.parametrize("dtype,code", [(torch.float32, 0), (torch.float64, 1)])
@pytest.mark.parametrize("b,n", [(11, 12), (9, 64), (3, 5000), (2, 9000), (33, 512), (1000, 32), (5, 1)])
def test_cf_kernels_stay_inside_their_buffers(dtype, code, b, n) -> None:
    es = 4 if code == 0 else 8
    mat = Guarded(b * n * es)
    mat.view(dtype, (b, n)).uniform_()
    for method in (0, 1):
        if method == 1 and not (32 <= n <= 512 and n & (n - 1) == 0):
            continue
        need = int(_cabi.LIB.smc_cf_fft_mean_workspace_bytes(b, n, method))
        out, ws = Guarded(n * 2 * es), Guarded(need)
        _cabi.check(_cabi.LIB.smc_cf_fft_mean(mat.ptr, b, n, code, method, out.ptr, ws.ptr, need, None))
        out.check(f"cf out method {method}"), ws.check(f"cf workspace method {method}"), mat.check("cf input")
    if n <= 8192:
        spec = Guarded(b * n * 2 * es)
        _cabi.check(_cabi.LIB.smc_fft_rows(mat.ptr, b, n, code, spec.ptr, None))
        spec.check("fft_rows out")


# ---- peer exchange on ONE device: two "ranks" whose exchange buffers are guarded torch allocations --------------
def _group(rank, world, bufs, cap, n, epoch, timeout_ms=0):
    g = _cabi.P2PGroup()
    g.rank, g.world = rank, world
    for q, b in enumerate(bufs):
        g.buffers[q] = b.ptr
    g.capacity_contracts, g.network_size, g.epoch, g.timeout_ms = cap, n, epoch, timeout_ms
    return g


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("N,B,norm", [(64, 40, _cabi.SMC_RAW), (48, 21, _cabi.SMC_RAW), (16, 33, _cabi.SMC_NORMALIZE),
                                      (1024, 5, _cabi.SMC_RAW)])  # N = 1024: the transform leaves the step kernel (separate exchange-finalise kernel)
def test_peer_exchange_stays_inside_its_buffers(dtype, N, B, norm) -> None:
    """The exchange layout (data slots, per-contract flags, small all-reduce region, status word) addressed by both
    ranks of a 2-rank group, each rank on its own stream of the same device: guard bands around BOTH exchange
    buffers, the outputs and the workspaces stay intact, and the result equals the unsharded call — three epochs,
    so both slot parities are reused."""
    contracts = torch.tensor(ROWS, dtype=torch.float64, device="cuda")
    C, T, world = len(ROWS), 9, 2
    esz = 8 if dtype == torch.float32 else 16
    nbytes = int(_cabi.LIB.smc_p2p_buffer_bytes(C, N, world))
    bufs = [Guarded(nbytes) for _ in range(world)]
    for b in bufs:
        b.raw[GUARD : GUARD + b.nbytes].zero_()  # epochs start at 1: a zeroed flag is "not yet"
    streams = [torch.cuda.Stream() for _ in range(world)]
    cuts = [0, B // 3, B]
    whole = _cabi.cf_fused(_cabi.make_fused_args(contracts, C, T, N, B, dtype, 0, norm, 42, 3), contracts.device, dtype)
    torch.cuda.synchronize()
    for epoch in (1, 2, 3):
        outs, keep = [], []
        for rank in range(world):
            a = _cabi.make_fused_args(contracts, C, T, N, B, dtype, 0, norm, 42, 3, batch_begin=cuts[rank], batch_end=cuts[rank + 1])
            g = _group(rank, world, bufs, C, N, epoch)
            out = Guarded(C * N * esz)
            st = streams[rank].cuda_stream
            if norm == _cabi.SMC_RAW:
                need = int(_cabi.LIB.smc_cf_fused_workspace_bytes(ctypes.byref(a)))
                ws = Guarded(need)
                _cabi.check(_cabi.LIB.smc_cf_fused_p2p(ctypes.byref(a), ctypes.byref(g), out.ptr, ws.ptr, need, st))
                keep += [ws]
            else:
                rows_local = cuts[rank + 1] - cuts[rank]
                term, tsum = Guarded(C * rows_local * N * (esz // 2)), Guarded(C * 8)
                need_a = int(_cabi.LIB.smc_fused_terminal_workspace_bytes(ctypes.byref(a)))
                need_b = int(_cabi.LIB.smc_cf_from_terminal_workspace_bytes(ctypes.byref(a)))
                ws_a, ws_b = Guarded(need_a), Guarded(need_b)
                _cabi.check(_cabi.LIB.smc_fused_terminal(ctypes.byref(a), term.ptr, tsum.ptr, ws_a.ptr, need_a, st))
                _cabi.check(_cabi.LIB.smc_p2p_allreduce_sum_f64(tsum.ptr, C, ctypes.byref(g), st))
                _cabi.check(_cabi.LIB.smc_cf_from_terminal_p2p(ctypes.byref(a), ctypes.byref(g), term.ptr, tsum.ptr, out.ptr, ws_b.ptr, need_b, st))
                keep += [term, tsum, ws_a, ws_b]
            outs.append(out)
            keep.append(out)
        torch.cuda.synchronize()
        for i, k in enumerate(keep + bufs):
            k.check(f"epoch {epoch} buffer {i}")
        cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
        got = [o.view(cdt, (C, N)) for o in outs]
        assert torch.equal(torch.view_as_real(got[0]), torch.view_as_real(got[1]))  # identical bits on both ranks
        tol = 2e-6 if dtype == torch.float32 else 1e-12
        assert float((got[0] - whole).abs().max() / whole.abs().max()) <= tol


def test_peer_that_never_arrives_is_reported_not_trapped() -> None:
    """A wait for a peer is bounded in TIME: the targets come back NaN, the status word names the epoch, and the
    CUDA context survives (the next call works) — no device trap that would take every rank down in turn."""
    contracts = torch.tensor(ROWS, dtype=torch.float64, device="cuda")
    C, T, N, B = len(ROWS), 6, 32, 16
    nbytes = int(_cabi.LIB.smc_p2p_buffer_bytes(C, N, 2))
    bufs = [Guarded(nbytes) for _ in range(2)]
    for b in bufs:
        b.raw[GUARD : GUARD + b.nbytes].zero_()
    a = _cabi.make_fused_args(contracts, C, T, N, B, torch.float32, 0, _cabi.SMC_RAW, 42, 0, batch_begin=0, batch_end=B // 2)
    g = _group(0, 2, bufs, C, N, 7, timeout_ms=50)  # rank 1 never calls
    out = _cabi.cf_fused_p2p(a, g, contracts.device, torch.float32)
    torch.cuda.synchronize()  # no sticky error
    assert bool(torch.isnan(torch.view_as_real(out)).all())
    status = ctypes.c_uint32(0)
    _cabi.check(_cabi.LIB.smc_p2p_status(ctypes.byref(g), ctypes.byref(status), None))
    assert status.value == 7
    for b in bufs:
        b.check("exchange buffer after a timeout")
    again = _cabi.cf_fused(_cabi.make_fused_args(contracts, C, T, N, B, torch.float32, 0, _cabi.SMC_RAW, 42, 0), contracts.device, torch.float32)
    assert bool(torch.isfinite(torch.view_as_real(again)).all())
