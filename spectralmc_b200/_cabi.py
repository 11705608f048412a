"""ctypes binding of ``libspectralmc_b200.so`` (declared in ``include/spectralmc_b200.h``).

This is the ONLY route from Python to the device kernels: there is no Numba, CuPy, Triton
or CPU fallback on the path.  If the shared library is missing the import fails loudly.
PyTorch is used purely as the owner of device memory and streams; the wrappers below pass
raw pointers, sizes and the current ``cudaStream_t``.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint64, c_void_p

import torch

# SPECTRALMC_B200_LIB overrides the in-tree location (deployment); a path that does not exist is an error
_LIB_PATH = os.environ.get("SPECTRALMC_B200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "lib", "libspectralmc_b200.so"
)

SMC_F32, SMC_F64 = 0, 1
SMC_LOG_EULER, SMC_SIMPLE_EULER, SMC_LOG_EULER_STEPWISE = 0, 1, 2
SMC_NORMALIZE, SMC_RAW = 0, 1
SMC_CF_MEAN_THEN_FFT, SMC_CF_ROW_FFT = 0, 1
SMC_EINVAL = 1
SMC_STREAM_PHILOX10, SMC_STREAM_PHILOX7 = 0, 1

EXPORTS = (
    "smc_version smc_last_error smc_device_info smc_philox_normals smc_philox_normals_v smc_gbm_paths_inplace "
    "smc_gbm_terminal_from_normals smc_normalize_rows_workspace_bytes smc_normalize_rows smc_payoff "
    "smc_means3_workspace_bytes smc_means3 smc_cf_fft_mean_workspace_bytes smc_cf_fft_mean smc_fft_rows "
    "smc_cf_fused_workspace_bytes smc_cf_fused_launch_count smc_cf_fused smc_fused_terminal_workspace_bytes smc_fused_terminal "
    "smc_cf_from_terminal_workspace_bytes smc_cf_from_terminal smc_cf_fused_host_workspace_bytes "
    "smc_cf_fused_host smc_pipe_calibrate "
    "smc_cvnn_workspace_bytes smc_cvnn_output_width smc_cvnn_forward smc_cvnn_loss_backward smc_adam_step "
    "smc_cvnn_train_step "
    "smc_p2p_buffer_bytes smc_p2p_alloc smc_p2p_open smc_p2p_close smc_p2p_free smc_cf_fused_p2p "
    "smc_p2p_allreduce_sum_f64 smc_cf_from_terminal_p2p smc_cf_fused_p2p_check smc_p2p_status smc_cf_fused_plan smc_diag_stream_fields_f32 smc_diag_stream_lags_f32"
).split()

SMC_LAYER_LINEAR, SMC_LAYER_MODRELU, SMC_LAYER_ZRELU = 0, 1, 2


class SmcError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, code: int, message: str) -> None:
        super().__init__(f"libspectralmc_b200 status {code}: {message}")
        self.code = code
        self.message = message


class FusedArgs(Structure):
    """``smc_fused_args`` (include/spectralmc_b200.h)."""

    _fields_ = [
        ("contracts", c_void_p),
        ("n_contracts", c_int64),
        ("timesteps", c_int64),
        ("network_size", c_int64),
        ("batches_total", c_int64),
        ("batch_begin", c_int64),
        ("batch_end", c_int64),
        ("dtype", c_int),
        ("scheme", c_int),
        ("normalization", c_int),
        ("seed", c_uint64),
        ("first_matrix_index", c_uint64),
        ("stream_version", c_int),
    ]


class P2PGroup(Structure):
    """``smc_p2p_group`` (include/spectralmc_b200.h)."""

    _fields_ = [
        ("rank", c_int),
        ("world", c_int),
        ("buffers", c_void_p * 16),
        ("capacity_contracts", c_int64),
        ("network_size", c_int64),
        ("epoch", ctypes.c_uint32),
        ("timeout_ms", ctypes.c_uint32),
    ]


class CvnnLayer(Structure):
    """``smc_cvnn_layer`` (include/spectralmc_b200.h)."""

    _fields_ = [
        ("kind", c_int),
        ("has_bias", c_int),
        ("in_features", c_int64),
        ("out_features", c_int64),
        ("param_offset", c_int64),
    ]


class CvnnNet(Structure):
    """``smc_cvnn_net``."""

    _fields_ = [
        ("layers", POINTER(CvnnLayer)),
        ("n_layers", c_int),
        ("dtype", c_int),
        ("n_inputs", c_int64),
        ("n_params", c_int64),
    ]


class AdamArgs(Structure):
    """``smc_adam_args``."""

    _fields_ = [("lr", c_double), ("beta1", c_double), ("beta2", c_double), ("eps", c_double)]


def _load() -> ctypes.CDLL:
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C spectralmc_b200/csrc`). spectralmc_b200 has no fallback path."
        )
    lib = ctypes.CDLL(_LIB_PATH)
    lib.smc_version.restype = c_int
    lib.smc_last_error.restype = c_char_p
    for name in EXPORTS:
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        if name.endswith("_workspace_bytes"):
            fn.restype = c_size_t
    lib.smc_normalize_rows_workspace_bytes.argtypes = [c_int64, c_int64]
    lib.smc_means3_workspace_bytes.argtypes = [c_int64]
    lib.smc_cf_fft_mean_workspace_bytes.argtypes = [c_int64, c_int64, c_int]
    for name in (
        "smc_cf_fused_workspace_bytes",
        "smc_fused_terminal_workspace_bytes",
        "smc_cf_from_terminal_workspace_bytes",
        "smc_cf_fused_host_workspace_bytes",
    ):
        getattr(lib, name).argtypes = [POINTER(FusedArgs)]
    lib.smc_cf_fused_launch_count.argtypes = [POINTER(FusedArgs)]
    lib.smc_cf_fused_launch_count.restype = c_int
    lib.smc_device_info.argtypes = [POINTER(c_int)] * 3
    lib.smc_philox_normals.argtypes = [c_void_p, c_int64, c_int64, c_int, c_uint64, c_uint64, c_void_p]
    lib.smc_gbm_paths_inplace.argtypes = [c_void_p, c_int64, c_int64, c_int] + [c_double] * 5 + [c_int, c_int, c_void_p]
    lib.smc_gbm_terminal_from_normals.argtypes = (
        [c_void_p, c_int64, c_int64, c_int] + [c_double] * 5 + [c_int, c_void_p, c_void_p]
    )
    lib.smc_normalize_rows.argtypes = [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smc_payoff.argtypes = [c_void_p, c_int64, c_int, c_double, c_double, c_void_p, c_void_p, c_void_p]
    lib.smc_means3.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smc_cf_fft_mean.argtypes = [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smc_fft_rows.argtypes = [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]
    lib.smc_cf_fused.argtypes = [POINTER(FusedArgs), c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smc_fused_terminal.argtypes = [POINTER(FusedArgs), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smc_cf_from_terminal.argtypes = [POINTER(FusedArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smc_cf_fused_host.argtypes = [POINTER(FusedArgs), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smc_pipe_calibrate.argtypes = [c_int, c_int64, POINTER(c_double), c_void_p, c_void_p]
    lib.smc_p2p_buffer_bytes.argtypes = [c_int64, c_int64, c_int]
    lib.smc_p2p_buffer_bytes.restype = c_size_t
    lib.smc_p2p_alloc.argtypes = [c_size_t, POINTER(c_void_p), c_void_p]
    lib.smc_p2p_open.argtypes = [c_void_p, POINTER(c_void_p)]
    lib.smc_p2p_close.argtypes = [c_void_p]
    lib.smc_p2p_free.argtypes = [c_void_p]
    lib.smc_cf_fused_p2p.argtypes = [POINTER(FusedArgs), POINTER(P2PGroup), c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smc_p2p_allreduce_sum_f64.argtypes = [c_void_p, c_int64, POINTER(P2PGroup), c_void_p]
    lib.smc_diag_stream_fields_f32.argtypes = [c_uint64, c_uint64, c_uint64, ctypes.c_uint32, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    lib.smc_diag_stream_lags_f32.argtypes = [c_uint64, c_uint64, ctypes.c_uint32, ctypes.c_uint32, c_void_p, c_int, c_void_p]
    lib.smc_philox_normals_v.argtypes = [c_void_p, c_int64, c_int64, c_int, c_uint64, c_uint64, c_int, c_void_p]
    lib.smc_cf_fused_plan.argtypes = [POINTER(FusedArgs), POINTER(c_int64), c_int]
    lib.smc_cf_fused_p2p_check.argtypes = [POINTER(FusedArgs), POINTER(P2PGroup), c_size_t]
    lib.smc_p2p_status.argtypes = [POINTER(P2PGroup), POINTER(ctypes.c_uint32), c_void_p]
    lib.smc_cf_from_terminal_p2p.argtypes = [POINTER(FusedArgs), POINTER(P2PGroup), c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smc_cvnn_workspace_bytes.argtypes = [POINTER(CvnnNet), c_int64, c_int]
    lib.smc_cvnn_output_width.argtypes = [POINTER(CvnnNet)]
    lib.smc_cvnn_output_width.restype = c_int64
    lib.smc_cvnn_forward.argtypes = [POINTER(CvnnNet)] + [c_void_p] * 3 + [c_int64] + [c_void_p] * 3 + [c_size_t, c_void_p]
    lib.smc_cvnn_loss_backward.argtypes = (
        [POINTER(CvnnNet)] + [c_void_p] * 4 + [c_int64] + [c_void_p] * 3 + [c_size_t, c_void_p]
    )
    lib.smc_adam_step.argtypes = [c_void_p] * 4 + [c_int64, c_int, c_void_p, POINTER(AdamArgs), c_void_p]
    lib.smc_cvnn_train_step.argtypes = (
        [POINTER(CvnnNet)] + [c_void_p] * 5 + [POINTER(AdamArgs)] + [c_void_p] * 3 + [c_int64] + [c_void_p] * 2 + [c_size_t, c_void_p]
    )
    return lib


LIB = _load()
LIB_PATH = _LIB_PATH


def check(status: int) -> None:
    if status != 0:
        raise SmcError(status, LIB.smc_last_error().decode())


def version() -> int:
    return int(LIB.smc_version())


def device_info() -> tuple[int, int, int]:
    a, b, c = c_int(), c_int(), c_int()
    check(LIB.smc_device_info(byref(a), byref(b), byref(c)))
    return a.value, b.value, c.value


# --------------------------------------------------------------------------- helpers
def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return SMC_F32
    if dtype == torch.float64:
        return SMC_F64
    raise TypeError(f"unsupported dtype {dtype}")


def complex_dtype(dtype: torch.dtype) -> torch.dtype:
    return torch.complex64 if dtype == torch.float32 else torch.complex128


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (spectralmc_b200 has no CPU path)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be C-contiguous")


def stream_handle(stream: object | None) -> int:
    """``cudaStream_t`` (as an int) of whatever stream object the caller holds: ``None`` = torch's current
    stream, a ``torch.cuda.Stream`` (``.cuda_stream``), a Numba stream (``.handle``, a ctypes pointer —
    what the reference passes at gbm.py:313-314,416), a CuPy stream (``.ptr``, gbm.py:313) or a raw integer."""
    if stream is None:
        return _stream()
    if isinstance(stream, int):
        return stream
    for attr in ("cuda_stream", "ptr", "handle"):
        if hasattr(stream, attr):
            h = getattr(stream, attr)
            h = getattr(h, "value", h)  # ctypes.c_void_p -> int | None
            return int(h) if h is not None else 0
    raise TypeError(f"cannot take a CUDA stream handle from {type(stream).__name__}")


_TYPESTR = {"<f4": torch.float32, "<f8": torch.float64, "=f4": torch.float32, "=f8": torch.float64}


def device_matrix(obj: object, name: str = "io") -> tuple[int, tuple[int, ...], torch.dtype, object]:
    """(device pointer, shape, dtype, keep-alive) of a C-contiguous float32/float64 device array given as a
    torch tensor, any object with ``__cuda_array_interface__`` (Numba device arrays — the reference passes
    ``cuda.as_cuda_array(sims)``, gbm.py:418 — and CuPy arrays) or a DLPack exporter.  Zero-copy in every case."""
    if isinstance(obj, torch.Tensor):
        _require_cuda(obj, name)
        dtype_code(obj.dtype)
        return obj.data_ptr(), tuple(obj.shape), obj.dtype, obj
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        shape = tuple(int(x) for x in cai["shape"])
        dtype = _TYPESTR.get(cai["typestr"])
        if dtype is None:
            raise TypeError(f"{name}: unsupported element type {cai['typestr']} (float32 / float64 only)")
        ptr, readonly = cai["data"]
        if readonly:
            raise ValueError(f"{name}: the device array is read-only")
        strides = cai.get("strides")
        if strides is not None:
            expect, acc = [], dtype.itemsize
            for dim in reversed(shape):
                expect.append(acc)
                acc *= dim
            if tuple(strides) != tuple(reversed(expect)):
                raise ValueError(f"{name} must be C-contiguous")
        if ptr is None or (ptr == 0 and all(shape)):
            raise ValueError(f"{name}: NULL device pointer")
        return int(ptr), shape, dtype, obj
    if hasattr(obj, "__dlpack__"):
        t = torch.from_dlpack(obj)
        return device_matrix(t, name)[:3] + (t,)
    raise TypeError(f"{name} must be a CUDA tensor or expose __cuda_array_interface__ / __dlpack__; got {type(obj).__name__}")


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------- wrappers
def philox_normals(out: torch.Tensor, seed: int, matrix_index: int, stream_version: int = SMC_STREAM_PHILOX10) -> torch.Tensor:
    """K1: fill ``out`` (rows, cols) with the ``matrix_index``-th matrix of stream ``seed`` (``stream_version``:
    Philox4x32-10 by default, Philox4x32-7 as an explicit opt-in)."""
    _require_cuda(out, "out")
    rows, cols = out.shape
    check(LIB.smc_philox_normals_v(out.data_ptr(), rows, cols, dtype_code(out.dtype), seed, matrix_index, stream_version, _stream()))
    return out


def gbm_paths_inplace(
    io: object, dt: float, X0: float, r: float, d: float, v: float, scheme: int, threads_per_block: int = 256,
    stream: object | None = None,
) -> None:
    """K2: the reference's ``SimulateBlackScholes`` launch on a materialised matrix.  ``io`` is any device
    array ``device_matrix`` accepts; ``stream`` anything ``stream_handle`` accepts (default: torch's current)."""
    ptr, shape, dtype, _keep = device_matrix(io, "io")
    if len(shape) != 2:
        raise ValueError(f"io must be a (timesteps, paths) matrix; got shape {shape}")
    check(
        LIB.smc_gbm_paths_inplace(
            ptr, shape[0], shape[1], dtype_code(dtype), dt, X0, r, d, v, scheme, threads_per_block, stream_handle(stream)
        )
    )


def gbm_terminal_from_normals(
    normals: torch.Tensor, dt: float, X0: float, r: float, d: float, v: float, scheme: int
) -> torch.Tensor:
    _require_cuda(normals, "normals")
    rows, cols = normals.shape
    out = torch.empty(cols, dtype=normals.dtype, device=normals.device)
    check(
        LIB.smc_gbm_terminal_from_normals(
            normals.data_ptr(), rows, cols, dtype_code(normals.dtype), dt, X0, r, d, v, scheme, out.data_ptr(), _stream()
        )
    )
    return out


def normalize_rows(sims: torch.Tensor, forwards: torch.Tensor) -> None:
    _require_cuda(sims, "sims")
    _require_cuda(forwards, "forwards")
    rows, cols = sims.shape
    if forwards.dtype != sims.dtype or forwards.numel() != rows:
        raise ValueError("forwards must have one entry per row, in the dtype of sims")
    ws = _workspace(LIB.smc_normalize_rows_workspace_bytes(rows, cols), sims.device)
    check(
        LIB.smc_normalize_rows(
            sims.data_ptr(), rows, cols, dtype_code(sims.dtype), forwards.data_ptr(), ws.data_ptr(), ws.numel(), _stream()
        )
    )


def payoff(terminal: torch.Tensor, K: float, df: float) -> tuple[torch.Tensor, torch.Tensor]:
    _require_cuda(terminal, "terminal")
    put = torch.empty_like(terminal)
    call = torch.empty_like(terminal)
    check(
        LIB.smc_payoff(
            terminal.data_ptr(), terminal.numel(), dtype_code(terminal.dtype), K, df, put.data_ptr(), call.data_ptr(), _stream()
        )
    )
    return put, call


def means3(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """Means of three equally long vectors -> device float64[3] (fixed-order reduction)."""
    for t, n in ((a, "a"), (b, "b"), (c, "c")):
        _require_cuda(t, n)
    n = a.numel()
    out = torch.empty(3, dtype=torch.float64, device=a.device)
    ws = _workspace(LIB.smc_means3_workspace_bytes(n), a.device)
    check(
        LIB.smc_means3(
            a.data_ptr(), b.data_ptr(), c.data_ptr(), n, dtype_code(a.dtype), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()
        )
    )
    return out


def cf_fft_mean(mat: torch.Tensor, method: int = SMC_CF_MEAN_THEN_FFT) -> torch.Tensor:
    """K7+K8: ``mean(fft(mat, axis=1), axis=0)`` of a (B, N) real matrix -> (N,) complex."""
    _require_cuda(mat, "mat")
    B, N = mat.shape
    out = torch.empty(N, dtype=complex_dtype(mat.dtype), device=mat.device)
    ws = _workspace(LIB.smc_cf_fft_mean_workspace_bytes(B, N, method), mat.device)
    check(
        LIB.smc_cf_fft_mean(
            mat.data_ptr(), B, N, dtype_code(mat.dtype), method, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()
        )
    )
    return out


def fft_rows(mat: torch.Tensor) -> torch.Tensor:
    """Forward DFT of every row of a real (B, N) matrix -> (B, N) complex (no batch mean)."""
    _require_cuda(mat, "mat")
    B, N = mat.shape
    out = torch.empty((B, N), dtype=complex_dtype(mat.dtype), device=mat.device)
    check(LIB.smc_fft_rows(mat.data_ptr(), B, N, dtype_code(mat.dtype), out.data_ptr(), _stream()))
    return out


def make_fused_args(
    contracts: torch.Tensor | None,
    n_contracts: int,
    timesteps: int,
    network_size: int,
    batches_total: int,
    dtype: torch.dtype,
    scheme: int,
    normalization: int,
    seed: int,
    first_matrix_index: int,
    batch_begin: int = 0,
    batch_end: int | None = None,
    stream_version: int = SMC_STREAM_PHILOX10,
) -> FusedArgs:
    return FusedArgs(
        contracts.data_ptr() if contracts is not None else None,
        n_contracts,
        timesteps,
        network_size,
        batches_total,
        batch_begin,
        batches_total if batch_end is None else batch_end,
        dtype_code(dtype),
        scheme,
        normalization,
        seed,
        first_matrix_index,
        stream_version,
    )


def cf_fused(args: FusedArgs, device: torch.device, dtype: torch.dtype, workspace: torch.Tensor | None = None) -> torch.Tensor:
    """The fused batch path: returns the (partial) CF targets ``[n_contracts, N]`` complex."""
    out = torch.empty((args.n_contracts, args.network_size), dtype=complex_dtype(dtype), device=device)
    need = LIB.smc_cf_fused_workspace_bytes(byref(args))
    ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, device)
    check(LIB.smc_cf_fused(byref(args), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out


def fused_terminal(args: FusedArgs, device: torch.device, dtype: torch.dtype) -> tuple[torch.Tensor, torch.Tensor]:
    paths_local = (args.batch_end - args.batch_begin) * args.network_size
    terminal = torch.empty((args.n_contracts, paths_local), dtype=dtype, device=device)
    tsum = torch.empty(args.n_contracts, dtype=torch.float64, device=device)
    ws = _workspace(LIB.smc_fused_terminal_workspace_bytes(byref(args)), device)
    check(LIB.smc_fused_terminal(byref(args), terminal.data_ptr(), tsum.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return terminal, tsum


def cf_from_terminal(
    args: FusedArgs, terminal: torch.Tensor, terminal_sum_global: torch.Tensor | None, dtype: torch.dtype
) -> torch.Tensor:
    _require_cuda(terminal, "terminal")
    out = torch.empty((args.n_contracts, args.network_size), dtype=complex_dtype(dtype), device=terminal.device)
    ws = _workspace(LIB.smc_cf_from_terminal_workspace_bytes(byref(args)), terminal.device)
    tsum_ptr = terminal_sum_global.data_ptr() if terminal_sum_global is not None else None
    check(
        LIB.smc_cf_from_terminal(
            byref(args), terminal.data_ptr(), tsum_ptr, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()
        )
    )
    return out


def cf_fused_host(args: FusedArgs, contracts_host: torch.Tensor, out_host: torch.Tensor, workspace: torch.Tensor) -> None:
    """Host-buffer entry point: H2D contracts, fused path, D2H targets, stream sync."""
    check(
        LIB.smc_cf_fused_host(
            byref(args), contracts_host.data_ptr(), out_host.data_ptr(), workspace.data_ptr(), workspace.numel(), _stream()
        )
    )


def cf_fused_plan(args: FusedArgs) -> dict:
    """How the simulation of ``args`` is cut into CTAs (``smc_cf_fused_plan``; no device access)."""
    out = (c_int64 * 8)()
    check(LIB.smc_cf_fused_plan(byref(args), out, 8))
    return {"tiles": int(out[0]), "tile_rows": int(out[1]), "row_lanes": int(out[2]), "levels": int(out[3]), "root_fan_in": int(out[4]),
            "main_tiles": int(out[5]), "tail_tile_rows": int(out[6])}


def cf_fused_host_workspace_bytes(args: FusedArgs) -> int:
    return int(LIB.smc_cf_fused_host_workspace_bytes(byref(args)))


def pipe_calibrate(kind: int, iters: int, device: torch.device) -> tuple[float, float]:
    """Run one calibration kernel; returns (lane-ops executed, milliseconds)."""
    sink = torch.zeros(4, dtype=torch.float32, device=device)
    ops = c_double()
    check(LIB.smc_pipe_calibrate(kind, max(iters // 16, 1), byref(ops), sink.data_ptr(), _stream()))  # warm-up
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    check(LIB.smc_pipe_calibrate(kind, iters, byref(ops), sink.data_ptr(), _stream()))
    stop.record()
    stop.synchronize()
    return ops.value, start.elapsed_time(stop)


# --------------------------------------------------------------------------- CVNN step
def make_cvnn_net(layers: list[tuple], n_inputs: int, dtype: torch.dtype) -> tuple[CvnnNet, int]:
    """Descriptor for ``layers`` = [("linear", in, out, bias) | ("modrelu", features) | ("zrelu",)].

    Parameter offsets follow ``module.parameters()`` order with no padding; returns the descriptor
    and the flat parameter count.  The layer array is kept alive as ``net._keep``.
    """
    arr = (CvnnLayer * len(layers))()
    offset = 0
    for slot, layer in zip(arr, layers):
        kind = layer[0]
        slot.param_offset = offset
        if kind == "linear":
            slot.kind, slot.in_features, slot.out_features, slot.has_bias = SMC_LAYER_LINEAR, layer[1], layer[2], int(layer[3])
            offset += 2 * layer[1] * layer[2] + (2 * layer[2] if layer[3] else 0)
        elif kind == "modrelu":
            slot.kind, slot.in_features = SMC_LAYER_MODRELU, layer[1]
            offset += layer[1]
        elif kind == "zrelu":
            slot.kind = SMC_LAYER_ZRELU
        else:
            raise ValueError(f"unsupported CVNN layer {kind!r}")
    net = CvnnNet(arr, len(layers), dtype_code(dtype), n_inputs, offset)
    net._keep = arr
    return net, offset


def cvnn_workspace_bytes(net: CvnnNet, rows: int, training: bool) -> int:
    n = int(LIB.smc_cvnn_workspace_bytes(byref(net), rows, int(training)))
    if n == 0:
        raise SmcError(SMC_EINVAL, LIB.smc_last_error().decode() or "inconsistent CVNN descriptor")
    return n


def cvnn_output_width(net: CvnnNet) -> int:
    w = int(LIB.smc_cvnn_output_width(byref(net)))
    if w < 0:
        raise SmcError(SMC_EINVAL, LIB.smc_last_error().decode())
    return w


def cvnn_forward(net: CvnnNet, params: torch.Tensor, in_r: torch.Tensor, in_i: torch.Tensor, out_r: torch.Tensor,
                 out_i: torch.Tensor, workspace: torch.Tensor) -> None:
    for t, n in ((params, "params"), (in_r, "in_r"), (in_i, "in_i"), (out_r, "out_r"), (out_i, "out_i")):
        _require_cuda(t, n)
    check(
        LIB.smc_cvnn_forward(
            byref(net), params.data_ptr(), in_r.data_ptr(), in_i.data_ptr(), in_r.shape[0], out_r.data_ptr(),
            out_i.data_ptr(), workspace.data_ptr(), workspace.numel(), _stream(),
        )
    )


def cvnn_loss_backward(net: CvnnNet, params: torch.Tensor, in_r: torch.Tensor, in_i: torch.Tensor, targets: torch.Tensor,
                       grads: torch.Tensor, loss: torch.Tensor, workspace: torch.Tensor) -> None:
    for t, n in ((params, "params"), (in_r, "in_r"), (in_i, "in_i"), (targets, "targets"), (grads, "grads"), (loss, "loss")):
        _require_cuda(t, n)
    check(
        LIB.smc_cvnn_loss_backward(
            byref(net), params.data_ptr(), in_r.data_ptr(), in_i.data_ptr(), targets.data_ptr(), in_r.shape[0],
            grads.data_ptr(), loss.data_ptr(), workspace.data_ptr(), workspace.numel(), _stream(),
        )
    )


def adam_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: torch.Tensor,
              hyper: AdamArgs) -> None:
    for t, n in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq"), (step, "step")):
        _require_cuda(t, n)
    check(
        LIB.smc_adam_step(
            params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), params.numel(),
            dtype_code(params.dtype), step.data_ptr(), byref(hyper), _stream(),
        )
    )


def cvnn_train_step(net: CvnnNet, params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
                    step: torch.Tensor, hyper: AdamArgs, in_r: torch.Tensor, in_i: torch.Tensor, targets: torch.Tensor,
                    loss: torch.Tensor, workspace: torch.Tensor) -> None:
    for t, n in ((params, "params"), (in_r, "in_r"), (in_i, "in_i"), (targets, "targets"), (loss, "loss")):
        _require_cuda(t, n)
    check(
        LIB.smc_cvnn_train_step(
            byref(net), params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), step.data_ptr(),
            byref(hyper), in_r.data_ptr(), in_i.data_ptr(), targets.data_ptr(), in_r.shape[0], loss.data_ptr(),
            workspace.data_ptr(), workspace.numel(), _stream(),
        )
    )


# --------------------------------------------------------------------------- stream audit
def diag_stream_fields(seed: int, matrix_index: int, n_blocks: int, cols: int, device: torch.device,
                       stream_version: int = SMC_STREAM_PHILOX10) -> dict:
    """Field histograms, tail counts and power sums of ``6 * n_blocks`` float32-stream normals (no matrix in HBM)."""
    radius = torch.zeros(1 << 21, dtype=torch.int32, device=device)
    angle = torch.zeros(1 << 21, dtype=torch.int32, device=device)
    tails = torch.zeros(4, dtype=torch.int64, device=device)
    sums = torch.zeros(4, dtype=torch.float64, device=device)
    check(LIB.smc_diag_stream_fields_f32(seed, matrix_index, n_blocks, cols, radius.data_ptr(), angle.data_ptr(), tails.data_ptr(),
                                         sums.data_ptr(), stream_version, _stream()))
    return {"radius_hist": radius, "angle_hist": angle, "tails": tails, "power_sums": sums}


def diag_stream_lags(seed: int, matrix_index: int, cols: int, rows: int, device: torch.device,
                     stream_version: int = SMC_STREAM_PHILOX10) -> torch.Tensor:
    sums = torch.zeros(7, dtype=torch.float64, device=device)
    check(LIB.smc_diag_stream_lags_f32(seed, matrix_index, cols, rows, sums.data_ptr(), stream_version, _stream()))
    return sums


# --------------------------------------------------------------------------- peer-memory exchange
def p2p_alloc(nbytes: int) -> tuple[int, bytes]:
    """Allocate a zeroed exchange buffer on the current device; returns (device pointer, 64-byte IPC handle)."""
    ptr = c_void_p()
    handle = ctypes.create_string_buffer(64)
    check(LIB.smc_p2p_alloc(nbytes, byref(ptr), handle))
    return int(ptr.value), handle.raw


def p2p_open(handle: bytes) -> int:
    ptr = c_void_p()
    check(LIB.smc_p2p_open(ctypes.create_string_buffer(handle, 64), byref(ptr)))
    return int(ptr.value)


def cf_fused_p2p(args: FusedArgs, group: P2PGroup, device: torch.device, dtype: torch.dtype,
                 workspace: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """Sharded fused batch path with the all-reduce fused into the finalise kernel: COMPLETE targets on every rank.
    ``out`` may be a pinned host tensor (device-addressable under unified addressing): the kernel then writes the
    targets straight to host memory."""
    if out is None:
        out = torch.empty((args.n_contracts, args.network_size), dtype=complex_dtype(dtype), device=device)
    need = LIB.smc_cf_fused_workspace_bytes(byref(args))
    ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, device)
    check(LIB.smc_cf_fused_p2p(byref(args), byref(group), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out


def p2p_allreduce_sum_f64(values: torch.Tensor, group: P2PGroup) -> None:
    """In-place sum over ranks of a float64 device vector through the exchange buffers (no collective call)."""
    _require_cuda(values, "values")
    if values.dtype != torch.float64:
        raise TypeError("values must be float64")
    check(LIB.smc_p2p_allreduce_sum_f64(values.data_ptr(), values.numel(), byref(group), _stream()))


def cf_from_terminal_p2p(args: FusedArgs, group: P2PGroup, terminal: torch.Tensor, terminal_sum_global: torch.Tensor | None,
                         dtype: torch.dtype) -> torch.Tensor:
    _require_cuda(terminal, "terminal")
    out = torch.empty((args.n_contracts, args.network_size), dtype=complex_dtype(dtype), device=terminal.device)
    ws = _workspace(LIB.smc_cf_from_terminal_workspace_bytes(byref(args)), terminal.device)
    tsum_ptr = terminal_sum_global.data_ptr() if terminal_sum_global is not None else None
    check(LIB.smc_cf_from_terminal_p2p(byref(args), byref(group), terminal.data_ptr(), tsum_ptr, out.data_ptr(), ws.data_ptr(),
                                       ws.numel(), _stream()))
    return out
