#!/usr/bin/env python
"""Time smc_cf_fused of ANY build of the library (SMC_LIB=path, default: the in-tree one) without going
through spectralmc_b200._cabi, so that older builds with a different export list can be A/B-ed in one
gpurun call.  Shapes: c2 (default), c3 (trainer-test size), c4, c2x8, or C,T,N,B,dtype via SMC_SHAPE.

    SMC_LIB=tools/tune/lib_r1.so python tools/bench_raw.py c2 c3
"""
import ctypes, json, os, sys
from ctypes import POINTER, Structure, byref, c_int, c_int64, c_size_t, c_uint64, c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBP = os.environ.get("SMC_LIB") or os.path.join(ROOT, "spectralmc_b200", "lib", "libspectralmc_b200.so")
lib = ctypes.CDLL(LIBP)


class FusedArgs(Structure):
    _fields_ = [("contracts", c_void_p), ("n_contracts", c_int64), ("timesteps", c_int64), ("network_size", c_int64),
                ("batches_total", c_int64), ("batch_begin", c_int64), ("batch_end", c_int64), ("dtype", c_int), ("scheme", c_int),
                ("normalization", c_int), ("seed", c_uint64), ("first_matrix_index", c_uint64), ("stream_version", c_int)]


lib.smc_cf_fused_workspace_bytes.restype = c_size_t
lib.smc_cf_fused_workspace_bytes.argtypes = [POINTER(FusedArgs)]
lib.smc_cf_fused.argtypes = [POINTER(FusedArgs), c_void_p, c_void_p, c_size_t, c_void_p]
lib.smc_last_error.restype = ctypes.c_char_p

SHAPES = {
    "c2": (1, 252, 128, 65536, "f32"), "c2x8": (8, 252, 128, 65536, "f32"), "c3": (1024, 1, 16, 4096, "f32"),
    "c3t2": (1024, 2, 16, 4096, "f32"), "c3t3": (1024, 3, 16, 4096, "f32"), "c3t16": (1024, 16, 16, 4096, "f32"),
    "c4": (512, 365, 256, 4096, "f64"), "c4s": (64, 365, 256, 4096, "f64"), "c2f64": (1, 252, 128, 65536, "f64"),
    "c2s8": (1, 252, 128, 8192, "f32"),  # one rank's share of a strong-scaled c2 on 8 GPUs
}


def run(name: str) -> dict:
    C, T, N, B, dt = SHAPES[name] if name in SHAPES else tuple(int(x) if x.isdigit() else x for x in name.split(","))
    dev = torch.device("cuda", 0)
    rows = torch.tensor([(100.0, 100.0 + 0.01 * i, 1.0, 0.05, 0.0, 0.2) for i in range(C)], dtype=torch.float64, device=dev)
    norm = 0 if os.environ.get("SMC_NORM") == "1" else 1
    a = FusedArgs(rows.data_ptr(), C, T, N, B, 0, B, 0 if dt == "f32" else 1, int(os.environ.get("SMC_SCHEME", "0")), norm, 7, 0,
                  int(os.environ.get("SMC_STREAM", "0")))  # builds older than ABI 102 ignore the trailing field
    ws = torch.empty(lib.smc_cf_fused_workspace_bytes(byref(a)) + 256, dtype=torch.uint8, device=dev)
    out = torch.empty((C, N), dtype=torch.complex64 if dt == "f32" else torch.complex128, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    times = []
    reps = int(os.environ.get("SMC_REPS", "12"))
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.smc_cf_fused(byref(a), out.data_ptr(), ws.data_ptr(), ws.numel(), st)
        e1.record()
        e1.synchronize()
        if rc:
            raise SystemExit(f"{name}: status {rc}: {lib.smc_last_error().decode()}")
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    times.sort()
    ms = times[0]
    return {"lib": os.path.basename(LIBP), "shape": name, "norm": 1 - norm, "ms_min": round(ms, 4), "ms_med": round(times[len(times) // 2], 4),
            "path_steps_per_s": float(f"{C * T * N * B / (ms * 1e-3):.4g}"), "dc": out[0, 0].real.item() / N,
            "env": {k: v for k, v in os.environ.items() if k.startswith("SMC_") and k not in ("SMC_LIB",)}}


if __name__ == "__main__":
    for name in sys.argv[1:] or ["c2"]:
        print(json.dumps(run(name)), flush=True)
