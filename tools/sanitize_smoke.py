#!/usr/bin/env python
"""Tiny run of every kernel family, for `compute-sanitizer --tool memcheck` (one tool per gpurun call).
Sizes are ragged on purpose (odd N, P not a multiple of the block, T % 6 != 0, unaligned views)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectralmc_b200 import _cabi

dev = torch.device("cuda", 0)
rows = torch.tensor([(100.0, 100.0, 1.0, 0.05, 0.0, 0.2), (37.5, 41.0, 2.5, -0.01, 0.03, 0.65)], dtype=torch.float64, device=dev)
for dtype in (torch.float32, torch.float64):
    for (T, N, B) in ((7, 12, 11), (12, 16, 64), (3, 300, 5), (1, 1, 3), (13, 128, 9)):
        for scheme in (_cabi.SMC_LOG_EULER, _cabi.SMC_SIMPLE_EULER, _cabi.SMC_LOG_EULER_STEPWISE):
            for norm in (_cabi.SMC_RAW, _cabi.SMC_NORMALIZE):
                args = _cabi.make_fused_args(rows, 2, T, N, B, dtype, scheme, norm, 42, 3)
                _cabi.cf_fused(args, dev, dtype)
        sh = _cabi.make_fused_args(rows, 2, T, N, B, dtype, 0, _cabi.SMC_NORMALIZE, 42, 3, batch_begin=1 if B > 1 else 0, batch_end=B)
        term, tsum = _cabi.fused_terminal(sh, dev, dtype)
        _cabi.cf_from_terminal(sh, term, tsum, dtype)
    for (r, c) in ((13, 1001), (6, 8), (5, 3), (24, 260)):
        z = torch.empty((r, c), dtype=dtype, device=dev)
        _cabi.philox_normals(z, 7, 1)
        base = torch.empty(r * c + 1, dtype=dtype, device=dev)
        _cabi.philox_normals(base[1:].view(r, c), 7, 1)
        t = _cabi.gbm_terminal_from_normals(z, 0.1, 50.0, 0.02, 0.01, 0.4, 0)
        for tpb in (32, 1024):
            _cabi.gbm_paths_inplace(z.clone(), 0.1, 50.0, 0.02, 0.01, 0.4, 1, tpb)
        _cabi.normalize_rows(z, torch.ones(r, dtype=dtype, device=dev))
        put, call = _cabi.payoff(t, 50.0, 0.9)
        _cabi.means3(t, put, call)
    for (b, n) in ((11, 12), (9, 64), (3, 5000), (2, 9000), (33, 512)):
        m = torch.rand((b, n), dtype=dtype, device=dev)
        _cabi.cf_fft_mean(m)
        if n in (64, 512):
            _cabi.cf_fft_mean(m, _cabi.SMC_CF_ROW_FFT)
        if n <= 8192:
            _cabi.fft_rows(m)
torch.cuda.synchronize()
print("sanitize smoke done")
