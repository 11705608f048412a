#!/usr/bin/env python
"""Times smc_philox_normals on a (252, 2M) float32 matrix (2.1 GB) — run with SMC_NORMALS_VEC=1|2|4."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectralmc_b200 import _cabi
for dtype, cols in ((torch.float32, 128 * 16384), (torch.float64, 128 * 8192)):
    z = torch.empty((252, cols), dtype=dtype, device="cuda")
    _cabi.philox_normals(z, 7, 0); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _cabi.philox_normals(z, 7, 1); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    print(json.dumps({"lib": os.environ.get("SMC_LIB", "default"), "vec": os.environ.get("SMC_NORMALS_VEC", "default"), "dtype": str(dtype), "ms": best, "GBps": z.numel() * z.element_size() / best / 1e6}))
