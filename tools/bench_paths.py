#!/usr/bin/env python
"""Times the in-place and terminal-only steppers on a (252, 2M) float32 matrix (A/B of library builds)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectralmc_b200 import _cabi
z = torch.empty((252, 128 * 16384), dtype=torch.float32, device="cuda")
res = {"lib": os.environ.get("SMC_LIB", "default")}
for tpb in (128, 256, 512):
    _cabi.philox_normals(z, 7, 0)
    best = 1e9
    for _ in range(4):
        _cabi.philox_normals(z, 7, 0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _cabi.gbm_paths_inplace(z, 1.0 / 252, 100.0, 0.05, 0.0, 0.2, 0, tpb); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    res[f"inplace_tpb{tpb}_GBps"] = round(2 * z.numel() * 4 / best / 1e6)
_cabi.philox_normals(z, 7, 0)
best = 1e9
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); _cabi.gbm_terminal_from_normals(z, 1.0 / 252, 100.0, 0.05, 0.0, 0.2, 0); b.record(); b.synchronize()
    best = min(best, a.elapsed_time(b))
res["terminal_GBps"] = round(z.numel() * 4 / best / 1e6)
print(json.dumps(res))
