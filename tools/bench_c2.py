#!/usr/bin/env python
"""c2 RAW log-Euler fused timing only (used to A/B library builds via SMC_LIB label)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectralmc_b200 import _cabi
dev = torch.device("cuda", 0)
contracts = torch.tensor([(100.0, 100.0, 1.0, 0.05, 0.0, 0.2)], dtype=torch.float64, device=dev)
norm = _cabi.SMC_NORMALIZE if os.environ.get("SMC_NORM") == "1" else _cabi.SMC_RAW
scheme = int(os.environ.get("SMC_SCHEME", "0"))
args = _cabi.make_fused_args(contracts, 1, 252, 128, 65536, torch.float32, scheme, norm, 7, 0)
ws = torch.empty(_cabi.LIB.smc_cf_fused_workspace_bytes(_cabi.byref(args)) + 256, dtype=torch.uint8, device=dev)
times = []
for i in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = _cabi.cf_fused(args, dev, torch.float32, ws); b.record(); b.synchronize()
    if i >= 2: times.append(a.elapsed_time(b))
print(json.dumps({"lib": os.environ.get("SMC_LIB", "default"), "norm": os.environ.get("SMC_NORM", "0"), "scheme": scheme, "ms_min": min(times), "ms_med": sorted(times)[len(times) // 2], "dc": out[0, 0].real.item() / 128}))
