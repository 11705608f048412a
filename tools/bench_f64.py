#!/usr/bin/env python
"""c2-shape float64 fused timing + float64 generator (used to A/B library builds)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectralmc_b200 import _cabi
dev = torch.device("cuda", 0)
contracts = torch.tensor([(100.0, 100.0, 1.0, 0.05, 0.0, 0.2)], dtype=torch.float64, device=dev)
for scheme, name in ((0, "log"), (1, "simple")):
    args = _cabi.make_fused_args(contracts, 1, 252, 128, 65536, torch.float64, scheme, _cabi.SMC_RAW, 7, 0)
    best = 1e9
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = _cabi.cf_fused(args, dev, torch.float64); b.record(); b.synchronize()
        if i: best = min(best, a.elapsed_time(b))
    print(json.dumps({"lib": os.environ.get("SMC_LIB", "default"), "what": "c2 fp64 fused " + name, "ms": best, "steps_per_s": 252 * 128 * 65536 / best * 1e3, "dc": out[0, 0].real.item() / 128}))
