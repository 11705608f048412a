#!/usr/bin/env python
"""Small driver for ncu: a few launches of one hot-path kernel at BASELINE config c2 size.

    python tools/prof_fused.py fused|fused_stepwise|fused_f64|fused_c3|normals|inplace|terminal [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spectralmc_b200 import _cabi

what = sys.argv[1] if len(sys.argv) > 1 else "fused"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
CANON = (100.0, 100.0, 1.0, 0.05, 0.0, 0.2)
T, N = 252, 128
if what == "fused_c3":  # trainer-test size: 1024 contracts, one timestep (short-path grouped layout)
    import numpy as np
    rows = np.tile(np.asarray(CANON), (1024, 1)); rows[:, 1] = np.linspace(80, 120, 1024)
    contracts = torch.tensor(rows, dtype=torch.float64, device=dev)
    for i in range(reps):
        args = _cabi.make_fused_args(contracts, 1024, 1, 16, 4096, torch.float32, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, 7, i * 1024)
        out = _cabi.cf_fused(args, dev, torch.float32)
    torch.cuda.synchronize()
    print(what, out[0, 0].item())
elif what.startswith("fused"):
    dtype = torch.float64 if what.endswith("f64") else torch.float32
    B = 8192 if dtype == torch.float64 else 65536
    scheme = _cabi.SMC_LOG_EULER_STEPWISE if "stepwise" in what else _cabi.SMC_SIMPLE_EULER if "simple" in what else _cabi.SMC_LOG_EULER
    contracts = torch.tensor([CANON], dtype=torch.float64, device=dev)
    for i in range(reps):
        args = _cabi.make_fused_args(contracts, 1, T, N, B, dtype, scheme, _cabi.SMC_RAW, 7, i)
        out = _cabi.cf_fused(args, dev, dtype)
    torch.cuda.synchronize()
    print(what, out[0, 0].item())
else:
    B = 16384
    z = torch.empty((T, N * B), dtype=torch.float32, device=dev)
    for i in range(reps):
        if what == "normals":
            _cabi.philox_normals(z, 7, i)
        elif what == "inplace":
            _cabi.philox_normals(z, 7, i)
            _cabi.gbm_paths_inplace(z, 1.0 / T, 100.0, 0.05, 0.0, 0.2, _cabi.SMC_LOG_EULER, 256)
        else:
            _cabi.philox_normals(z, 7, i)
            t = _cabi.gbm_terminal_from_normals(z, 1.0 / T, 100.0, 0.05, 0.0, 0.2, _cabi.SMC_LOG_EULER)
    torch.cuda.synchronize()
    print(what, z[0, 0].item())
