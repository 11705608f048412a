#!/usr/bin/env python
"""How far is the fused float32 path from float64 arithmetic on the SAME normals, over the whole
Sobol contract domain (incl. 200 % vol, 10 y)?  Prints the distribution of the norm-wise CF error
and of the per-path terminal-price error.  (Evidence for DESIGN.md; not a test.)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import gbm as ogbm, philox
from oracle.sobol import sobol_contracts
from spectralmc_b200 import _cabi

T, N, B = 24, 64, 256
rows = sobol_contracts(64, seed=31)
contracts = torch.tensor(rows, device="cuda")
for scheme, name in ((_cabi.SMC_LOG_EULER, "log_euler"), (_cabi.SMC_SIMPLE_EULER, "simple_euler")):
    args = _cabi.make_fused_args(contracts, 64, T, N, B, torch.float32, scheme, _cabi.SMC_RAW, 7, 0)
    cf = _cabi.cf_fused(args, contracts.device, torch.float32).cpu().numpy()
    term, _ = _cabi.fused_terminal(args, contracts.device, torch.float32)
    term = term.cpu().numpy().astype(np.float64)
    cf_err, path_err = [], []
    for i, row in enumerate(rows):
        z = philox.normals_matrix(T, N * B, np.float32, 7, i)
        c = ogbm.Contract(*row)
        sr = ogbm.simulate(c, z, scheme=name, normalization=ogbm.RAW)
        pr = ogbm.price(c, sr)
        ref = ogbm.cf_estimate_linear(pr.put_price, B, N)
        if np.max(np.abs(ref)) > 0:
            cf_err.append(np.max(np.abs(cf[i] - ref)) / np.max(np.abs(ref)))
        t = sr.sims[-1].astype(np.float64)
        ok = t > 1e-30
        path_err.append(np.max(np.abs(term[i][ok] - t[ok]) / t[ok]))
    cf_err, path_err = np.array(cf_err), np.array(path_err)
    print(json.dumps({"scheme": name, "contracts": len(rows), "cf_err_median": float(np.median(cf_err)), "cf_err_p90": float(np.quantile(cf_err, 0.9)),
                      "cf_err_max": float(cf_err.max()), "cf_over_1e-5": int((cf_err > 1e-5).sum()),
                      "path_err_median": float(np.median(path_err)), "path_err_max": float(path_err.max())}))
