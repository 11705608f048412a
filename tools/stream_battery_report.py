#!/usr/bin/env python
"""Numbers behind tests/test_gpu_stream_battery.py, for the record (profiles/r2_stream_battery.txt)."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import stats
from spectralmc_b200 import _cabi
from tests.test_gpu_stream_battery import _spec_tail_probability

dev = torch.device("cuda", 0)
n_blocks = (1 << 33) // 3 + 1
torch.cuda.synchronize(); t0 = time.perf_counter()
out = _cabi.diag_stream_fields(20260318, 5, n_blocks, 1 << 22, dev)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"float32 stream audit: {n_blocks} Philox blocks = {3 * n_blocks} draws of each 21-bit field type = {6 * n_blocks} normals in {dt:.2f} s")
dof = (1 << 21) - 1
for which in ("radius_hist", "angle_hist"):
    c = out[which].cpu().numpy().astype(np.float64); e = c.sum() / (1 << 21)
    chi2 = float(np.sum((c - e) ** 2) / e)
    print(f"  chi-square {which:11s}: {chi2:.1f} on {dof} dof, z = {(chi2 - dof) / math.sqrt(2 * dof):+.2f}; cell counts {int(c.min())}..{int(c.max())} around {e:.1f}")
n = 6.0 * n_blocks
for t, got in zip((4.0, 5.0, 5.5, 6.0), out["tails"].cpu().tolist()):
    spec, ideal = _spec_tail_probability(t), 2 * stats.norm.sf(t)
    print(f"  |z| > {t}: observed {got}, specification {n * spec:.1f} (z = {(got - n * spec) / math.sqrt(n * spec):+.2f}), normal law {n * ideal:.1f} (specification / normal law - 1 = {spec / ideal - 1:+.2e})")
s = [float(v) / n for v in out["power_sums"].cpu()]
print(f"  moments: mean {s[0]:+.2e} (sd {1 / math.sqrt(n):.1e}), E z^2 - 1 = {s[1] - 1:+.2e} (sd {math.sqrt(2 / n):.1e}), E z^3 = {s[2]:+.2e} (sd {math.sqrt(15 / n):.1e}), E z^4 - 3 = {s[3] - 3:+.2e} (sd {math.sqrt(96 / n):.1e})")
cols, rows = 1 << 17, 1536
lags = _cabi.diag_stream_lags(99, 1, cols, rows, dev).cpu().numpy()
print(f"serial correlation over {cols} columns x {rows} rows (sd {1 / math.sqrt(cols * rows):.1e}):")
print("  lags 1..6 along a path: " + ", ".join(f"{lags[k] / (cols * (rows - k - 1)):+.2e}" for k in range(6)) + f"; adjacent columns: {lags[6] / ((cols // 32 * 31) * rows):+.2e}")

# float64 stream (two pairs per block: 43-bit radius + 21-bit angle field per pair), audited from its output
for sv, name in ((0, "Philox4x32-10"), (1, "Philox4x32-7 (opt-in)")):
    rows, cols = 64, 1 << 21
    z = torch.empty((rows, cols), dtype=torch.float64, device=dev)
    _cabi.philox_normals(z, 20260318, 9, stream_version=sv)
    n = z.numel()
    print(f"float64 stream audit, {name}: {n} normals ({rows} rows x {cols} columns)")
    m = [float((z**p).mean()) for p in (1, 2, 3, 4)]
    print(f"  moments: mean {m[0]:+.2e} (sd {1 / math.sqrt(n):.1e}), E z^2 - 1 = {m[1] - 1:+.2e} (sd {math.sqrt(2 / n):.1e}), E z^3 = {m[2]:+.2e} (sd {math.sqrt(15 / n):.1e}), E z^4 - 3 = {m[3] - 3:+.2e} (sd {math.sqrt(96 / n):.1e})")
    a = z.abs()
    print("  tails: " + ", ".join(f"|z| > {t}: {int((a > t).sum())} (normal law {n * 2 * stats.norm.sf(t):.1f})" for t in (3.0, 4.0, 5.0)))
    even, odd = z[0::2], z[1::2]
    u1 = torch.exp(-0.5 * (even * even + odd * odd)); u2 = torch.atan2(odd, even) / (2 * math.pi) + 0.5
    pairs, cells = u1.numel(), 4096
    for nm, u in (("radius uniform exp(-r^2/2)", u1), ("angle uniform atan2/2pi+1/2", u2)):
        c = torch.bincount((u * cells).long().clamp_(0, cells - 1).ravel(), minlength=cells).double()
        chi2 = float(((c - pairs / cells) ** 2).sum() / (pairs / cells))
        print(f"  chi-square {nm}: {chi2:.1f} on {cells - 1} dof, z = {(chi2 - (cells - 1)) / math.sqrt(2 * (cells - 1)):+.2f}")
    j = torch.bincount(((u1 * 64).long().clamp_(0, 63) * 64 + (u2 * 64).long().clamp_(0, 63)).ravel(), minlength=4096).double()
    chi2 = float(((j - pairs / 4096) ** 2).sum() / (pairs / 4096))
    print(f"  chi-square joint 64 x 64 (independence of radius and angle): {chi2:.1f} on 4095 dof, z = {(chi2 - 4095) / math.sqrt(2 * 4095):+.2f}")
    print(f"  correlations (sd {1 / math.sqrt(n):.1e}): lags 1..5 along a path " + ", ".join(f"{float((z[l:] * z[:-l]).mean()):+.2e}" for l in range(1, 6)) + f"; adjacent columns {float((z[:, 1:] * z[:, :-1]).mean()):+.2e}")
    del z, a, even, odd, u1, u2
