#!/usr/bin/env python
"""z-scores of the float64 stream's second moment, |z| > 3 tail count and lag-1..8 correlations over many (seed, matrix) pairs:
a deviation that is chance averages out over seeds, a defect of the construction does not.  Prints the per-statistic mean z-score
times sqrt(runs) (itself a standard normal under the null)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scipy import stats
from spectralmc_b200 import _cabi

dev = torch.device("cuda", 0)
rows, cols = 64, 1 << 21
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sv = int(sys.argv[2]) if len(sys.argv) > 2 else 0
z = torch.empty((rows, cols), dtype=torch.float64, device=dev)
n = z.numel()
tot = [0.0] * 11
for r in range(runs):
    seed, k = 1000003 * (r + 1) + 17, r * 5
    _cabi.philox_normals(z, seed, k, stream_version=sv)
    m2 = (float((z * z).mean()) - 1) / math.sqrt(2 / n)
    m4 = (float((z**4).mean()) - 3) / math.sqrt(96 / n)
    e3 = n * 2 * stats.norm.sf(3.0)
    t3 = (int((z.abs() > 3.0).sum()) - e3) / math.sqrt(e3)
    lags = [float((z[l:] * z[:-l]).mean()) * math.sqrt((rows - l) * cols) for l in range(1, 9)]
    for i, v in enumerate([m2, m4, t3] + lags):
        tot[i] += v
names = ["E z^2", "E z^4", "|z|>3"] + [f"lag {l}" for l in range(1, 9)]
print(f"float64 stream_version {sv}: {runs} runs of {n} normals; combined z-scores (sum / sqrt(runs)):")
print("  " + "  ".join(f"{nm} {t / math.sqrt(runs):+.2f}" for nm, t in zip(names, tot)))
