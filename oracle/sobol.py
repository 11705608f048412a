"""Oracle (test-only): the Sobol contract batch.

Follows /root/reference/src/spectralmc/sobol_sampler.py:187-203 (construction:
``Sobol(d, scramble=True, seed)`` then ``fast_forward(skip)``) and :238-239
(``lower + (upper - lower) * raw``).  The arithmetic is SciPy's (third party,
``scipy>=1.13,<2.0``); bit-exactness means calling the same SciPy routine, which is what
both this oracle and the product's ``SobolSampler`` do.
"""

from __future__ import annotations

import numpy as np
from scipy.stats.qmc import Sobol

# tests/helpers/factories.py:108-115 default Black-Scholes bounds, field order X0,K,T,r,d,v
DEFAULT_BOUNDS = (
    (0.001, 10_000.0),
    (0.001, 20_000.0),
    (0.0, 10.0),
    (-0.20, 0.20),
    (-0.20, 0.20),
    (0.0, 2.0),
)


def sobol_contracts(n: int, seed: int, skip: int = 0, bounds=DEFAULT_BOUNDS) -> np.ndarray:
    lower = np.array([b[0] for b in bounds], dtype=np.float64)
    upper = np.array([b[1] for b in bounds], dtype=np.float64)
    sampler = Sobol(d=len(bounds), scramble=True, seed=seed)
    if skip:
        sampler.fast_forward(skip)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # n not a power of two only warns
        raw = sampler.random(n)
    return lower + (upper - lower) * raw
