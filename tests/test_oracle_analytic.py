"""Closed-form Black-76 (oracle.black76 and the product's spectralmc_b200.analytic)."""

from __future__ import annotations

import math

import numpy as np
import pytest

from oracle import gbm, philox
from oracle.black76 import black76
from oracle.sobol import sobol_contracts


def test_known_value_and_parity() -> None:
    p = black76(100.0, 100.0, 1.0, 0.05, 0.0, 0.2)
    assert p["call_price"] == pytest.approx(10.450583572185565, rel=1e-12)  # textbook Black-Scholes
    assert p["put_price"] == pytest.approx(5.573526022256971, rel=1e-12)
    fwd, df = 100.0 * math.exp(0.05), math.exp(-0.05)
    assert p["call_price"] - p["put_price"] == pytest.approx(df * (fwd - 100.0), rel=1e-12)


def test_zero_std_returns_intrinsic() -> None:
    for T, v in ((0.0, 0.3), (2.0, 0.0)):
        p = black76(90.0, 100.0, T, 0.03, 0.01, v)
        assert p["put_price"] == p["put_price_intrinsic"] and p["call_price"] == p["call_price_intrinsic"]


def test_oracle_mc_converges_to_black76() -> None:
    """The restated path kernel prices like the closed form (z-score, as tests/test_gbm.py:113-139)."""
    rows = sobol_contracts(8, seed=31)
    zs = []
    for i, row in enumerate(rows):
        c = gbm.Contract(*row)
        vals = []
        for rep in range(8):
            z = philox.normals_matrix(4, 16 * 1024, np.float64, seed=7, matrix_index=8 * i + rep)
            _, pr = gbm.simulate_fft(c, z, 16, normalization=gbm.RAW)
            vals.append(gbm.host_price(pr)["put_price"])
        vals = np.array(vals)
        se = vals.std(ddof=1) / math.sqrt(len(vals))
        ref = black76(*row)["put_price"]
        zs.append(abs(vals.mean() - ref) / se if se > 0 else 0.0 if abs(vals.mean() - ref) < 1e-8 else 9.0)
    assert np.mean(np.array(zs) > 3.5) <= 0.25
