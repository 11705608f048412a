"""Oracle (test-only): closed-form Black-76, standing in for QuantLib.

Follows /root/reference/src/spectralmc/quantlib.py:19-39, whose arithmetic is
``ql.blackFormula(type, K, fwd, std, df)`` from QuantLib >=1.37,<2.0
(pyproject.binary.toml:73-ish; QuantLib is not installed in this image).  The
published formula: d1 = ln(F/K)/s + s/2, d2 = d1 - s, s = v*sqrt(T);
call = df*(F*Phi(d1) - K*Phi(d2)); put = df*(K*Phi(-d2) - F*Phi(-d1)); s == 0 gives
the discounted intrinsic value.
"""

from __future__ import annotations

import math


def _phi(x: float) -> float:
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def black76(X0: float, K: float, T: float, r: float, d: float, v: float) -> dict[str, float]:
    std = v * math.sqrt(T)  # quantlib.py:21
    df = math.exp(-r * T)  # quantlib.py:22
    fwd = X0 * math.exp((r - d) * T)  # quantlib.py:23
    put_intr = df * max(K - fwd, 0.0)  # quantlib.py:28
    call_intr = df * max(fwd - K, 0.0)  # quantlib.py:29
    if std <= 0.0:
        put, call = put_intr, call_intr
    else:
        d1 = math.log(fwd / K) / std + 0.5 * std
        d2 = d1 - std
        call = df * (fwd * _phi(d1) - K * _phi(d2))
        put = df * (K * _phi(-d2) - fwd * _phi(-d1))
    return dict(
        put_price_intrinsic=put_intr,
        call_price_intrinsic=call_intr,
        underlying=fwd,
        put_convexity=put - put_intr,
        call_convexity=call - call_intr,
        put_price=put,
        call_price=call,
    )
