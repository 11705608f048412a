"""Closed-form Black-76 prices — the role ``spectralmc.quantlib.bs_price_quantlib`` plays in the
reference (/root/reference/src/spectralmc/quantlib.py:19-39, a test/validation helper built on
``ql.blackFormula``).  QuantLib is not installed in this image, so the published formula is
evaluated directly; the result model and field meanings are unchanged.
"""

from __future__ import annotations

import math

from spectralmc_b200.gbm import BlackScholes

Inputs = BlackScholes.Inputs
HostPriceResults = BlackScholes.HostPricingResults

__all__ = ["bs_price_analytic", "bs_price_quantlib"]


def _phi(x: float) -> float:
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def bs_price_analytic(inp: Inputs) -> HostPriceResults:
    std = inp.v * math.sqrt(inp.T)
    df = math.exp(-inp.r * inp.T)
    fwd = inp.X0 * math.exp((inp.r - inp.d) * inp.T)
    put_intr = df * max(inp.K - fwd, 0.0)
    call_intr = df * max(fwd - inp.K, 0.0)
    if std > 0.0:
        d1 = math.log(fwd / inp.K) / std + 0.5 * std
        d2 = d1 - std
        call = df * (fwd * _phi(d1) - inp.K * _phi(d2))
        put = df * (inp.K * _phi(-d2) - fwd * _phi(-d1))
    else:
        put, call = put_intr, call_intr
    return HostPriceResults(
        put_price_intrinsic=put_intr,
        call_price_intrinsic=call_intr,
        underlying=fwd,
        put_convexity=put - put_intr,
        call_convexity=call - call_intr,
        put_price=put,
        call_price=call,
    )


bs_price_quantlib = bs_price_analytic  # reference name
