"""The CPU oracle against outputs of the reference itself (tests/golden/*.npz).

The fixtures were produced by running the reference's own BlackScholes engine
(tests/golden/make_golden.py); this is what pins the oracle before it is trusted as the checker
of the CUDA path.  float32: bit-exact (same NumPy expressions on the same data);
float64: <= 1e-14 relative (math.exp under the simulator vs numpy.exp differ by an ulp).
"""

from __future__ import annotations

import numpy as np

from oracle import gbm
from tests.helpers import rel_max

_FIELDS = ("X0", "K", "T", "r", "d", "v")


def _run(g):
    c = gbm.Contract(*g["contract"])
    sr = gbm.simulate(c, g["normals"].copy(), scheme=str(g["scheme"]), normalization=str(g["normalization"]))
    pr = gbm.price(c, sr)
    cf = gbm.cf_estimate(pr.put_price, int(g["batches"]), int(g["network_size"]))
    return c, sr, pr, cf


def test_oracle_matches_reference_outputs(golden) -> None:
    c, sr, pr, cf = _run(golden)
    f32 = golden["normals"].dtype == np.float32
    tol = 0.0 if f32 else 1e-14
    for name, mine in (("times", sr.times), ("forwards", sr.forwards), ("df", sr.df), ("sims", sr.sims),
                       ("put_price", pr.put_price), ("call_price", pr.call_price), ("underlying", pr.underlying),
                       ("cf", cf)):
        ref = golden[name]
        assert mine.dtype == ref.dtype, name
        assert mine.shape == ref.shape, name
        assert rel_max(mine, ref) <= tol, (golden["name"], name, rel_max(mine, ref))
    host = gbm.host_price(pr)
    assert rel_max(list(host.values()), golden["host"]) <= (0.0 if f32 else 1e-14)


def test_linearity_of_the_cf_estimate(golden) -> None:
    """mean_b FFT_n(mat) == FFT_n(mean_b mat): the identity the CUDA path relies on."""
    _, _, pr, cf = _run(golden)
    lin = gbm.cf_estimate_linear(pr.put_price, int(golden["batches"]), int(golden["network_size"]))
    tol = 2e-6 if pr.put_price.dtype == np.float32 else 1e-13
    assert rel_max(cf, lin) <= tol


def test_dc_bin_carries_the_price(golden) -> None:
    """CF[0] = N * mean(put_price)."""
    _, _, pr, cf = _run(golden)
    n = int(golden["network_size"])
    expect = n * float(np.mean(pr.put_price.astype(np.float64)))
    assert abs(cf[0].real - expect) <= 1e-5 * max(abs(expect), 1.0)
    assert abs(cf[0].imag) <= 1e-12 * max(abs(expect), 1.0)


def test_degenerate_contracts() -> None:
    """T = 0 leaves X = X0 with df = 1; v = 0 follows the forward deterministically."""
    z = np.random.default_rng(1).standard_normal((5, 64))
    sr = gbm.simulate(gbm.Contract(90.0, 100.0, 0.0, 0.03, 0.01, 0.3), z.copy(), normalization=gbm.RAW)
    assert np.all(sr.sims == 90.0) and np.all(sr.df == 1.0)
    c = gbm.Contract(90.0, 100.0, 2.0, 0.03, 0.01, 0.0)
    sr = gbm.simulate(c, z.copy(), normalization=gbm.RAW)
    assert np.allclose(sr.sims[-1], 90.0 * np.exp((0.03 - 0.01) * 2.0), rtol=1e-14)
