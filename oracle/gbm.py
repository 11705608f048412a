"""Oracle (test-only): NumPy restatement of the reference's GBM/CF path.

Every function cites the reference lines it follows
(/root/reference/src/spectralmc/...).  Pinned against the reference's own outputs in
``tests/golden/*.npz`` (see ``tests/golden/make_golden.py``).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

LOG_EULER = "log_euler"  # effects/montecarlo.py:27
SIMPLE_EULER = "simple_euler"  # effects/montecarlo.py:28
NORMALIZE = "normalize_forwards"  # effects/montecarlo.py:34
RAW = "raw_paths"  # effects/montecarlo.py:35


@dataclass(frozen=True)
class Contract:
    """BlackScholes.Inputs (gbm.py:267-277): field order X0, K, T, r, d, v."""

    X0: float
    K: float
    T: float
    r: float
    d: float
    v: float

    def __post_init__(self) -> None:
        # Python floats, as the reference's Pydantic model holds them: a NumPy float64
        # scalar would not be a "weak" scalar and would widen float32 engine arithmetic.
        for f in ("X0", "K", "T", "r", "d", "v"):
            object.__setattr__(self, f, float(getattr(self, f)))

    def as_row(self) -> list[float]:
        return [self.X0, self.K, self.T, self.r, self.d, self.v]


def simulate_paths_inplace(
    io: np.ndarray, timesteps: int, dt: float, X0: float, r: float, d: float, v: float, log_flag: bool
) -> None:
    """The path kernel, gbm.py:241-257, all paths at once.

    Arithmetic is float64 for BOTH storage dtypes: Numba types the Python-float scalar
    arguments as float64, so a float32 element is widened on load and the running value
    ``X`` lives in a float64 register and is narrowed only on store (PTX
    opcode evidence recorded in tests/golden/make_golden.py).
    """
    sqrt_dt = np.sqrt(np.float64(dt))  # gbm.py:243
    X = np.full(io.shape[1], X0, dtype=np.float64)  # gbm.py:244
    if log_flag:
        drift = r - d - 0.5 * v * v  # gbm.py:246
        for i in range(timesteps):
            dW = io[i].astype(np.float64) * sqrt_dt  # gbm.py:248
            X = X * np.exp(drift * dt + v * dW)  # gbm.py:249
            io[i] = X  # gbm.py:250 (narrowing store for float32)
    else:
        drift = r - d  # gbm.py:252
        for i in range(timesteps):
            dW = io[i].astype(np.float64) * sqrt_dt  # gbm.py:254
            X = X + (drift * X * dt + v * X * dW)  # gbm.py:255
            X = np.abs(X)  # gbm.py:256
            io[i] = X  # gbm.py:257


@dataclass
class SimResults:
    """gbm.py:279-284."""

    times: np.ndarray
    sims: np.ndarray
    forwards: np.ndarray
    df: np.ndarray


def simulate(
    c: Contract, normals: np.ndarray, *, scheme: str = LOG_EULER, normalization: str = NORMALIZE
) -> SimResults:
    """BlackScholes._simulate, gbm.py:400-447, on an injected normal matrix.

    ``normals`` is consumed in place (as the reference does, gbm.py:405-426) and its
    dtype is the engine dtype.
    """
    timesteps = normals.shape[0]
    dtype = normals.dtype
    dt = c.T / timesteps  # gbm.py:411
    simulate_paths_inplace(normals, timesteps, dt, c.X0, c.r, c.d, c.v, scheme == LOG_EULER)
    sims = normals
    times = np.linspace(dt, c.T, timesteps, dtype=dtype)  # gbm.py:429
    forwards = (c.X0 * np.exp((c.r - c.d) * times)).astype(dtype)  # gbm.py:430 (array dtype)
    df = np.exp(-c.r * times).astype(dtype)  # gbm.py:431
    if normalization == NORMALIZE:
        row_means = np.mean(sims, axis=1, keepdims=True).squeeze()  # gbm.py:437
        sims *= np.expand_dims(forwards / row_means, 1)  # gbm.py:438
    return SimResults(times=times, sims=sims, forwards=forwards, df=df)


@dataclass
class PricingResults:
    """gbm.py:286-292."""

    put_price_intrinsic: np.ndarray
    call_price_intrinsic: np.ndarray
    underlying: np.ndarray
    put_price: np.ndarray
    call_price: np.ndarray


def price(c: Contract, sr: SimResults) -> PricingResults:
    """BlackScholes.price, gbm.py:464-474 (all arithmetic in the engine dtype)."""
    dtype = sr.sims.dtype
    F = sr.forwards[-1]  # gbm.py:465
    df_last = sr.df[-1]  # gbm.py:466
    K = np.asarray(c.K, dtype=dtype)  # gbm.py:467
    put_intr = df_last * np.maximum(K - F, 0)  # gbm.py:469
    call_intr = df_last * np.maximum(F - K, 0)  # gbm.py:470
    terminal = sr.sims[-1]  # gbm.py:472
    put_price = df_last * np.maximum(K - terminal, 0)  # gbm.py:473
    call_price = df_last * np.maximum(terminal - K, 0)  # gbm.py:474
    return PricingResults(put_intr, call_intr, terminal, put_price, call_price)


def host_price(pr: PricingResults) -> dict[str, float]:
    """BlackScholes.get_host_price, gbm.py:491-513."""
    put_intr = float(pr.put_price_intrinsic)
    call_intr = float(pr.call_price_intrinsic)
    underlying = float(pr.underlying.mean())
    put = float(pr.put_price.mean())
    call = float(pr.call_price.mean())
    return dict(
        put_price_intrinsic=put_intr,
        call_price_intrinsic=call_intr,
        underlying=underlying,
        put_convexity=put - put_intr,
        call_convexity=call - call_intr,
        put_price=put,
        call_price=call,
    )


def cf_estimate(put_price: np.ndarray, batches: int, network_size: int) -> np.ndarray:
    """GbmCVNNPricer._simulate_fft, gbm_trainer.py:814-817 (== :409-412).

    ``mat = put_price.reshape(B, N)`` (C order: path = b*N + n), forward unnormalised DFT
    along ``n``, mean over ``b``.  ``numpy.fft.fft`` keeps complex64 for float32 input
    (NumPy >= 2), matching cuFFT's c64 plan.
    """
    mat = put_price.reshape(batches, network_size)
    return np.mean(np.fft.fft(mat, axis=1), axis=0)


def cf_estimate_linear(put_price: np.ndarray, batches: int, network_size: int) -> np.ndarray:
    """Same estimate by linearity, ``FFT_n(mean_b mat)``, in float64 throughout.

    Not a reference code path; used by tests as the high-precision value the two
    formulations both converge to (SURVEY.md key fact 5).
    """
    mat = put_price.reshape(batches, network_size).astype(np.float64)
    return np.fft.fft(mat.mean(axis=0))


def simulate_fft(
    c: Contract,
    normals: np.ndarray,
    network_size: int,
    *,
    scheme: str = LOG_EULER,
    normalization: str = NORMALIZE,
) -> tuple[np.ndarray, PricingResults]:
    """One contract end to end: _simulate -> price -> FFT/mean (gbm_trainer.py:806-817)."""
    sr = simulate(c, normals, scheme=scheme, normalization=normalization)
    pr = price(c, sr)
    batches = normals.shape[1] // network_size
    return cf_estimate(pr.put_price, batches, network_size), pr


def seed_stream(mc_seed: int, skips: int, count: int) -> np.ndarray:
    """Per-matrix CuPy seeds of the reference, async_normals.py:319-326,391.

    Matrix ``k`` of the reference is seeded by the ``k``-th draw of
    ``default_rng(mc_seed).integers(0, 1e9)``; restoring discards ``skips`` draws.
    Recorded for documentation/tests of the snapshot contract — the B200 stream keys
    Philox on ``(mc_seed, k)`` directly (oracle.philox).
    """
    rng = np.random.default_rng(mc_seed)
    rng.integers(0, 1_000_000_000, size=skips)
    return rng.integers(0, 1_000_000_000, size=count)
