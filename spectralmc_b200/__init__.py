"""spectralmc_b200 — B200-native batch-generation hot path of SpectralMC.

Importing the package loads ``lib/libspectralmc_b200.so`` (hand-written sm_100a kernels behind
a C ABI).  There is no fallback: a missing library raises ImportError.
"""

from spectralmc_b200 import _cabi  # noqa: F401  (fails loudly if the library is missing)
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import (
    BlackScholes,
    BlackScholesConfig,
    SimulateBlackScholes,
    SimulationParams,
    build_black_scholes_config,
    build_simulation_params,
)
from spectralmc_b200.numerical import Precision

__version__ = "0.1.0"
__all__ = [
    "BlackScholes",
    "BlackScholesConfig",
    "ForwardNormalization",
    "PathScheme",
    "Precision",
    "SimulateBlackScholes",
    "SimulationParams",
    "build_black_scholes_config",
    "build_simulation_params",
]
