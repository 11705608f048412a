"""Error ADTs returned (never raised) by the hot-path API.

Same names, fields and ``kind`` discriminators as the reference's
/root/reference/src/spectralmc/errors/async_normals.py:10-45, errors/gbm.py:21-79 and
errors/sampler.py, so ``match``-based callers keep working.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Union

from pydantic import ValidationError


# ---- normal supply (errors/async_normals.py) -------------------------------------------
@dataclass(frozen=True)
class InvalidDType:
    requested: str
    kind: Literal["InvalidDType"] = "InvalidDType"


@dataclass(frozen=True)
class InvalidShape:
    rows: int
    cols: int
    kind: Literal["InvalidShape"] = "InvalidShape"


@dataclass(frozen=True)
class SeedOutOfRange:
    seed: int
    kind: Literal["SeedOutOfRange"] = "SeedOutOfRange"


@dataclass(frozen=True)
class QueueEmpty:
    kind: Literal["QueueEmpty"] = "QueueEmpty"


@dataclass(frozen=True)
class QueueBusy:
    kind: Literal["QueueBusy"] = "QueueBusy"


# ---- engine (errors/gbm.py) ------------------------------------------------------------------
@dataclass(frozen=True)
class CudaRNGUnavailable:
    reason: str
    kind: Literal["CudaRNGUnavailable"] = "CudaRNGUnavailable"


NormGeneratorError = Union[InvalidShape, InvalidDType, QueueBusy, QueueEmpty, SeedOutOfRange, CudaRNGUnavailable]


@dataclass(frozen=True)
class InvalidSimulationParams:
    error: ValidationError
    kind: Literal["InvalidSimulationParams"] = "InvalidSimulationParams"


@dataclass(frozen=True)
class GPUMemoryLimitExceeded:
    total_paths: int
    max_paths: int
    network_size: int
    batches_per_mc_run: int
    kind: Literal["GPUMemoryLimitExceeded"] = "GPUMemoryLimitExceeded"


@dataclass(frozen=True)
class InvalidBlackScholesConfig:
    error: ValidationError
    kind: Literal["InvalidBlackScholesConfig"] = "InvalidBlackScholesConfig"


@dataclass(frozen=True)
class NormalsUnavailable:
    error: NormGeneratorError
    kind: Literal["NormalsUnavailable"] = "NormalsUnavailable"


@dataclass(frozen=True)
class NormalsGenerationFailed:
    error: NormGeneratorError
    kind: Literal["NormalsGenerationFailed"] = "NormalsGenerationFailed"


@dataclass(frozen=True)
class DeviceKernelFailed:
    """A C-ABI call failed (status + message from ``smc_last_error``).  New in this package: the
    reference surfaces device errors as exceptions from Numba/CuPy."""

    status: int
    message: str
    kind: Literal["DeviceKernelFailed"] = "DeviceKernelFailed"


# ---- Sobol sampler (errors/sampler.py) --------------------------------------------------------
@dataclass(frozen=True)
class DimensionMismatch:
    kind: Literal["DimensionMismatch"] = "DimensionMismatch"
    expected_fields: tuple[str, ...] = ()
    provided_fields: tuple[str, ...] = ()


@dataclass(frozen=True)
class InvalidBounds:
    message: str
    kind: Literal["InvalidBounds"] = "InvalidBounds"


@dataclass(frozen=True)
class NegativeSamples:
    n_samples: int
    kind: Literal["NegativeSamples"] = "NegativeSamples"


@dataclass(frozen=True)
class BoundSpecInvalid:
    lower: float
    upper: float
    kind: Literal["BoundSpecInvalid"] = "BoundSpecInvalid"


@dataclass(frozen=True)
class SamplerValidationFailed:
    error: ValidationError
    kind: Literal["SamplerValidationFailed"] = "SamplerValidationFailed"


SamplerError = Union[DimensionMismatch, InvalidBounds, BoundSpecInvalid, NegativeSamples, SamplerValidationFailed]


# ---- trainer (errors/trainer.py; gbm_trainer.py:525-555) --------------------------------------
@dataclass(frozen=True)
class SamplerInitFailed:
    error: object
    kind: Literal["SamplerInitFailed"] = "SamplerInitFailed"


@dataclass(frozen=True)
class InvalidTrainerConfig:
    message: str
    kind: Literal["InvalidTrainerConfig"] = "InvalidTrainerConfig"


@dataclass(frozen=True)
class InvalidTrainingConfig:
    num_batches: int
    batch_size: int
    learning_rate: float
    message: str
    kind: Literal["InvalidTrainingConfig"] = "InvalidTrainingConfig"


@dataclass(frozen=True)
class OptimizerStateSerializationFailed:
    message: str
    kind: Literal["OptimizerStateSerializationFailed"] = "OptimizerStateSerializationFailed"


@dataclass(frozen=True)
class PredictionFailed:
    message: str
    kind: Literal["PredictionFailed"] = "PredictionFailed"


@dataclass(frozen=True)
class DeviceDTypeError:
    """The CVNN's parameters do not share one device / full-precision dtype, or the dtype differs
    from the simulation's (the reference asserts the latter, gbm_trainer.py:685-687)."""

    message: str
    kind: Literal["DeviceDTypeError"] = "DeviceDTypeError"


@dataclass(frozen=True)
class DeviceNotCUDA:
    device: str
    message: str
    kind: Literal["DeviceNotCUDA"] = "DeviceNotCUDA"
