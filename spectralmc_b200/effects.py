"""The Monte-Carlo effect vocabulary of the reference, as plain data.

``PathScheme`` / ``ForwardNormalization`` and the three operator ADTs
``GenerateNormals`` / ``SimulatePaths`` / ``ComputeFFT`` keep the names, fields and defaults of
/root/reference/src/spectralmc/effects/montecarlo.py:24-112; they are the operator seam the
reference's ``MonteCarloInterpreter`` interprets (effects/interpreter.py:552-712).
``spectralmc_b200.interpreter.MonteCarloOperators`` executes them on the C ABI.
"""

from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Literal


class PathScheme(str, Enum):
    LOG_EULER = "log_euler"
    SIMPLE_EULER = "simple_euler"


class ForwardNormalization(str, Enum):
    NORMALIZE = "normalize_forwards"
    RAW = "raw_paths"


@dataclass(frozen=True)
class GenerateNormals:
    kind: Literal["GenerateNormals"] = "GenerateNormals"
    rows: int = 0
    cols: int = 0
    seed: int = 0
    skip: int = 0
    output_tensor_id: str = "normals"


@dataclass(frozen=True)
class SimulatePaths:
    kind: Literal["SimulatePaths"] = "SimulatePaths"
    spot: float = 100.0
    strike: float = 100.0
    rate: float = 0.05
    dividend: float = 0.0
    vol: float = 0.2
    expiry: float = 1.0
    timesteps: int = 252
    batches: int = 1024
    path_scheme: PathScheme = PathScheme.LOG_EULER
    normalization: ForwardNormalization = ForwardNormalization.NORMALIZE
    input_normals_id: str = ""
    output_tensor_id: str = "paths"


@dataclass(frozen=True)
class ComputeFFT:
    kind: Literal["ComputeFFT"] = "ComputeFFT"
    input_tensor_id: str = ""
    axis: int = -1
    output_tensor_id: str = "fft"


MonteCarloEffect = GenerateNormals | SimulatePaths | ComputeFFT
