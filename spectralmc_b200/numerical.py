"""``Precision`` — the reference's dtype enum (/root/reference/src/spectralmc/models/numerical.py:124-182)
with torch/NumPy conversions (CuPy is not on this path)."""

from __future__ import annotations

from enum import Enum

import numpy as np
import torch

_TO_COMPLEX = {"float32": "complex64", "float64": "complex128"}
_TO_REAL = {v: k for k, v in _TO_COMPLEX.items()}


class Precision(str, Enum):
    float32 = "float32"
    float64 = "float64"
    complex64 = "complex64"
    complex128 = "complex128"

    def to_numpy(self) -> np.dtype:
        return np.dtype(self.value)

    def to_torch(self) -> torch.dtype:
        return getattr(torch, self.value)

    @classmethod
    def from_torch(cls, dtype: torch.dtype) -> "Precision":
        return cls(str(dtype).replace("torch.", ""))

    def to_complex(self) -> "Precision":
        return Precision(_TO_COMPLEX.get(self.value, self.value))

    def to_real(self) -> "Precision":
        return Precision(_TO_REAL.get(self.value, self.value))

    def is_real(self) -> bool:
        return self.value in _TO_COMPLEX
