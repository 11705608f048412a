"""Effect-operator adapter: the three Monte-Carlo effects of the reference executed on the C ABI.

Mirrors ``MonteCarloInterpreter`` (/root/reference/src/spectralmc/effects/interpreter.py:537-712)
over a name -> tensor registry (the role of ``SharedRegistry.register_tensor/get_cupy_array`` and
``register_kernel``, effects/registry.py:95-137,502-529):

* ``GenerateNormals``  -> ``smc_philox_normals`` (matrix index = ``skip``; the reference "skips" by
  drawing ``skip`` scalars from one XORWOW stream, :578-581 — here ``skip`` selects the matrix of the
  counter-based stream, which is what the engine's ``skip`` means, gbm.py:372-380);
* ``SimulatePaths``    -> copy of the normals (:623) + ``smc_gbm_paths_inplace`` (256 threads per
  block, :626) + ``smc_normalize_rows`` when normalisation is requested (:660-665);
* ``ComputeFFT``       -> forward DFT along ``axis`` (:703): ``smc_fft_rows`` for a real 2-D tensor
  transformed along its last axis with ``N <= 8192`` (every case the Monte-Carlo path produces).
  This operator returns the FULL per-row spectrum, which the training path never needs (it
  consumes the batch MEAN, gbm_trainer.py:814-817 -> ``smc_cf_fft_mean``).  Other ranks / axes /
  complex inputs are outside the path and are rejected with a ``MonteCarloError``.

Unlike the reference interpreter, which is float32-only (:583), the dtype is a constructor
argument.  ``interpret`` is ``async`` like the reference's; ``run`` is the synchronous form.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Literal

import numpy as np
import torch

from spectralmc_b200 import _cabi
from spectralmc_b200.effects import ComputeFFT, ForwardNormalization, GenerateNormals, MonteCarloEffect, PathScheme, SimulatePaths
from spectralmc_b200.gbm import SimulateBlackScholes
from spectralmc_b200.result import Failure, Result, Success


@dataclass(frozen=True)
class MonteCarloError:
    """Same shape as the reference's effects/errors.py MonteCarloError."""

    message: str
    kind: Literal["MonteCarloError"] = "MonteCarloError"


class TensorRegistry:
    """Named device tensors and kernels shared between operators."""

    def __init__(self) -> None:
        self._tensors: dict[str, torch.Tensor] = {}
        self._kernels: dict[str, Callable[..., object]] = {"SimulateBlackScholes": SimulateBlackScholes}

    def register_tensor(self, tensor_id: str, tensor: torch.Tensor) -> Result[None, str]:
        if not tensor_id:
            return Failure("empty tensor id")
        self._tensors[tensor_id] = tensor
        return Success(None)

    def get_tensor(self, tensor_id: str) -> Result[torch.Tensor, str]:
        t = self._tensors.get(tensor_id)
        return Success(t) if t is not None else Failure(f"unknown tensor {tensor_id!r}")

    def get_torch_tensor(self, tensor_id: str) -> Result[torch.Tensor, str]:
        """Reference name (effects/registry.py:208); tensors here ARE torch tensors."""
        return self.get_tensor(tensor_id)

    def has_tensor(self, tensor_id: str) -> bool:
        return tensor_id in self._tensors

    def register_kernel(self, name: str, fn: Callable[..., object]) -> Result[None, str]:
        if not name:
            return Failure("empty kernel name")
        self._kernels[name] = fn
        return Success(None)

    def get_kernel(self, name: str) -> Result[Callable[..., object], str]:
        k = self._kernels.get(name)
        return Success(k) if k is not None else Failure(f"unknown kernel {name!r}")

    def has_kernel(self, name: str) -> bool:
        return name in self._kernels

    def clear_tensors(self) -> None:
        self._tensors.clear()


class MonteCarloOperators:
    def __init__(self, registry: TensorRegistry, dtype: torch.dtype = torch.float32) -> None:
        self._registry, self._dtype = registry, dtype

    async def interpret(self, effect: MonteCarloEffect) -> Result[object, MonteCarloError]:
        return self.run(effect)

    def run(self, effect: MonteCarloEffect) -> Result[object, MonteCarloError]:
        try:
            if isinstance(effect, GenerateNormals):
                return self._generate_normals(effect)
            if isinstance(effect, SimulatePaths):
                return self._simulate_paths(effect)
            if isinstance(effect, ComputeFFT):
                return self._compute_fft(effect)
        except (_cabi.SmcError, RuntimeError, ValueError) as exc:
            return Failure(MonteCarloError(message=str(exc)))
        raise AssertionError(f"not a Monte-Carlo effect: {effect!r}")

    def _store(self, tensor_id: str, tensor: torch.Tensor) -> Result[object, MonteCarloError]:
        stored = self._registry.register_tensor(tensor_id, tensor)
        if isinstance(stored, Failure):
            return Failure(MonteCarloError(message=f"Registry tensor error: {stored.error}"))
        return Success(tensor)

    def _generate_normals(self, effect: GenerateNormals) -> Result[object, MonteCarloError]:
        out = torch.empty((effect.rows, effect.cols), dtype=self._dtype, device="cuda")
        _cabi.philox_normals(out, effect.seed, effect.skip)
        return self._store(effect.output_tensor_id, out)

    def _simulate_paths(self, effect: SimulatePaths) -> Result[object, MonteCarloError]:
        got = self._registry.get_tensor(effect.input_normals_id)
        if isinstance(got, Failure):
            return Failure(MonteCarloError(message=f"Normals tensor not found: {effect.input_normals_id}"))
        sims = got.value.clone()  # the kernel works in place (reference :623)
        dt = effect.expiry / effect.timesteps
        kernel = self._registry.get_kernel("SimulateBlackScholes").unwrap()
        blocks = (sims.shape[1] + 255) // 256
        kernel[blocks, 256, torch.cuda.current_stream()](
            sims, effect.timesteps, dt, effect.spot, effect.rate, effect.dividend, effect.vol,
            effect.path_scheme is PathScheme.LOG_EULER,
        )
        if effect.normalization is ForwardNormalization.NORMALIZE:
            np_dtype = np.float32 if sims.dtype == torch.float32 else np.float64
            times = np.linspace(dt, effect.expiry, effect.timesteps, dtype=np_dtype)
            forwards = (effect.spot * np.exp((effect.rate - effect.dividend) * times)).astype(np_dtype)
            _cabi.normalize_rows(sims, torch.from_numpy(forwards).to(sims.device))
        return self._store(effect.output_tensor_id, sims)

    def _compute_fft(self, effect: ComputeFFT) -> Result[object, MonteCarloError]:
        got = self._registry.get_tensor(effect.input_tensor_id)
        if isinstance(got, Failure):
            return Failure(MonteCarloError(message=f"Tensor not found: {effect.input_tensor_id}"))
        t = got.value
        axis = effect.axis if effect.axis >= 0 else t.dim() + effect.axis
        if t.dim() != 2 or axis != 1 or t.is_complex():
            return Failure(MonteCarloError(message="ComputeFFT supports a real 2-D tensor transformed along its last axis"))
        return self._store(effect.output_tensor_id, _cabi.fft_rows(t.contiguous()))


# the reference's names for the two classes (effects/registry.py:95, effects/interpreter.py:537)
SharedRegistry = TensorRegistry
MonteCarloInterpreter = MonteCarloOperators
