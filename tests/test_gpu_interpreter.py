"""The effect-operator seam (GenerateNormals / SimulatePaths / ComputeFFT) on the C ABI, checked
against the oracle; the reference's real MonteCarloInterpreter is never exercised by its own tests
(SURVEY.md §3D), so the acceptance here is numerical parity with the engine path."""

from __future__ import annotations

import asyncio

import numpy as np
import pytest
import torch

from oracle import gbm as ogbm
from oracle import philox
from spectralmc_b200.effects import ComputeFFT, ForwardNormalization, GenerateNormals, PathScheme, SimulatePaths
from spectralmc_b200.interpreter import MonteCarloOperators, TensorRegistry
from tests.helpers import expect_failure, expect_success, rel_elem, rel_max

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,np_dtype,tol", [(torch.float32, np.float32, 1e-5), (torch.float64, np.float64, 1e-12)])
@pytest.mark.parametrize("norm", [ForwardNormalization.RAW, ForwardNormalization.NORMALIZE])
def test_effect_sequence_matches_oracle(dtype, np_dtype, tol, norm) -> None:
    reg = TensorRegistry()
    ops = MonteCarloOperators(reg, dtype=dtype)
    T, P = 12, 1024
    expect_success(ops.run(GenerateNormals(rows=T, cols=P, seed=42, skip=3, output_tensor_id="generated_normals")))
    sims = expect_success(asyncio.run(ops.interpret(SimulatePaths(
        spot=100.0, strike=100.0, rate=0.05, dividend=0.0, vol=0.2, expiry=1.0, timesteps=T, batches=P,
        path_scheme=PathScheme.LOG_EULER, normalization=norm, input_normals_id="generated_normals", output_tensor_id="paths"))))
    z = philox.normals_matrix(T, P, np_dtype, 42, 3)
    sr = ogbm.simulate(ogbm.Contract(100.0, 100.0, 1.0, 0.05, 0.0, 0.2), z, normalization=norm.value)
    assert rel_elem(sims.cpu().numpy(), sr.sims) <= tol
    # the normals in the registry are untouched (the kernel ran on a copy)
    kept = expect_success(reg.get_tensor("generated_normals")).cpu().numpy()
    assert np.max(np.abs(kept - philox.normals_matrix(T, P, np_dtype, 42, 3))) <= 1e-4
    fft = expect_success(ops.run(ComputeFFT(input_tensor_id="paths", axis=1, output_tensor_id="fft")))
    assert rel_max(fft.cpu().numpy(), np.fft.fft(sr.sims.astype(np.float64), axis=1)) <= (1e-5 if dtype == torch.float32 else 1e-12)


def test_missing_inputs_are_failures_not_exceptions() -> None:
    ops = MonteCarloOperators(TensorRegistry())
    assert "not found" in expect_failure(ops.run(SimulatePaths(input_normals_id="nope"))).message
    assert "not found" in expect_failure(ops.run(ComputeFFT(input_tensor_id="nope"))).message
    assert "shape" in expect_failure(ops.run(GenerateNormals(rows=0, cols=4, seed=1))).message
