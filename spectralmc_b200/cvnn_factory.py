"""Declarative construction of the CVNN that consumes the hot path's targets.

The configuration schema and ``build_model`` of the reference's ``spectralmc.cvnn_factory``
(/root/reference/src/spectralmc/cvnn_factory.py: schemas :56-156, builder :177-368): the same
class names, fields and defaults, the same nesting of the emitted modules (an activation wraps its
layer in a two-element ``ComplexSequential``, a one-element sequence collapses to the element, a
width mismatch at the end appends a projection ``ComplexLinear``) and the same RNG discipline
(weights drawn under ``torch.manual_seed(cfg.seed)`` inside ``fork_rng`` on the CPU) — so a model
built here has the reference's parameter names, shapes AND initial values
(tests/test_oracle_cvnn_factory.py pins all three on outputs of the reference's own factory).

Out of scope: the safetensors / protobuf (de)serialisation helpers of the reference's module
(``load_model``, ``get_safetensors``, :370-431), which belong to its storage layer.
"""

from __future__ import annotations

from enum import Enum
from typing import Union

import torch
from pydantic import BaseModel, ConfigDict, PositiveInt
from torch import nn

from spectralmc_b200.cvnn import (
    ComplexLinear,
    ComplexResidual,
    ComplexSequential,
    CovarianceComplexBatchNorm,
    NaiveComplexBatchNorm,
    modReLU,
    zReLU,
)
from spectralmc_b200.numerical import Precision
from spectralmc_b200.result import Failure, Result, Success
from spectralmc_b200.validation import validate_model

_FROZEN = ConfigDict(frozen=True, extra="forbid")


class ActivationKind(str, Enum):
    Z_RELU = "zReLU"
    MOD_RELU = "modReLU"


class LayerKind(str, Enum):
    LINEAR = "ComplexLinear"
    BN_NAIVE = "NaiveComplexBatchNorm"
    BN_COV = "CovarianceComplexBatchNorm"
    SEQ = "Sequential"
    RES = "Residual"


class WidthSpec(BaseModel):
    model_config = _FROZEN


class PreserveWidth(WidthSpec):
    """Keep the width of the incoming signal."""


class ExplicitWidth(WidthSpec):
    value: PositiveInt


class ActivationCfg(BaseModel):
    kind: ActivationKind
    model_config = _FROZEN


class LinearCfg(BaseModel):
    kind: LayerKind = LayerKind.LINEAR
    width: Union[ExplicitWidth, PreserveWidth] = PreserveWidth()
    bias: bool = True
    activation: ActivationCfg | None = None
    model_config = _FROZEN


class _BatchNormCfg(BaseModel):
    eps: float = 1e-5
    momentum: float = 0.1
    affine: bool = True
    track_running_stats: bool = True
    activation: ActivationCfg | None = None
    model_config = _FROZEN


class NaiveBNCfg(_BatchNormCfg):
    kind: LayerKind = LayerKind.BN_NAIVE


class CovBNCfg(_BatchNormCfg):
    kind: LayerKind = LayerKind.BN_COV


class SequentialCfg(BaseModel):
    kind: LayerKind = LayerKind.SEQ
    layers: list["LayerCfg"]
    activation: ActivationCfg | None = None
    model_config = _FROZEN


class ResidualCfg(BaseModel):
    kind: LayerKind = LayerKind.RES
    body: SequentialCfg
    projection: LinearCfg | None = None
    activation: ActivationCfg | None = None
    model_config = _FROZEN


LayerCfg = Union[LinearCfg, NaiveBNCfg, CovBNCfg, SequentialCfg, ResidualCfg]
SequentialCfg.model_rebuild()
ResidualCfg.model_rebuild()


class CVNNConfig(BaseModel):
    dtype: Precision
    layers: list[LayerCfg]
    seed: PositiveInt
    final_activation: ActivationCfg | None = None
    model_config = _FROZEN


class WidthMismatch(BaseModel):
    """A residual's explicit projection does not produce the width of its body."""

    message: str
    model_config = _FROZEN


def build_cvnn_config(*, dtype: Precision | torch.dtype | str, layers: list[LayerCfg], seed: int,
                      final_activation: ActivationCfg | None = None) -> Result[CVNNConfig, object]:
    if isinstance(dtype, torch.dtype):
        dtype = {torch.float32: Precision.float32, torch.float64: Precision.float64}.get(dtype, str(dtype))
    return validate_model(CVNNConfig, dtype=dtype, layers=layers, seed=seed, final_activation=final_activation)


# ----------------------------------------------------------------------------- builder
def _activation(cfg: ActivationCfg, width: int) -> nn.Module:
    return modReLU(width) if cfg.kind is ActivationKind.MOD_RELU else zReLU()


def _collapse(mods: list[nn.Module]) -> nn.Module:
    return mods[0] if len(mods) == 1 else ComplexSequential(*mods)


def _activated(mod: nn.Module, act: ActivationCfg | None, width: int) -> nn.Module:
    return mod if act is None else ComplexSequential(mod, _activation(act, width))


def _chain(cfgs: list[LayerCfg], width: int) -> Result[tuple[list[nn.Module], int], WidthMismatch]:
    mods: list[nn.Module] = []
    for cfg in cfgs:
        built = _build(cfg, width)
        if isinstance(built, Failure):
            return built
        mod, width = built.value
        mods.append(mod)
    return Success((mods, width))


def _build(cfg: LayerCfg, width: int) -> Result[tuple[nn.Module, int], WidthMismatch]:
    """Module for ``cfg`` on a signal of ``width`` features, and the width it leaves."""
    if isinstance(cfg, LinearCfg):
        out_w = cfg.width.value if isinstance(cfg.width, ExplicitWidth) else width
        return Success((_activated(ComplexLinear(width, out_w, bias=cfg.bias), cfg.activation, out_w), out_w))
    if isinstance(cfg, (NaiveBNCfg, CovBNCfg)):
        cls = NaiveComplexBatchNorm if isinstance(cfg, NaiveBNCfg) else CovarianceComplexBatchNorm
        bn = cls(width, eps=cfg.eps, momentum=cfg.momentum, affine=cfg.affine, track_running_stats=cfg.track_running_stats)
        return Success((_activated(bn, cfg.activation, width), width))
    if isinstance(cfg, SequentialCfg):
        inner = _chain(cfg.layers, width)
        if isinstance(inner, Failure):
            return inner
        mods, out_w = inner.value
        return Success((_activated(_collapse(mods), cfg.activation, out_w), out_w))
    # ResidualCfg: body, then the skip path (explicit projection, none if widths agree, else an automatic one)
    body = _build(cfg.body, width)
    if isinstance(body, Failure):
        return body
    body_mod, body_w = body.value
    proj: nn.Module | None = None
    if cfg.projection is not None:
        built = _build(cfg.projection, width)
        if isinstance(built, Failure):
            return built
        proj, proj_w = built.value
        if proj_w != body_w:
            return Failure(WidthMismatch(message=f"Residual projection width {proj_w} does not match body width {body_w}."))
    elif body_w != width:
        proj = ComplexLinear(width, body_w)
    post = _activation(cfg.activation, body_w) if cfg.activation is not None else None
    return Success((ComplexResidual(body=body_mod, proj=proj, post_act=post), body_w))


def build_model(*, n_inputs: int, n_outputs: int, cfg: CVNNConfig) -> Result[nn.Module, WidthMismatch]:
    """Materialise ``cfg`` on the CPU in ``cfg.dtype`` without disturbing the caller's RNG stream
    (reference :343-368); move the result with ``.to(device, dtype)`` as the reference's callers do."""
    dtype = cfg.dtype.to_torch()
    previous = torch.get_default_dtype()
    with torch.random.fork_rng(devices=[]):
        torch.set_default_dtype(dtype)
        try:
            torch.manual_seed(cfg.seed)
            with torch.device("cpu"):
                chain = _chain(cfg.layers, n_inputs)
                if isinstance(chain, Failure):
                    return chain
                mods, width = chain.value
                net = _collapse(mods)
                if width != n_outputs:
                    net, width = ComplexSequential(net, ComplexLinear(width, n_outputs)), n_outputs
                net = _activated(net, cfg.final_activation, width)
        finally:
            torch.set_default_dtype(previous)
    return Success(net)
