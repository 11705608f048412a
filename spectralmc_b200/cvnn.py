"""Minimal complex-valued network — the downstream CONSUMER of the hot path's targets.

Out of the hot-path scope (the CVNN "stays PyTorch"); this is just enough of the
reference's ``spectralmc.cvnn`` surface (``ComplexLinear`` = four real matmuls, cvnn.py:65-146;
``modReLU`` cvnn.py:168-210; ``ComplexSequential`` cvnn.py:439-470) to run a training step on the
``[C, N]`` complex targets, with models mapping ``(real, imag) -> (real, imag)``.
"""

from __future__ import annotations

import math

import torch
from torch import nn


class ComplexLinear(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True) -> None:
        super().__init__()
        self.real_weight = nn.Parameter(torch.empty(out_features, in_features))
        self.imag_weight = nn.Parameter(torch.empty(out_features, in_features))
        self.real_bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        self.imag_bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        bound = 1.0 / math.sqrt(in_features)
        nn.init.uniform_(self.real_weight, -bound, bound)
        nn.init.uniform_(self.imag_weight, -bound, bound)

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        out_r = real @ self.real_weight.T - imag @ self.imag_weight.T
        out_i = real @ self.imag_weight.T + imag @ self.real_weight.T
        if self.real_bias is not None:
            out_r, out_i = out_r + self.real_bias, out_i + self.imag_bias
        return out_r, out_i


class modReLU(nn.Module):
    """z -> ReLU(|z| + b) * z / |z|."""

    def __init__(self, num_features: int) -> None:
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(num_features))

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        mod = torch.sqrt(real * real + imag * imag + 1e-12)
        scale = torch.relu(mod + self.bias) / mod
        return real * scale, imag * scale


class ComplexSequential(nn.Module):
    def __init__(self, *layers: nn.Module) -> None:
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        for layer in self.layers:
            real, imag = layer(real, imag)
        return real, imag


def make_cvnn(n_inputs: int, n_outputs: int, *, hidden_width: int = 32, seed: int = 0,
              dtype: torch.dtype = torch.float32, device: torch.device | str = "cuda") -> ComplexSequential:
    """6 -> hidden (modReLU) -> N, the shape of the reference's test network
    (tests/helpers/factories.py:69-105); weights seeded inside a forked RNG as ``build_model`` does."""
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        net = ComplexSequential(ComplexLinear(n_inputs, hidden_width), modReLU(hidden_width), ComplexLinear(hidden_width, n_outputs))
    return net.to(device=device, dtype=dtype)
