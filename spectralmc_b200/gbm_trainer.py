"""Data path of the reference's ``GbmCVNNPricer`` on the fused device path.

Covers what SURVEY.md §8a puts on the hot path from /root/reference/src/spectralmc/gbm_trainer.py:
``_simulate_fft`` :806-817 (one contract -> CF estimate), the per-batch target construction of
``_run_batch`` :1539-1565 (Sobol batch -> ``[C, N]`` complex targets -> DLPack/torch hand-off ->
``_torch_step`` :819-835), ``_split_inputs`` :1775-1783, ``predict_price`` :1709-1735 and the
``snapshot`` bookkeeping (``sobol_skip``, ``global_step``, engine ``skip``) :756-800.  The
reference's commit plans, S3 store, TensorBoard logging and effect descriptions are out of scope.

The per-contract Python loop with >= 5 host synchronisations per contract becomes
ONE C-ABI call per training step; targets are produced directly as a torch tensor, so the
``cp.asarray`` stacking and ``torch.from_dlpack`` copy of the reference vanish.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import numpy as np
import torch
from torch import nn, optim

from spectralmc_b200.distributed import sharded_cf_targets
from spectralmc_b200.gbm import BlackScholes, BlackScholesConfig
from spectralmc_b200.result import Failure, Result, Success
from spectralmc_b200.sobol_sampler import DomainBounds, SobolConfig, SobolSampler


def _split_inputs(rows: np.ndarray | Sequence[BlackScholes.Inputs], *, dtype: torch.dtype, device: torch.device) -> tuple[torch.Tensor, torch.Tensor]:
    """Contracts -> CVNN ``(real, imag)`` inputs in field order X0,K,T,r,d,v (reference :1775-1783)."""
    if not isinstance(rows, np.ndarray):
        fields = list(BlackScholes.Inputs.model_fields.keys())
        rows = np.array([[float(getattr(inp, f)) for f in fields] for inp in rows], dtype=np.float64)
    real = torch.as_tensor(rows, dtype=dtype, device=device)
    return real, torch.zeros_like(real)


@dataclass
class TrainingConfig:
    num_batches: int
    batch_size: int
    learning_rate: float = 1e-2


@dataclass
class PricerSnapshot:
    cfg: BlackScholesConfig
    sobol_skip: int
    global_step: int
    cvnn_state: dict = field(default_factory=dict)
    optimizer_state: dict | None = None


class GbmCVNNPricer:
    """Trains a CVNN on CF targets produced by the fused Monte-Carlo path."""

    def __init__(self, cfg: BlackScholesConfig, domain_bounds: DomainBounds, cvnn: nn.Module, *, sobol_skip: int = 0,
                 global_step: int = 0, process_group=None) -> None:
        self._cfg, self._sp = cfg, cfg.sim_params
        self._engine = BlackScholes(cfg)
        self._cvnn = cvnn
        self._dtype = self._sp.dtype.to_torch()
        self._device = self._engine._device
        self._domain_bounds = domain_bounds
        self._sobol_skip, self._global_step = sobol_skip, global_step
        self._group = process_group
        self._optimizer: optim.Optimizer | None = None
        # the sampler is seeded with mc_seed and resumed with sobol_skip (reference :703-710)
        self._sampler_result = SobolSampler.create(BlackScholes.Inputs, domain_bounds, config=SobolConfig(seed=self._sp.mc_seed, skip=sobol_skip))

    # ------------------------------------------------------------------ data path
    def _simulate_fft(self, contract: BlackScholes.Inputs) -> Result[torch.Tensor, object]:
        """One contract -> ``[N]`` complex CF estimate (reference :806-817)."""
        return self._engine.simulate_fft(contract)

    def targets(self, rows: np.ndarray) -> Result[torch.Tensor, object]:
        """``[C, 6]`` contracts -> ``[C, N]`` complex targets on device (one C-ABI call sequence)."""
        host = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64)).pin_memory()
        dev = host.to(self._device, non_blocking=True)
        if self._group is not None:
            return Success(sharded_cf_targets(self._engine, dev, group=self._group))
        return self._engine.cf_targets(dev)

    def _torch_step(self, real_in, imag_in, targets, optimizer) -> tuple[torch.Tensor, float]:
        """Forward / MSE on real and imaginary parts / backward / Adam (reference :819-835)."""
        pred_r, pred_i = self._cvnn(real_in, imag_in)
        loss = nn.functional.mse_loss(pred_r, torch.real(targets)) + nn.functional.mse_loss(pred_i, torch.imag(targets))
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        optimizer.step()
        grad_norm = float(torch.nn.utils.clip_grad_norm_(self._cvnn.parameters(), float("inf")))
        return loss, grad_norm

    def train(self, config: TrainingConfig) -> Result[list[float], object]:
        if isinstance(self._sampler_result, Failure):
            return self._sampler_result
        sampler = self._sampler_result.value
        if self._optimizer is None:
            self._optimizer = optim.Adam(self._cvnn.parameters(), lr=config.learning_rate)
        self._cvnn.train()
        losses: list[torch.Tensor] = []
        for _ in range(config.num_batches):
            drawn = sampler.sample_array(config.batch_size)
            if isinstance(drawn, Failure):
                return drawn
            self._sobol_skip += config.batch_size
            got = self.targets(drawn.value)
            if isinstance(got, Failure):
                return got
            targets = got.value.detach()  # already a torch tensor: the DLPack hand-off is the identity
            real_in, imag_in = _split_inputs(drawn.value, dtype=self._dtype, device=self._device)
            loss, _ = self._torch_step(real_in, imag_in, targets, self._optimizer)
            losses.append(loss.detach())
            self._global_step += 1
        return Success([float(x) for x in torch.stack(losses).cpu()] if losses else [])

    def predict_price(self, inputs: Sequence[BlackScholes.Inputs]) -> list[float]:
        """CVNN forward -> ifft -> mean -> real part = DC / N (reference :1709-1735)."""
        self._cvnn.eval()
        real_in, imag_in = _split_inputs(inputs, dtype=self._dtype, device=self._device)
        with torch.no_grad():
            pred_r, pred_i = self._cvnn(real_in, imag_in)
            spectrum = torch.complex(pred_r, pred_i)
            price = torch.fft.ifft(spectrum, dim=1).mean(dim=1).real
        return [float(x) for x in price.cpu()]

    def snapshot(self) -> PricerSnapshot:
        cfg = self._engine.snapshot()
        if isinstance(cfg, Failure):
            raise AssertionError(f"engine snapshot failed: {cfg.error}")
        return PricerSnapshot(
            cfg=cfg.value, sobol_skip=self._sobol_skip, global_step=self._global_step,
            cvnn_state={k: v.detach().cpu().clone() for k, v in self._cvnn.state_dict().items()},
            optimizer_state=None if self._optimizer is None else _to_cpu(self._optimizer.state_dict()),
        )

    @classmethod
    def restore(cls, snap: PricerSnapshot, domain_bounds: DomainBounds, cvnn: nn.Module, *, learning_rate: float = 1e-2) -> "GbmCVNNPricer":
        cvnn.load_state_dict(snap.cvnn_state)
        self = cls(snap.cfg, domain_bounds, cvnn, sobol_skip=snap.sobol_skip, global_step=snap.global_step)
        if snap.optimizer_state is not None:
            self._optimizer = optim.Adam(cvnn.parameters(), lr=learning_rate)
            self._optimizer.load_state_dict(snap.optimizer_state)
        return self


def _to_cpu(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().clone()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj
