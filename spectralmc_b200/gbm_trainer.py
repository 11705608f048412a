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

SURVEY.md §8f-4: when the network is a ComplexSequential of ComplexLinear / modReLU / zReLU
(``spectralmc_b200.cvnn.describe``), ``_torch_step`` runs through the C ABI's fused CVNN step
(``smc_cvnn_train_step``) and, by default, replays as ONE CUDA graph per training step — a step is
then: pinned H2D of the contracts, the simulation launches (their matrix index changes every
step, so they stay ordinary launches), two small device copies into the graph's static inputs,
one graph launch.  The losses stay on the device until ``train`` returns.  Any other network
takes the generic torch route below (``fused_step=False`` forces it).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import numpy as np
import torch
from torch import nn, optim

from spectralmc_b200.cvnn import FusedCVNN, describe
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
                 global_step: int = 0, process_group=None, fused_step: bool | None = None, cuda_graph: bool = True) -> None:
        self._cfg, self._sp = cfg, cfg.sim_params
        self._engine = BlackScholes(cfg)
        self._cvnn = cvnn
        self._dtype = self._sp.dtype.to_torch()
        self._device = self._engine._device
        self._domain_bounds = domain_bounds
        self._sobol_skip, self._global_step = sobol_skip, global_step
        self._group = process_group
        self._optimizer: optim.Optimizer | None = None
        supported = describe(cvnn) is not None
        if fused_step and not supported:
            raise ValueError("fused_step=True needs a ComplexSequential of ComplexLinear / modReLU / zReLU")
        self._use_fused = supported if fused_step is None else fused_step
        self._use_graph = cuda_graph
        self._fused: FusedCVNN | None = None
        self._graphs: dict[int, _StepGraph] = {}
        # the sampler is seeded with mc_seed and resumed with sobol_skip (reference :703-710)
        self._sampler_result = SobolSampler.create(BlackScholes.Inputs, domain_bounds, config=SobolConfig(seed=self._sp.mc_seed, skip=sobol_skip))

    # ------------------------------------------------------------------ data path
    def _simulate_fft(self, contract: BlackScholes.Inputs) -> Result[torch.Tensor, object]:
        """One contract -> ``[N]`` complex CF estimate (reference :806-817)."""
        return self._engine.simulate_fft(contract)

    def _upload(self, rows: np.ndarray) -> torch.Tensor:
        host = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64)).pin_memory()
        return host.to(self._device, non_blocking=True)

    def targets(self, rows: np.ndarray | torch.Tensor) -> Result[torch.Tensor, object]:
        """``[C, 6]`` contracts -> ``[C, N]`` complex targets on device (one C-ABI call sequence)."""
        dev = rows if isinstance(rows, torch.Tensor) else self._upload(rows)
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

    def _fused_step(self, contracts: torch.Tensor, targets: torch.Tensor, loss_slot: torch.Tensor) -> None:
        """``_torch_step`` through the C ABI; ``contracts`` is the ``[C, 6]`` float64 device batch."""
        fused = self._fused
        assert fused is not None
        rows = contracts.shape[0]
        if not self._use_graph:
            real_in = contracts.to(self._dtype)
            loss_slot.copy_(fused.train_step(real_in, torch.zeros_like(real_in), targets))
            return
        g = self._graphs.get(rows)
        if g is None:
            g = self._graphs[rows] = _StepGraph(fused, rows)
        g.real_in.copy_(contracts)  # float64 -> dtype, as torch.as_tensor(rows, dtype) narrows (reference :1779-1781)
        g.targets.copy_(targets)
        g.graph.replay()
        loss_slot.copy_(g.loss)

    def train(self, config: TrainingConfig) -> Result[list[float], object]:
        if isinstance(self._sampler_result, Failure):
            return self._sampler_result
        sampler = self._sampler_result.value
        if self._use_fused and self._fused is None:
            self._fused = FusedCVNN(self._cvnn, lr=config.learning_rate)
        elif self._use_fused and self._fused.hyper.lr != config.learning_rate:
            self._fused.hyper.lr = config.learning_rate  # hyper-parameters are baked into captured launches
            self._graphs.clear()
        elif not self._use_fused and self._optimizer is None:
            self._optimizer = optim.Adam(self._cvnn.parameters(), lr=config.learning_rate)
        self._cvnn.train()
        losses = torch.zeros(max(config.num_batches, 1), dtype=torch.float64, device=self._device)
        for i in range(config.num_batches):
            drawn = sampler.sample_array(config.batch_size)
            if isinstance(drawn, Failure):
                return drawn
            self._sobol_skip += config.batch_size
            contracts = self._upload(drawn.value)
            got = self.targets(contracts)
            if isinstance(got, Failure):
                return got
            targets = got.value.detach()  # already a torch tensor: the DLPack hand-off is the identity
            if self._use_fused:
                self._fused_step(contracts, targets, losses[i : i + 1])
            else:
                real_in = contracts.to(self._dtype)
                loss, _ = self._torch_step(real_in, torch.zeros_like(real_in), targets, self._optimizer)
                losses[i : i + 1].copy_(loss.detach())
            self._global_step += 1
        return Success([float(x) for x in losses[: config.num_batches].cpu()])

    def predict_price(self, inputs: Sequence[BlackScholes.Inputs]) -> list[float]:
        """CVNN forward -> ifft -> mean -> real part = DC / N (reference :1709-1735)."""
        self._cvnn.eval()
        real_in, imag_in = _split_inputs(inputs, dtype=self._dtype, device=self._device)
        if self._use_fused:
            # mean_n ifft(S)[n] = S[0] / N: the price is the DC bin over N, no transform needed
            fused = self._fused if self._fused is not None else FusedCVNN(self._cvnn)
            self._fused = fused
            pred_r, _ = fused.forward(real_in, imag_in)
            return [float(x) for x in (pred_r[:, 0] / pred_r.shape[1]).cpu()]
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
            optimizer_state=self._optimizer_state(),
        )

    def _optimizer_state(self) -> dict | None:
        """Adam state in ``torch.optim.Adam.state_dict()`` layout for both step routes."""
        if self._use_fused:
            return None if self._fused is None or int(self._fused.step.item()) == 0 else _to_cpu(self._fused.optimizer_state_dict())
        return None if self._optimizer is None else _to_cpu(self._optimizer.state_dict())

    @classmethod
    def restore(cls, snap: PricerSnapshot, domain_bounds: DomainBounds, cvnn: nn.Module, *, learning_rate: float = 1e-2,
                **kwargs) -> "GbmCVNNPricer":
        cvnn.load_state_dict(snap.cvnn_state)
        self = cls(snap.cfg, domain_bounds, cvnn, sobol_skip=snap.sobol_skip, global_step=snap.global_step, **kwargs)
        if snap.optimizer_state is not None:
            if self._use_fused:
                self._fused = FusedCVNN(cvnn, lr=learning_rate)
                self._fused.load_optimizer_state_dict(snap.optimizer_state)
            else:
                self._optimizer = optim.Adam(cvnn.parameters(), lr=learning_rate)
                self._optimizer.load_state_dict(snap.optimizer_state)
        return self


class _StepGraph:
    """One captured ``smc_cvnn_train_step`` for a fixed batch size, with its static inputs."""

    def __init__(self, fused: FusedCVNN, rows: int) -> None:
        dev, dt = fused.device, fused.dtype
        self.real_in = torch.zeros((rows, fused.n_inputs), dtype=dt, device=dev)
        self.imag_in = torch.zeros_like(self.real_in)  # imag_in = 0 (reference :1782)
        self.targets = torch.zeros((rows, fused.n_outputs), dtype=torch.complex64 if dt == torch.float32 else torch.complex128, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float64, device=dev)
        fused.warm_up(rows)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            fused.train_step(self.real_in, self.imag_in, self.targets, self.loss)


def _to_cpu(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().clone()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj
