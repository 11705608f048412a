"""Data path of the reference's ``GbmCVNNPricer`` on the fused device path.

Covers what SURVEY.md §8a puts on the hot path from /root/reference/src/spectralmc/gbm_trainer.py:
``_simulate_fft`` :806-817 (one contract -> CF estimate), the per-batch target construction of
``_run_batch`` :1539-1565 (Sobol batch -> ``[C, N]`` complex targets -> DLPack/torch hand-off ->
``_torch_step`` :819-835), ``_split_inputs`` :1775-1783, ``predict_price`` :1709-1735 and the
``snapshot`` bookkeeping (``sobol_skip``, ``global_step``, engine ``skip``) :756-800.  The
reference's commit plans, S3 store, TensorBoard logging and effect descriptions are out of scope.
The public shape follows the reference: ``build_training_config`` :261-298, ``GbmCVNNPricerConfig``
:301-313, ``StepMetrics`` / ``TrainingResult`` :336-361, ``GbmCVNNPricer.create(cfg) -> Result``
:600-663, ``train(config, logger=...) -> Result[TrainingResult, ...]`` :1456-1683,
``snapshot() -> Result[GbmCVNNPricerConfig, ...]`` :756-800 and
``predict_price(inputs) -> Result[list[HostPricingResults], ...]`` :1709-1769, so the reference's
tests/test_gbm_trainer.py reads the same against this module.  Deviation: ``optimizer_state`` is
``torch.optim.Adam.state_dict()`` with CPU tensors (the reference wraps the same dictionary in its
``AdamOptimizerState`` serialisation model, out of scope here).

The per-contract Python loop with >= 5 host synchronisations per contract becomes
ONE C-ABI call per training step; targets are produced directly as a torch tensor, so the
``cp.asarray`` stacking and ``torch.from_dlpack`` copy of the reference vanish.

SURVEY.md §8f-4: when the network is a ComplexSequential of ComplexLinear / modReLU / zReLU
(``spectralmc_b200.cvnn.describe``), ``_torch_step`` runs through the C ABI's fused CVNN step
(``smc_cvnn_train_step``) and, by default, replays as ONE CUDA graph per training step — a step is
then: pinned H2D of the contracts, the simulation launches (their matrix index changes every
step, so they stay ordinary launches), two small device copies into the graph's static inputs,
one graph launch.  The losses stay on the device until ``train`` returns.  Any other network
(batch norms, residual blocks) takes the torch route (``fused_step=False`` forces it): autograd for the
network, ``smc_adam_step`` over the flattened parameters (``FlatAdam``), captured as one CUDA graph too;
``cuda_graph=False`` gives the reference's eager step with ``torch.optim.Adam``.  In both graphed routes the step
runs on a high-priority second stream, one batch behind the simulation.
"""

from __future__ import annotations

import copy
import math
import time
import warnings
from dataclasses import dataclass
from typing import Callable, Sequence

import numpy as np
import torch
from pydantic import BaseModel, ConfigDict
from torch import nn, optim

from spectralmc_b200.cvnn import FlatAdam, FusedCVNN, describe
from spectralmc_b200.distributed import PeerExchange, sharded_cf_targets
from spectralmc_b200 import _cabi
from spectralmc_b200.errors import (
    DeviceDTypeError,
    DeviceKernelFailed,
    DeviceNotCUDA,
    InvalidTrainingConfig,
    PredictionFailed,
    SamplerInitFailed,
)
from spectralmc_b200.gbm import BlackScholes, BlackScholesConfig
from spectralmc_b200.result import Failure, Result, Success
from spectralmc_b200.sobol_sampler import DomainBounds, SobolConfig, SobolSampler
from spectralmc_b200.validation import validate_model


def _split_inputs(rows: np.ndarray | Sequence[BlackScholes.Inputs], *, dtype: torch.dtype, device: torch.device) -> tuple[torch.Tensor, torch.Tensor]:
    """Contracts -> CVNN ``(real, imag)`` inputs in field order X0,K,T,r,d,v (reference :1775-1783)."""
    if not isinstance(rows, np.ndarray):
        fields = list(BlackScholes.Inputs.model_fields.keys())
        rows = np.array([[float(getattr(inp, f)) for f in fields] for inp in rows], dtype=np.float64)
    real = torch.as_tensor(rows, dtype=dtype, device=device)
    return real, torch.zeros_like(real)


# complex multiply-adds per training step (rows x sum of in x out over the linear layers) above which the graphed torch
# route is faster than the C-ABI step on a B200: 0.6x at 8.2e8, 1.04x at 1.2e9 and 2.7e9, 1.44x at 9.7e9
FUSED_STEP_MAX_MACS = 1_000_000_000


@dataclass(frozen=True)
class TrainingConfig:
    """Training hyper-parameters (reference :252-258); validate with ``build_training_config``."""

    num_batches: int
    batch_size: int
    learning_rate: float = 1e-2


def build_training_config(*, num_batches: int, batch_size: int, learning_rate: float) -> Result[TrainingConfig, InvalidTrainingConfig]:
    """Reference :261-298: positive counts, learning rate in (0, 1)."""

    def bad(message: str) -> Failure[InvalidTrainingConfig]:
        return Failure(InvalidTrainingConfig(num_batches=num_batches, batch_size=batch_size, learning_rate=learning_rate, message=message))

    if num_batches <= 0:
        return bad("num_batches must be > 0")
    if batch_size <= 0:
        return bad("batch_size must be > 0")
    if not (0.0 < learning_rate < 1.0):
        return bad("learning_rate must be in (0, 1)")
    return Success(TrainingConfig(num_batches=num_batches, batch_size=batch_size, learning_rate=learning_rate))


class GbmCVNNPricerConfig(BaseModel):
    """Frozen snapshot of a trainer (reference :301-313)."""

    cfg: BlackScholesConfig
    domain_bounds: DomainBounds
    cvnn: nn.Module
    optimizer_state: dict | None = None  # torch.optim.Adam.state_dict() layout, CPU tensors
    global_step: int = 0
    sobol_skip: int = 0
    torch_cpu_rng_state: bytes | None = None
    torch_cuda_rng_states: list[bytes] | None = None

    model_config = ConfigDict(arbitrary_types_allowed=True, frozen=True, extra="forbid")


def build_gbm_cvnn_pricer_config(**kwargs: object) -> Result[GbmCVNNPricerConfig, object]:
    return validate_model(GbmCVNNPricerConfig, **kwargs)


@dataclass(frozen=True)
class StepMetrics:
    """Scalars handed to the ``logger`` callback after every optimiser step (reference :336-346).
    ``optimizer`` is the ``torch.optim.Adam`` of the torch route or the ``FusedCVNN`` of the C-ABI route."""

    step: int
    batch_time: float
    loss: float
    grad_norm: float
    lr: float
    optimizer: object
    model: nn.Module


StepLogger = Callable[[StepMetrics], None]


@dataclass(frozen=True)
class TrainingResult:
    """Outcome of ``train`` (reference :349-361) plus the per-step losses, which this implementation
    keeps on the device during the run and reads back once."""

    updated_config: GbmCVNNPricerConfig
    final_loss: float
    total_batches: int
    final_grad_norm: float
    losses: tuple[float, ...] = ()


def _module_device_dtype(module: nn.Module) -> Result[tuple[torch.device, torch.dtype], DeviceDTypeError]:
    tensors = [t for t in module.state_dict().values() if t.is_floating_point()]  # BatchNorm keeps an int64 batch counter
    if not tensors:
        return Failure(DeviceDTypeError(message="CVNN has no parameters"))
    devices, dtypes = {t.device for t in tensors}, {t.dtype for t in tensors}
    if len(devices) != 1 or len(dtypes) != 1:
        return Failure(DeviceDTypeError(message=f"CVNN tensors span devices {sorted(map(str, devices))} / dtypes {sorted(map(str, dtypes))}"))
    dtype = dtypes.pop()
    if dtype not in (torch.float32, torch.float64):
        return Failure(DeviceDTypeError(message=f"CVNN must use float32 or float64 parameters, got {dtype}"))
    return Success((devices.pop(), dtype))


class GbmCVNNPricer:
    """Trains a CVNN on CF targets produced by the fused Monte-Carlo path."""

    @staticmethod
    def create(cfg: GbmCVNNPricerConfig, *, process_group=None, fused_step: bool | None = None, cuda_graph: bool = True,
               peer_exchange: bool = False) -> Result["GbmCVNNPricer", DeviceDTypeError | DeviceNotCUDA]:
        """Validated construction (reference :600-663): the CVNN must live on one CUDA device in the
        simulation's precision."""
        got = _module_device_dtype(cfg.cvnn)
        if isinstance(got, Failure):
            return got
        device, dtype = got.value
        if device.type != "cuda":
            return Failure(DeviceNotCUDA(device=str(device), message=f"Model on {device}, but CUDA required for training"))
        if dtype != cfg.cfg.sim_params.dtype.to_torch():
            return Failure(DeviceDTypeError(message=f"gbm sim dtype {cfg.cfg.sim_params.dtype} does not match cvnn dtype {dtype}"))
        self = GbmCVNNPricer(cfg.cfg, cfg.domain_bounds, cfg.cvnn, sobol_skip=cfg.sobol_skip, global_step=cfg.global_step,
                             process_group=process_group, fused_step=fused_step, cuda_graph=cuda_graph, peer_exchange=peer_exchange)
        self._optimizer_state = cfg.optimizer_state
        if cfg.torch_cpu_rng_state is not None:  # reference :714-721
            torch.set_rng_state(torch.from_numpy(np.frombuffer(cfg.torch_cpu_rng_state, dtype=np.uint8).copy()))
        if cfg.torch_cuda_rng_states is not None and len(cfg.torch_cuda_rng_states) == torch.cuda.device_count():
            torch.cuda.set_rng_state_all([torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()) for b in cfg.torch_cuda_rng_states])
        return Success(self)

    def __init__(self, cfg: BlackScholesConfig, domain_bounds: DomainBounds, cvnn: nn.Module, *, sobol_skip: int = 0,
                 global_step: int = 0, process_group=None, fused_step: bool | None = None, cuda_graph: bool = True,
                 peer_exchange: bool = False) -> None:
        self._cfg, self._sp = cfg, cfg.sim_params
        self._engine = BlackScholes(cfg)
        self._cvnn = cvnn
        self._dtype = self._sp.dtype.to_torch()
        self._device = self._engine._device
        self._domain_bounds = domain_bounds
        self._sobol_skip, self._global_step = sobol_skip, global_step
        self._group = process_group
        # multi-GPU RAW runs: fuse the all-reduce of the partial sums into the finalise kernel over peer memory
        self._want_exchange = peer_exchange and process_group is not None
        self._exchange: PeerExchange | None = None
        self._optimizer_state: dict | None = None  # Adam state between train() calls (reference :671, :1646)
        supported = describe(cvnn) is not None
        if fused_step and not supported:
            raise ValueError("fused_step=True needs a ComplexSequential of ComplexLinear / modReLU / zReLU")
        self._use_fused = supported if fused_step is None else fused_step
        # fused_step=None: per batch size, the C-ABI step below FUSED_STEP_MAX_MACS complex multiply-adds per step and
        # torch autograd (cuBLAS SGEMM) + smc_adam_step above it — measured crossover, profiles/r2_cvnn_widths.md
        self._route_by_size = fused_step is None
        self._use_graph = cuda_graph
        self._fused: FusedCVNN | None = None
        self._graphs: dict[int, _StepGraph] = {}
        self._torch_graphs: dict[int, _TorchStepGraph] = {}  # torch route, one captured step per batch size
        self._flat_adam: FlatAdam | None = None               # Adam of the graphed torch route (device step counter)
        self._staging: dict[tuple, list] = {}
        self._staging_next = 0
        self._nn_stream: torch.cuda.Stream | None = None  # CVNN steps run here, one step behind the simulation
        # the sampler is seeded with mc_seed and resumed with sobol_skip (reference :703-710)
        self._sampler_result = SobolSampler.create(BlackScholes.Inputs, domain_bounds, config=SobolConfig(seed=self._sp.mc_seed, skip=sobol_skip))

    # ------------------------------------------------------------------ data path
    def _simulate_fft(self, contract: BlackScholes.Inputs) -> Result[torch.Tensor, object]:
        """One contract -> ``[N]`` complex CF estimate (reference :806-817)."""
        return self._engine.simulate_fft(contract)

    def _upload(self, rows: np.ndarray) -> torch.Tensor:
        """Contracts -> device through a small ring of reusable pinned staging buffers (a slot is rewritten
        only after the copy that last read it has completed)."""
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        ring = self._staging.get(rows.shape)
        if ring is None:
            ring = self._staging[rows.shape] = [[torch.empty(rows.shape, dtype=torch.float64).pin_memory(), None] for _ in range(4)]
        slot = ring[self._staging_next % len(ring)]
        self._staging_next += 1
        if slot[1] is not None:
            slot[1].synchronize()
        slot[0].numpy()[...] = rows
        dev = slot[0].to(self._device, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        return dev

    def targets(self, rows: np.ndarray | torch.Tensor) -> Result[torch.Tensor, object]:
        """``[C, 6]`` contracts -> ``[C, N]`` complex targets on device (one C-ABI call sequence)."""
        dev = rows if isinstance(rows, torch.Tensor) else self._upload(rows)
        if self._group is not None:
            if self._want_exchange and (self._exchange is None or self._exchange.capacity_contracts < dev.shape[0]):
                if self._exchange is not None:
                    self._exchange.close()
                self._exchange = PeerExchange(int(dev.shape[0]), self._sp.network_size, group=self._group)  # collective set-up
            try:
                # "auto": whole contracts per rank + one all-gather when a batch-row shard would be too small to fill a GPU
                return Success(sharded_cf_targets(self._engine, dev, group=self._group, exchange=self._exchange, shard="auto"))
            except _cabi.SmcError as exc:  # same error ADT as the single-GPU path (BlackScholes.cf_targets)
                return Failure(DeviceKernelFailed(status=exc.code, message=exc.message))
        return self._engine.cf_targets(dev)

    def close(self) -> None:
        """Release the peer-exchange buffers (collective: every rank of the process group must call it).
        A pricer without a process group has nothing to release."""
        if self._exchange is not None:
            exchange, self._exchange = self._exchange, None
            exchange.close()

    def _torch_step(self, real_in, imag_in, targets, optimizer) -> tuple[torch.Tensor, torch.Tensor]:
        """Forward / MSE on real and imaginary parts / backward / Adam (reference :819-835); the
        gradient norm stays a device scalar (the reference converts it to a float every step)."""
        pred_r, pred_i = self._cvnn(real_in, imag_in)
        loss = nn.functional.mse_loss(pred_r, torch.real(targets)) + nn.functional.mse_loss(pred_i, torch.imag(targets))
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        optimizer.step()
        grad_norm = torch.nn.utils.clip_grad_norm_(self._cvnn.parameters(), float("inf"))
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
        if self._route_by_size and fused.complex_macs(rows) > FUSED_STEP_MAX_MACS:
            # wide network x large batch: the SIMT complex GEMM of the C-ABI step (~25 TFLOP/s) falls behind cuBLAS; the
            # same flat parameter / gradient / Adam buffers are stepped by the graphed torch route instead
            tg = self._torch_graphs.get(rows)
            if tg is None:
                tg = self._torch_graphs[rows] = _TorchStepGraph(self._cvnn, fused.adam, rows, contracts.shape[1], targets.shape[1], self._dtype)
            tg.real_in.copy_(contracts)
            tg.targets.copy_(targets)
            tg.graph.replay()
            loss_slot.copy_(tg.loss)
            return
        g = self._graphs.get(rows)
        if g is None:
            g = self._graphs[rows] = _StepGraph(fused, rows)
        g.real_in.copy_(contracts)  # float64 -> dtype, as torch.as_tensor(rows, dtype) narrows (reference :1779-1781)
        g.targets.copy_(targets)
        g.graph.replay()
        loss_slot.copy_(g.loss)

    def _attach_optimizer(self, learning_rate: float) -> object:
        """A fresh Adam per ``train`` call with the previous call's state re-attached (reference :1513-1529)."""
        if self._use_fused:
            if self._fused is None:
                self._fused = FusedCVNN(self._cvnn, lr=learning_rate)
            elif self._fused.hyper.lr != learning_rate:
                self._fused.hyper.lr = learning_rate  # hyper-parameters are baked into captured launches
                self._graphs.clear()
            if self._optimizer_state is not None:
                self._fused.load_optimizer_state_dict(dict(self._optimizer_state, param_groups=[
                    dict(self._optimizer_state["param_groups"][0], lr=learning_rate)]))
                self._optimizer_state = None  # now lives in the fused buffers until the next snapshot
            return self._fused
        if self._use_graph:
            # graphed torch route: autograd for the network, smc_adam_step over the flattened parameters (capturable,
            # float64 bias corrections like the eager optimiser; torch's own capturable Adam keeps `step` in float32)
            if self._flat_adam is None:
                self._flat_adam = FlatAdam(list(self._cvnn.parameters()), lr=learning_rate)
            elif self._flat_adam.hyper.lr != learning_rate:
                self._flat_adam.hyper.lr = learning_rate  # baked into the captured launch
                self._torch_graphs.clear()
            if self._optimizer_state is not None:
                self._flat_adam.load_state_dict(_with_lr(self._optimizer_state, learning_rate))
                self._optimizer_state = None
            return self._flat_adam
        adam = optim.Adam(self._cvnn.parameters(), lr=learning_rate)
        if self._optimizer_state is not None:
            adam.load_state_dict(_with_lr(self._optimizer_state, learning_rate))
        return adam

    def _torch_graph_step(self, contracts: torch.Tensor, targets: torch.Tensor, loss_slot: torch.Tensor) -> torch.Tensor:
        """The torch route of ``_torch_step`` replayed as one CUDA graph (networks the fused step does not cover)."""
        rows = contracts.shape[0]
        g = self._torch_graphs.get(rows)
        if g is None:
            g = self._torch_graphs[rows] = _TorchStepGraph(self._cvnn, self._flat_adam, rows, contracts.shape[1], targets.shape[1], self._dtype)
        g.real_in.copy_(contracts)
        g.targets.copy_(targets)
        g.graph.replay()
        loss_slot.copy_(g.loss)
        return g.grad_norm

    def train(self, config: TrainingConfig, *, logger: StepLogger | None = None) -> Result[TrainingResult, object]:
        """``config.num_batches`` optimiser steps (reference :1456-1683).  Without a ``logger`` nothing
        synchronises the host until the losses are read back at the end."""
        if isinstance(self._sampler_result, Failure):
            return Failure(SamplerInitFailed(error=self._sampler_result.error))
        sampler = self._sampler_result.value
        optimizer = self._attach_optimizer(config.learning_rate)
        self._cvnn.train()
        losses = torch.zeros(max(config.num_batches, 1), dtype=torch.float64, device=self._device)
        grad_norm: torch.Tensor | None = None
        # Software pipeline (fused route, no per-step logger): the CVNN step of batch i is issued on a second
        # stream and overlaps the simulation of batch i + 1 — the two are independent, and the step's small
        # launches fit into the SM slots the tile kernel frees.  Same kernels in the same per-stream order, so
        # results are bit-identical to the serial schedule.
        graphed_torch = not self._use_fused and self._use_graph
        pipelined = (self._use_fused or graphed_torch) and logger is None
        sim_stream = torch.cuda.current_stream(self._device)
        if pipelined:
            if self._nn_stream is None:
                # high priority: the step's small CTAs take SM slots as soon as the simulation kernel frees any
                self._nn_stream = torch.cuda.Stream(self._device, priority=-1)
            self._nn_stream.wait_stream(sim_stream)  # parameters / optimiser state produced so far
        for i in range(config.num_batches):
            t0 = time.perf_counter()
            drawn = sampler.sample_array(config.batch_size)
            if isinstance(drawn, Failure):
                return Failure(SamplerInitFailed(error=drawn.error))
            self._sobol_skip += config.batch_size
            contracts = self._upload(drawn.value)
            with _nvtx("smc.targets"):  # visible in ncu / nsys timelines (the reference has no profiler hooks)
                got = self.targets(contracts)
            if isinstance(got, Failure):
                return got
            targets = got.value.detach()  # already a torch tensor: the DLPack hand-off is the identity
            with _nvtx("smc.cvnn_step"):
                if pipelined:
                    ready = torch.cuda.Event()
                    ready.record(sim_stream)
                    with torch.cuda.stream(self._nn_stream):
                        self._nn_stream.wait_event(ready)
                        contracts.record_stream(self._nn_stream)  # allocated on the simulation stream, read here
                        targets.record_stream(self._nn_stream)
                        if self._use_fused:
                            self._fused_step(contracts, targets, losses[i : i + 1])
                        else:
                            grad_norm = self._torch_graph_step(contracts, targets, losses[i : i + 1])
                elif self._use_fused:
                    self._fused_step(contracts, targets, losses[i : i + 1])
                elif graphed_torch:
                    grad_norm = self._torch_graph_step(contracts, targets, losses[i : i + 1])
                else:
                    real_in = contracts.to(self._dtype)
                    loss, grad_norm = self._torch_step(real_in, torch.zeros_like(real_in), targets, optimizer)
                    losses[i : i + 1].copy_(loss.detach())
            self._global_step += 1
            if logger is not None:  # per-step host metrics, as the reference produces them (:1567-1583)
                gn = self._fused.grads.norm() if self._use_fused else grad_norm
                logger(StepMetrics(step=self._global_step, batch_time=time.perf_counter() - t0, loss=float(losses[i]),
                                   grad_norm=float(gn), lr=config.learning_rate, optimizer=optimizer, model=self._cvnn))
        if pipelined:
            sim_stream.wait_stream(self._nn_stream)
        if config.num_batches > 0:
            grad_norm = self._fused.grads.norm() if self._use_fused else grad_norm
        host = [float(x) for x in losses[: config.num_batches].cpu()]
        if not self._use_fused and not self._use_graph:
            self._optimizer_state = _to_cpu(optimizer.state_dict())
        snap = self.snapshot()
        if isinstance(snap, Failure):
            return snap
        return Success(TrainingResult(updated_config=snap.value, final_loss=host[-1] if host else 0.0, total_batches=config.num_batches,
                                      final_grad_norm=float(grad_norm) if grad_norm is not None else 0.0, losses=tuple(host)))

    def predict_price(self, inputs: Sequence[BlackScholes.Inputs]) -> Result[list[BlackScholes.HostPricingResults], PredictionFailed]:
        """CVNN forward -> mean of the inverse DFT -> put price; call by parity (reference :1709-1769).
        ``mean_n ifft(S)[n] == S[0] / N``, so the fused route needs no transform."""
        if len(inputs) == 0:
            return Success([])
        self._cvnn.eval()
        real_in, imag_in = _split_inputs(inputs, dtype=self._dtype, device=self._device)
        try:
            if self._use_fused:
                if self._fused is None:
                    self._fused = FusedCVNN(self._cvnn)
                pred_r, pred_i = self._fused.forward(real_in, imag_in)
                n = pred_r.shape[1]
                coeffs = torch.stack((pred_r[:, 0] / n, pred_i[:, 0] / n), dim=1).cpu().tolist()
            else:
                with torch.no_grad():
                    pred_r, pred_i = self._cvnn(real_in, imag_in)
                    avg = torch.fft.ifft(torch.complex(pred_r, pred_i), dim=1).mean(dim=1)
                coeffs = torch.stack((avg.real, avg.imag), dim=1).cpu().tolist()
        except Exception as exc:  # noqa: BLE001 - mirrored from the reference (:1768)
            return Failure(PredictionFailed(message=str(exc)))
        out = []
        for (real_val, imag_val), c in zip(coeffs, inputs):
            if abs(imag_val) > 1.0e-6:
                warnings.warn(f"IFFT imaginary component {imag_val:.3e} exceeds tolerance.", RuntimeWarning)
            discount = math.exp(-c.r * c.T)
            forward = c.X0 * math.exp((c.r - c.d) * c.T)
            put, call = real_val, real_val + forward - c.K * discount
            put_i, call_i = discount * max(c.K - forward, 0.0), discount * max(forward - c.K, 0.0)
            out.append(BlackScholes.HostPricingResults(underlying=forward, put_price=put, call_price=call, put_price_intrinsic=put_i,
                                                       call_price_intrinsic=call_i, put_convexity=put - put_i, call_convexity=call - call_i))
        return Success(out)

    # ------------------------------------------------------------------ checkpointing
    def _current_optimizer_state(self) -> dict | None:
        """Adam state in ``torch.optim.Adam.state_dict()`` layout (CPU) for both step routes."""
        if self._use_fused and self._fused is not None and int(self._fused.step.item()) > 0:
            return _to_cpu(self._fused.optimizer_state_dict())
        if self._flat_adam is not None and int(self._flat_adam.step.item()) > 0:
            return _to_cpu(self._flat_adam.state_dict())
        return self._optimizer_state

    def snapshot(self) -> Result[GbmCVNNPricerConfig, object]:
        """Deterministic snapshot (reference :756-800): engine config with the advanced ``skip``,
        Sobol position, step count, optimiser state, torch RNG states; ``cvnn`` is the live module
        (clone it before training on, as the reference's tests do)."""
        cfg = self._engine.snapshot()
        if isinstance(cfg, Failure):
            return cfg
        cuda_rng = [s.cpu().numpy().tobytes() for s in torch.cuda.get_rng_state_all()] if torch.cuda.is_available() else None
        return Success(GbmCVNNPricerConfig(cfg=cfg.value, domain_bounds=self._domain_bounds, cvnn=self._cvnn,
                                           optimizer_state=self._current_optimizer_state(), global_step=self._global_step,
                                           sobol_skip=self._sobol_skip, torch_cpu_rng_state=torch.get_rng_state().numpy().tobytes(),
                                           torch_cuda_rng_states=cuda_rng))


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
        # captured on a high-priority stream: kernel nodes keep that priority, so a replay that overlaps the next
        # simulation gets SM slots as soon as the big kernel frees any (see GbmCVNNPricer.train)
        with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(dev, priority=-1)):
            fused.train_step(self.real_in, self.imag_in, self.targets, self.loss)


def _with_lr(state: dict, lr) -> dict:
    """``state`` (Adam ``state_dict`` layout) with every group's learning rate replaced."""
    return dict(state, param_groups=[dict(g, lr=lr) for g in state["param_groups"]])


class _TorchStepGraph:
    """One captured ``_torch_step`` of an arbitrary torch CVNN for a fixed batch size: forward, two MSE terms and
    backward by torch autograd (gradients accumulate into the views of ``FlatAdam``'s zeroed gradient buffer),
    then one ``smc_adam_step`` and the gradient norm.  cuBLAS handles and workspaces are initialised by one
    forward/backward on a deep copy of the network, so the real parameters and statistics are untouched before
    the capture."""

    def __init__(self, net: nn.Module, adam: FlatAdam, rows: int, n_inputs: int, n_outputs: int, dtype: torch.dtype) -> None:
        dev = adam.device
        self.real_in = torch.zeros((rows, n_inputs), dtype=dtype, device=dev)
        self.imag_in = torch.zeros_like(self.real_in)
        self.targets = torch.zeros((rows, n_outputs), dtype=torch.complex64 if dtype == torch.float32 else torch.complex128, device=dev)
        self.loss = torch.zeros((), dtype=torch.float64, device=dev)
        self.grad_norm = torch.zeros((), dtype=dtype, device=dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            twin = copy.deepcopy(net)
            pr, pi = twin(torch.randn_like(self.real_in), self.imag_in)
            (pr.square().mean() + pi.square().mean()).backward()
            scratch = [torch.zeros(1, dtype=dtype, device=dev) for _ in range(4)]
            _adam_warm_up(scratch, adam)
            del twin, pr, pi
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(dev, priority=-1)):
            pred_r, pred_i = net(self.real_in, self.imag_in)
            loss = nn.functional.mse_loss(pred_r, torch.real(self.targets)) + nn.functional.mse_loss(pred_i, torch.imag(self.targets))
            adam.grads.zero_()
            loss.backward()
            adam.apply()
            self.grad_norm.copy_(adam.grads.norm())
            self.loss.copy_(loss.detach())


def _adam_warm_up(scratch: list[torch.Tensor], adam: FlatAdam) -> None:
    """Load the Adam kernels outside a capture, on scratch buffers."""

    _cabi.adam_step(scratch[0], scratch[1], scratch[2], scratch[3], torch.zeros(1, dtype=torch.int64, device=adam.device), adam.hyper)


class _nvtx:
    """NVTX range around a phase of the training step."""

    def __init__(self, name: str) -> None:
        self.name = name

    def __enter__(self) -> None:
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc) -> None:
        torch.cuda.nvtx.range_pop()


def _to_cpu(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().clone()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj
