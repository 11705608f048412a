"""GBM / Black-Scholes Monte-Carlo engine — drop-in for the reference's ``spectralmc.gbm``.

Keeps the reference's API (/root/reference/src/spectralmc/gbm.py): ``SimulationParams`` :77-103,
``validate_simulation_params_memory`` :106-137, ``BlackScholesConfig`` :143-161,
``build_simulation_params`` / ``build_black_scholes_config`` :164-217, the launch object
``SimulateBlackScholes[blocks, threads, stream](io, timesteps, dt, X0, r, d, v, log_flag)``
:224-257, and ``BlackScholes`` with ``Inputs`` / ``SimResults`` / ``PricingResults`` /
``HostPricingResults``, ``_simulate`` :400, ``price`` :450, ``get_host_price`` :491,
``price_to_host`` :515, ``snapshot`` :332 — and routes every device operation through the C ABI
(``spectralmc_b200._cabi``): no Numba, no CuPy, no CPU fallback.

Deviations, all forced by "no CuPy on this path": device arrays in the result models are
``torch.Tensor`` (DLPack-native, so ``torch.from_dlpack`` consumers keep working) and stream
arguments are ``torch.cuda.Stream``.

New on top of the reference API: ``BlackScholes.cf_targets(contracts)`` — the fused batch
path (one launch sequence for a whole Sobol batch, normals never materialised) that replaces the
trainer's per-contract loop (gbm_trainer.py:1546-1553).
"""

from __future__ import annotations

import weakref
from math import exp
from typing import Annotated, Literal, Sequence, TypeAlias

import numpy as np
import torch
from pydantic import BaseModel, ConfigDict, Field

from spectralmc_b200 import _cabi
from spectralmc_b200.async_normals import BufferConfig, ConcurrentNormGenerator, ConcurrentNormGeneratorConfig
from spectralmc_b200.effects import ForwardNormalization, GenerateNormals, PathScheme, SimulatePaths
from spectralmc_b200.errors import (
    DeviceKernelFailed,
    GPUMemoryLimitExceeded,
    InvalidBlackScholesConfig,
    InvalidSimulationParams,
    NormalsGenerationFailed,
    NormalsUnavailable,
)
from spectralmc_b200.numerical import Precision
from spectralmc_b200.result import Failure, Result, Success
from spectralmc_b200.validation import validate_model

PosFloat = Annotated[float, Field(gt=0)]
NonNegFloat = Annotated[float, Field(ge=0)]
ThreadsPerBlock: TypeAlias = Literal[32, 64, 128, 256, 512, 1024]
NormalsError: TypeAlias = NormalsUnavailable | NormalsGenerationFailed | DeviceKernelFailed

_SCHEME_CODE = {PathScheme.LOG_EULER: _cabi.SMC_LOG_EULER, PathScheme.SIMPLE_EULER: _cabi.SMC_SIMPLE_EULER}
_NORM_CODE = {ForwardNormalization.NORMALIZE: _cabi.SMC_NORMALIZE, ForwardNormalization.RAW: _cabi.SMC_RAW}


class SimulationParams(BaseModel):
    """Immutable run-time parameters of one engine (reference gbm.py:77-103)."""

    timesteps: int = Field(..., gt=0)
    network_size: int = Field(..., gt=0)
    batches_per_mc_run: int = Field(..., gt=0)
    threads_per_block: ThreadsPerBlock
    mc_seed: int = Field(..., gt=0)
    buffer_size: int = Field(..., gt=0)
    skip: int = Field(0, ge=0)
    dtype: Precision
    # NOT in the reference: which counter-based generator draws the normals.  0 = Philox4x32-10, the default and the
    # stream every published number is measured on; 1 = the opt-in Philox4x32-7 stream (about 13 % faster in the fused
    # float32 kernel, passes the same battery, a DIFFERENT sample set).  include/spectralmc_b200.h: smc_stream_version.
    stream_version: Literal[0, 1] = 0

    model_config = ConfigDict(frozen=True, extra="forbid")

    def total_paths(self) -> int:
        return self.network_size * self.batches_per_mc_run

    def total_blocks(self) -> int:
        return (self.total_paths() + self.threads_per_block - 1) // self.threads_per_block


def validate_simulation_params_memory(params: SimulationParams) -> Result[SimulationParams, GPUMemoryLimitExceeded]:
    """Soft cap on total paths (reference gbm.py:106-137: 1e9 for float32, 5e8 for float64)."""
    total = params.total_paths()
    limit = 1_000_000_000 if params.dtype.value == "float32" else 500_000_000
    if total > limit:
        return Failure(
            GPUMemoryLimitExceeded(
                total_paths=total,
                max_paths=limit,
                network_size=params.network_size,
                batches_per_mc_run=params.batches_per_mc_run,
            )
        )
    return Success(params)


class BlackScholesConfig(BaseModel):
    """Frozen engine configuration (reference gbm.py:143-161)."""

    sim_params: SimulationParams
    path_scheme: PathScheme = PathScheme.LOG_EULER
    normalization: ForwardNormalization = ForwardNormalization.NORMALIZE

    model_config = ConfigDict(frozen=True, extra="forbid")


def build_simulation_params(
    *,
    timesteps: int,
    network_size: int,
    batches_per_mc_run: int,
    threads_per_block: ThreadsPerBlock,
    mc_seed: int,
    buffer_size: int,
    dtype: Precision,
    skip: int = 0,
    stream_version: int = 0,
) -> Result[SimulationParams, InvalidSimulationParams | GPUMemoryLimitExceeded]:
    made = validate_model(
        SimulationParams,
        timesteps=timesteps,
        network_size=network_size,
        batches_per_mc_run=batches_per_mc_run,
        threads_per_block=threads_per_block,
        mc_seed=mc_seed,
        buffer_size=buffer_size,
        skip=skip,
        dtype=dtype,
        stream_version=stream_version,
    )
    if isinstance(made, Failure):
        return Failure(InvalidSimulationParams(error=made.error))
    return validate_simulation_params_memory(made.value)


def build_black_scholes_config(
    *,
    sim_params: SimulationParams,
    path_scheme: PathScheme = PathScheme.LOG_EULER,
    normalization: ForwardNormalization = ForwardNormalization.NORMALIZE,
) -> Result[BlackScholesConfig, InvalidBlackScholesConfig]:
    made = validate_model(BlackScholesConfig, sim_params=sim_params, path_scheme=path_scheme, normalization=normalization)
    return made if isinstance(made, Success) else Failure(InvalidBlackScholesConfig(error=made.error))


# ───────────────────────────── kernel launch object ─────────────────────────────
class _Launch:
    def __init__(self, blocks: int, threads: int, stream: object | None) -> None:
        # `blocks` is accepted for call compatibility and not used: the grid follows io.shape[1] and the
        # kernel's vector width (the reference's own total_blocks() is ceil(paths / threads), gbm.py:101-103)
        self._threads, self._stream = threads, stream

    def __call__(
        self, io: object, timesteps: int, dt: float, X0: float, r: float, d: float, v: float, simulate_log_return: bool
    ) -> None:
        _ptr, shape, _dtype, _keep = _cabi.device_matrix(io, "io")
        if len(shape) != 2 or shape[0] != timesteps:
            raise ValueError(f"io must have shape (timesteps={timesteps}, paths); got {tuple(shape)}")
        scheme = _cabi.SMC_LOG_EULER if simulate_log_return else _cabi.SMC_SIMPLE_EULER
        _cabi.gbm_paths_inplace(io, dt, X0, r, d, v, scheme, self._threads, stream=self._stream)


class _SimulateBlackScholes:
    """``SimulateBlackScholes[blocks, threads, stream](io, timesteps, dt, X0, r, d, v, log_flag)``.

    Same call shape as the reference's Numba kernel object (gbm.py:224-257, launched at
    gbm.py:413-426 and effects/interpreter.py:645-654), and it accepts what the reference passes there:
    ``io`` is a torch CUDA tensor, anything with ``__cuda_array_interface__`` (Numba's
    ``cuda.as_cuda_array(sims)``, a CuPy array) or a DLPack exporter — zero-copy, mutated in place;
    ``stream`` is a ``torch.cuda.Stream``, a Numba stream (``.handle``), a CuPy stream (``.ptr``), a raw
    ``cudaStream_t`` integer, or absent (torch's current stream).  ``blocks`` is accepted and ignored: the
    grid is derived from ``io.shape[1]`` and the vector width; ``threads`` is the CTA size.
    """

    def __getitem__(self, cfg: tuple) -> _Launch:
        blocks, threads = cfg[0], cfg[1]
        stream = cfg[2] if len(cfg) > 2 else None
        return _Launch(blocks, threads, stream)


SimulateBlackScholes = _SimulateBlackScholes()


# ─────────────────────────────────── engine ───────────────────────────────────
class BlackScholes:
    """Single-GPU Monte-Carlo pricing engine (reference gbm.py:263-521)."""

    class Inputs(BaseModel):
        """One European option contract (reference gbm.py:267-277)."""

        X0: PosFloat
        K: PosFloat
        T: NonNegFloat
        r: float
        d: float
        v: NonNegFloat

        model_config = ConfigDict(frozen=True, extra="forbid")

    class SimResults(BaseModel):
        model_config = ConfigDict(arbitrary_types_allowed=True, extra="forbid")
        times: torch.Tensor
        sims: torch.Tensor
        forwards: torch.Tensor
        df: torch.Tensor

    class PricingResults(BaseModel):
        model_config = ConfigDict(arbitrary_types_allowed=True, extra="forbid")
        put_price_intrinsic: torch.Tensor
        call_price_intrinsic: torch.Tensor
        underlying: torch.Tensor
        put_price: torch.Tensor
        call_price: torch.Tensor

    class HostPricingResults(BaseModel):
        put_price_intrinsic: float
        call_price_intrinsic: float
        underlying: float
        put_convexity: float
        call_convexity: float
        put_price: float
        call_price: float

        model_config = ConfigDict(frozen=True, extra="forbid")

    def __init__(self, cfg: BlackScholesConfig) -> None:
        self._cfg = cfg
        self._sp = cfg.sim_params
        self._dtype = self._sp.dtype.to_torch()
        self._np_dtype = self._sp.dtype.to_numpy()
        # resolved lazily enough that the host-side logic (argument packing, sharding, snapshots) can
        # be exercised on a CPU-only box; any device operation there fails loudly in torch/_cabi
        self._device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self._served = self._sp.skip  # matrices of the normal stream consumed so far
        ngen_cfg = ConcurrentNormGeneratorConfig.create(
            rows=self._sp.timesteps, cols=self._sp.total_paths(), seed=self._sp.mc_seed, dtype=self._sp.dtype, skips=self._sp.skip,
            stream_version=self._sp.stream_version,
        )
        if isinstance(ngen_cfg, Failure):
            raise AssertionError(f"Invalid norm generator config: {ngen_cfg.error}")
        # The materialised pool is created on first use: the fused path never needs it, and at
        # production sizes one matrix is gigabytes.
        self._ngen: Result[ConcurrentNormGenerator, object] | None = None
        self._host_cache: tuple | None = None  # (weakref sims, weakref forwards, F, df_last) of the last _simulate
        self._workspace: torch.Tensor | None = None

    # ----------------------------------------------------------------- normal supply
    def _generator(self) -> Result[ConcurrentNormGenerator, NormalsUnavailable]:
        if self._ngen is None:
            cfg = ConcurrentNormGeneratorConfig.create(
                rows=self._sp.timesteps, cols=self._sp.total_paths(), seed=self._sp.mc_seed, dtype=self._sp.dtype, skips=self._served,
                stream_version=self._sp.stream_version,
            )
            buf = BufferConfig.create(self._sp.buffer_size, self._sp.timesteps, self._sp.total_paths())
            self._ngen = cfg if isinstance(cfg, Failure) else ConcurrentNormGenerator.create(buf, cfg.value)
        if isinstance(self._ngen, Failure):
            return Failure(NormalsUnavailable(error=self._ngen.error))
        return self._ngen

    def snapshot(self) -> Result[BlackScholesConfig, NormalsUnavailable]:
        """Deep copy of the configuration with ``skip`` = matrices consumed (reference gbm.py:332-339)."""
        if isinstance(self._ngen, Failure):
            return Failure(NormalsUnavailable(error=self._ngen.error))
        sp = self._sp.model_copy(update={"skip": self._served}, deep=True)
        return Success(self._cfg.model_copy(update={"sim_params": sp}, deep=True))

    def build_simulation_effects(self, inputs: "BlackScholes.Inputs") -> Result[tuple[object, ...], NormalsUnavailable]:
        """Pure description of one simulation as operator ADTs (reference gbm.py:342-397)."""
        return Success(
            (
                GenerateNormals(rows=self._sp.timesteps, cols=self._sp.total_paths(), seed=self._sp.mc_seed, skip=self._served),
                SimulatePaths(
                    spot=inputs.X0,
                    strike=inputs.K,
                    rate=inputs.r,
                    dividend=inputs.d,
                    vol=inputs.v,
                    expiry=inputs.T,
                    timesteps=self._sp.timesteps,
                    batches=self._sp.total_paths(),
                    path_scheme=self._cfg.path_scheme,
                    normalization=self._cfg.normalization,
                    input_normals_id="generated_normals",
                ),
            )
        )

    # ----------------------------------------------------------------- materialised path
    def _simulate(self, inputs: "BlackScholes.Inputs") -> Result["BlackScholes.SimResults", NormalsError]:
        """Full path matrix, as the reference returns it (gbm.py:400-447)."""
        gen = self._generator()
        if isinstance(gen, Failure):
            return gen
        drawn = gen.value.get_matrix()
        if isinstance(drawn, Failure):
            return Failure(NormalsGenerationFailed(error=drawn.error))
        sims = drawn.value
        self._served += 1
        dt = inputs.T / self._sp.timesteps
        try:
            SimulateBlackScholes[self._sp.total_blocks(), self._sp.threads_per_block](
                sims, self._sp.timesteps, dt, inputs.X0, inputs.r, inputs.d, inputs.v, self._cfg.path_scheme is PathScheme.LOG_EULER
            )
            # gbm.py:429-431 — length-`timesteps` vectors in the engine dtype.  Formed on the host
            # (three tiny NumPy expressions instead of five device launches) and copied once.
            times_h = np.linspace(dt, inputs.T, self._sp.timesteps, dtype=self._np_dtype)
            forwards_h = (inputs.X0 * np.exp((inputs.r - inputs.d) * times_h)).astype(self._np_dtype)
            df_h = np.exp(-inputs.r * times_h).astype(self._np_dtype)
            packed = torch.from_numpy(np.stack([times_h, forwards_h, df_h])).to(self._device)
            times, forwards, df = packed[0], packed[1], packed[2]
            if self._cfg.normalization is ForwardNormalization.NORMALIZE:
                _cabi.normalize_rows(sims, forwards.contiguous())  # gbm.py:437-438
        except _cabi.SmcError as exc:
            return Failure(DeviceKernelFailed(status=exc.code, message=exc.message))
        # forwards[-1] / df[-1] on the host for price(), keyed on the IDENTITY of the tensors just made (device
        # addresses are reused by the caching allocator, so they cannot tell two SimResults apart)
        self._host_cache = (weakref.ref(sims), weakref.ref(forwards), float(forwards_h[-1]), float(df_h[-1]))
        made = validate_model(self.SimResults, times=times, sims=sims, forwards=forwards, df=df)
        if isinstance(made, Failure):
            raise AssertionError(f"SimResults validation failed: {made.error}")
        return made

    def price(
        self, *, inputs: "BlackScholes.Inputs", sr_result: Result["BlackScholes.SimResults", NormalsError] | None = None
    ) -> Result["BlackScholes.PricingResults", NormalsError]:
        """Per-path discounted payoffs (reference gbm.py:450-488); no host synchronisation."""
        sim_result = sr_result or self._simulate(inputs)
        if isinstance(sim_result, Failure):
            return sim_result
        sr = sim_result.value
        if self._host_cache is not None and self._host_cache[0]() is sr.sims and self._host_cache[1]() is sr.forwards:
            F_h, df_h = self._host_cache[2], self._host_cache[3]
        else:  # foreign SimResults: one small device->host read
            F_h, df_h = float(sr.forwards[-1].item()), float(sr.df[-1].item())
        np_t = self._np_dtype.type
        K_h = np_t(inputs.K)  # gbm.py:467
        put_intr = torch.tensor(np_t(df_h) * max(K_h - np_t(F_h), np_t(0)), dtype=self._dtype, device=self._device)  # :469
        call_intr = torch.tensor(np_t(df_h) * max(np_t(F_h) - K_h, np_t(0)), dtype=self._dtype, device=self._device)  # :470
        terminal = sr.sims[-1]  # gbm.py:472
        try:
            put_price, call_price = _cabi.payoff(terminal, float(inputs.K), df_h)  # gbm.py:473-474
        except _cabi.SmcError as exc:
            return Failure(DeviceKernelFailed(status=exc.code, message=exc.message))
        made = validate_model(
            self.PricingResults,
            put_price_intrinsic=put_intr,
            call_price_intrinsic=call_intr,
            underlying=terminal,
            put_price=put_price,
            call_price=call_price,
        )
        if isinstance(made, Failure):
            raise AssertionError(f"PricingResults validation failed: {made.error}")
        return made

    def get_host_price(self, pr: "BlackScholes.PricingResults") -> "BlackScholes.HostPricingResults":
        """Scalar host prices (reference gbm.py:491-513) from ONE fused reduction + one D2H copy."""
        means = _cabi.means3(pr.underlying.contiguous(), pr.put_price, pr.call_price)
        scalars = torch.cat([means, pr.put_price_intrinsic.reshape(1).double(), pr.call_price_intrinsic.reshape(1).double()]).cpu()
        underlying, put_price, call_price, put_intr, call_intr = (float(x) for x in scalars)
        made = validate_model(
            self.HostPricingResults,
            put_price_intrinsic=put_intr,
            call_price_intrinsic=call_intr,
            underlying=underlying,
            put_convexity=put_price - put_intr,
            call_convexity=call_price - call_intr,
            put_price=put_price,
            call_price=call_price,
        )
        if isinstance(made, Failure):
            raise AssertionError(f"HostPricingResults validation failed: {made.error}")
        return made.value

    def price_to_host(self, inputs: "BlackScholes.Inputs") -> Result["BlackScholes.HostPricingResults", NormalsError]:
        priced = self.price(inputs=inputs)
        return priced if isinstance(priced, Failure) else Success(self.get_host_price(priced.value))

    # ----------------------------------------------------------------- fused batch path
    def fused_args(
        self,
        contracts_dev: torch.Tensor | None,
        n_contracts: int,
        *,
        batch_begin: int = 0,
        batch_end: int | None = None,
        scheme: int | None = None,
        matrix_offset: int = 0,
    ) -> _cabi.FusedArgs:
        """``smc_fused_args`` for ``n_contracts`` matrices of this engine's stream, starting ``matrix_offset`` after the
        next unconsumed one (a rank that handles contracts [c0, c1) of a batch passes ``matrix_offset=c0``)."""
        return _cabi.make_fused_args(
            contracts_dev,
            n_contracts,
            self._sp.timesteps,
            self._sp.network_size,
            self._sp.batches_per_mc_run,
            self._dtype,
            _SCHEME_CODE[self._cfg.path_scheme] if scheme is None else scheme,
            _NORM_CODE[self._cfg.normalization],
            self._sp.mc_seed,
            self._served + matrix_offset,
            batch_begin,
            batch_end,
            self._sp.stream_version,
        )

    def consume(self, n_matrices: int) -> None:
        """Advance the normal stream by ``n_matrices`` (each priced contract consumes one, gbm.py:405)."""
        self._served += n_matrices
        if isinstance(self._ngen, Success):
            self._ngen.value.skip(n_matrices)

    def contract_rows(self, contracts: Sequence["BlackScholes.Inputs"] | torch.Tensor) -> torch.Tensor:
        """``[C, 6]`` float64 contiguous rows X0,K,T,r,d,v on the engine's device — what the kernels read.
        Every fused entry point (single-GPU and sharded) coerces through here, so a float32, CPU or strided
        tensor is converted instead of being read as float64 rows."""
        if isinstance(contracts, torch.Tensor):
            if contracts.dim() != 2 or contracts.shape[1] != 6:
                raise ValueError(f"contracts must have shape [C, 6] (X0, K, T, r, d, v); got {tuple(contracts.shape)}")
            return contracts.to(device=self._device, dtype=torch.float64).contiguous()
        host = torch.tensor([[c.X0, c.K, c.T, c.r, c.d, c.v] for c in contracts], dtype=torch.float64).reshape(-1, 6)
        return host.pin_memory().to(self._device, non_blocking=True)

    def cf_targets(self, contracts: Sequence["BlackScholes.Inputs"] | torch.Tensor) -> Result[torch.Tensor, NormalsError]:
        """CF training targets ``[C, N]`` (complex) for a batch of contracts, fully on device.

        Equivalent to ``stack([mean(fft(price(c).put_price.reshape(B, N), axis=1), axis=0) for c in
        contracts])`` (reference gbm_trainer.py:1546-1553, 806-817) with the normals drawn in
        registers.  ``contracts`` may be a ``[C, 6]`` float64 CUDA tensor (columns X0,K,T,r,d,v).
        """
        rows = self.contract_rows(contracts)
        n = rows.shape[0]
        if n == 0:
            return Success(torch.empty((0, self._sp.network_size), dtype=_cabi.complex_dtype(self._dtype), device=self._device))
        args = self.fused_args(rows, n)
        need = _cabi.LIB.smc_cf_fused_workspace_bytes(_cabi.byref(args))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self._device)
        try:
            out = _cabi.cf_fused(args, self._device, self._dtype, self._workspace)
        except _cabi.SmcError as exc:
            return Failure(DeviceKernelFailed(status=exc.code, message=exc.message))
        self.consume(n)
        return Success(out)

    def simulate_fft(self, contract: "BlackScholes.Inputs") -> Result[torch.Tensor, NormalsError]:
        """One contract's CF estimate ``[N]`` (reference ``_simulate_fft``, gbm_trainer.py:806-817)."""
        got = self.cf_targets([contract])
        return got if isinstance(got, Failure) else Success(got.value[0])


def analytic_forward_df(inputs: BlackScholes.Inputs) -> tuple[float, float]:
    """``forwards[-1]`` and ``df[-1]`` in float64 (gbm.py:430-431 evaluated at t = T)."""
    return inputs.X0 * exp((inputs.r - inputs.d) * inputs.T), exp(-inputs.r * inputs.T)


__all__ = (
    "BlackScholes",
    "BlackScholesConfig",
    "SimulateBlackScholes",
    "SimulationParams",
    "ThreadsPerBlock",
    "build_black_scholes_config",
    "build_simulation_params",
    "validate_simulation_params_memory",
)
