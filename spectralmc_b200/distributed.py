"""Multi-GPU sharding of the fused batch path: one process per GPU, batch rows partitioned across
ranks, ONE sum-allreduce of the complex partial CF sums (plus one tiny allreduce of terminal-price
sums in NORMALIZE mode, because the payoff is non-linear in the normalisation factor).

The reference is single-GPU by policy (models/torch.py:160); this is new.  Every
rank simulates ALL contracts over batch rows ``[begin, end)`` of the GLOBAL ``batches_per_mc_run``;
Philox counters use the global path index, so the union over ranks is the single-GPU sample set and
results are independent of the GPU count up to summation order.

The device work goes through ``spectralmc_b200._cabi``; the collective is
``torch.distributed.all_reduce`` (NCCL over NVLink on GPUs).  ``ops`` exists so the CPU test-suite
can exercise the sharding/collective logic under ``gloo`` with a stand-in for the device calls.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Protocol

import torch
import torch.distributed as dist

from spectralmc_b200 import _cabi
from spectralmc_b200.effects import ForwardNormalization
from spectralmc_b200.gbm import BlackScholes


@dataclass(frozen=True)
class BatchShard:
    """Rows ``[begin, end)`` of the global batch dimension owned by ``rank``."""

    rank: int
    world_size: int
    batches_total: int
    begin: int
    end: int

    @property
    def rows(self) -> int:
        return self.end - self.begin


def shard_batches(batches_total: int, world_size: int, rank: int) -> BatchShard:
    """Contiguous, near-equal split in whole rows of ``network_size`` paths (FFT rows never split)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    if batches_total < world_size:
        raise ValueError(f"cannot shard {batches_total} batch rows over {world_size} ranks")
    base, extra = divmod(batches_total, world_size)
    begin = rank * base + min(rank, extra)
    return BatchShard(rank, world_size, batches_total, begin, begin + base + (1 if rank < extra else 0))


def shard_contracts(n_contracts: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, near-equal split of a contract batch: ``[begin, end)`` of ``rank`` (possibly empty)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_contracts, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


# below this many paths per contract and rank a batch-row shard no longer fills a GPU (c3 at the trainer-test size
# leaves each of 8 ranks 512 rows x 16 columns per contract) and whole contracts are dealt out instead
CONTRACT_SHARD_MAX_PATHS = 100_000


class DeviceOps(Protocol):
    def cf_fused(self, args: _cabi.FusedArgs) -> torch.Tensor: ...
    def fused_terminal(self, args: _cabi.FusedArgs) -> tuple[torch.Tensor, torch.Tensor]: ...
    def cf_from_terminal(self, args: _cabi.FusedArgs, terminal: torch.Tensor, tsum: torch.Tensor) -> torch.Tensor: ...


class _CabiOps:
    def __init__(self, device: torch.device, dtype: torch.dtype) -> None:
        self.device, self.dtype = device, dtype

    def cf_fused(self, args):
        return _cabi.cf_fused(args, self.device, self.dtype)

    def fused_terminal(self, args):
        return _cabi.fused_terminal(args, self.device, self.dtype)

    def cf_from_terminal(self, args, terminal, tsum):
        return _cabi.cf_from_terminal(args, terminal, tsum, self.dtype)


class PeerExchange:
    """Exchange buffers for the peer-memory all-reduce fused into the finalise kernel
    (``smc_cf_fused_p2p``): one cudaMalloc'ed buffer per rank, exported over cudaIpc and mapped by
    every other rank of ``group`` (all ranks on one node, NVLink / NVSwitch peers).

    ``torch.distributed`` is used once, to swap the 64-byte handles and to fence the set-up; the data
    path afterwards has no collective call.  Every rank must call ``sharded_cf_targets(..., exchange=)``
    the same number of times (the epoch counter advances in lock-step).
    """

    def __init__(self, capacity_contracts: int, network_size: int, *, group=None, timeout_ms: int = 0) -> None:
        self._own, self._peers, self.timeout_ms = None, [], int(timeout_ms)  # 0: SMC_P2P_TIMEOUT_MS or two minutes
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise ValueError("the peer exchange supports at most 16 ranks")
        self.capacity_contracts, self.network_size = capacity_contracts, network_size
        nbytes = int(_cabi.LIB.smc_p2p_buffer_bytes(capacity_contracts, network_size, self.world))
        # Set-up is collective: every rank takes part in both exchanges below even if its own allocation or
        # mapping failed, so that a failure anywhere raises on ALL ranks instead of leaving the others blocked.
        handle, problem = None, None
        try:
            self._own, handle = _cabi.p2p_alloc(nbytes)
        except Exception as exc:  # noqa: BLE001 - reported to every rank below
            problem = f"rank {self.rank}: {exc}"
        handles: list[bytes | None] = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        if problem is None and all(h is not None for h in handles):
            try:
                for q in range(self.world):  # one at a time, so that _release() sees what was opened if one fails
                    self._peers.append(self._own if q == self.rank else _cabi.p2p_open(handles[q]))
            except Exception as exc:  # noqa: BLE001
                problem = f"rank {self.rank}: {exc}"
        elif problem is None:
            problem = f"allocation failed on rank(s) {[q for q, h in enumerate(handles) if h is None]}"
        problems: list[str | None] = [None] * self.world
        dist.all_gather_object(problems, problem, group=group)
        if any(p is not None for p in problems):
            self._release()
            raise RuntimeError("peer exchange unavailable: " + "; ".join(p for p in problems if p is not None))
        self.epoch = 0
        dist.barrier(group=group)  # every buffer is zeroed and mapped before anyone writes

    def _release(self) -> None:
        for q, ptr in enumerate(self._peers):  # may be a prefix of the ranks if the set-up failed part-way
            if q != self.rank:
                _cabi.LIB.smc_p2p_close(ptr)
        self._peers = []
        if self._own is not None:
            _cabi.LIB.smc_p2p_free(self._own)
            self._own = None

    def _group_at(self, epoch: int) -> "_cabi.P2PGroup":
        g = _cabi.P2PGroup()
        g.rank, g.world = self.rank, self.world
        for q, ptr in enumerate(self._peers):
            g.buffers[q] = ptr
        g.capacity_contracts, g.network_size, g.epoch = self.capacity_contracts, self.network_size, epoch
        g.timeout_ms = self.timeout_ms
        return g

    def peek_group(self) -> "_cabi.P2PGroup":
        """The group of the NEXT exchange, without advancing the epoch: everything that can fail on the host
        (argument checks, workspace and output allocation) is done against this, and only then ``commit()``
        advances — a rank whose call raises before the launch stays in step with its peers."""
        return self._group_at(self.epoch + 1)

    def commit(self) -> None:
        self.epoch += 1

    def next_group(self) -> "_cabi.P2PGroup":
        self.commit()
        return self._group_at(self.epoch)

    def cf_fused(self, args: "_cabi.FusedArgs", device: torch.device, dtype: torch.dtype,
                 workspace: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        """``smc_cf_fused_p2p`` with the epoch advanced only once the call can no longer fail on the host."""
        group = self.peek_group()
        need = int(_cabi.LIB.smc_cf_fused_workspace_bytes(_cabi.byref(args)))
        _cabi.check(_cabi.LIB.smc_cf_fused_p2p_check(_cabi.byref(args), _cabi.byref(group), need))
        ws = workspace if workspace is not None and workspace.numel() >= need else torch.empty(max(need, 256), dtype=torch.uint8, device=device)
        if out is None:
            out = torch.empty((args.n_contracts, args.network_size), dtype=_cabi.complex_dtype(dtype), device=device)
        self.commit()
        return _cabi.cf_fused_p2p(args, group, device, dtype, ws, out=out)

    def timed_out_epoch(self) -> int:
        """Epoch of the first exchange on this rank that gave up waiting for a peer (0: none).  The targets of
        that call are NaN.  Synchronises the current stream."""
        status = ctypes.c_uint32(0)
        group = self._group_at(max(self.epoch, 1))
        _cabi.check(_cabi.LIB.smc_p2p_status(_cabi.byref(group), ctypes.byref(status), _cabi.stream_handle(None)))
        return int(status.value)

    def check(self) -> None:
        epoch = self.timed_out_epoch()
        if epoch:
            raise RuntimeError(f"peer exchange: rank {self.rank} timed out waiting for a peer in exchange {epoch} "
                               f"(of {self.epoch}); the targets of that call are NaN")

    def close(self) -> None:
        if self._peers:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)  # nobody is still writing into a buffer that is about to go away
            own, self._own = self._own, None
            self._release()                 # unmap the peers' buffers
            dist.barrier(group=self.group)  # ... before any owner frees its own
            _cabi.check(_cabi.LIB.smc_p2p_free(own))


def _contract_sharded(engine: BlackScholes, contracts: torch.Tensor, world: int, rank: int, group, ops: DeviceOps | None) -> torch.Tensor:
    """Whole contracts dealt out to the ranks; contract c still consumes matrix ``skip + c`` of the stream."""
    if ops is None:
        contracts = engine.contract_rows(contracts)
    else:
        if contracts.dim() != 2 or contracts.shape[1] != 6:
            raise ValueError(f"contracts must have shape [C, 6]; got {tuple(contracts.shape)}")
        contracts = contracts.to(torch.float64).contiguous()
    n, width = contracts.shape[0], engine._sp.network_size
    begin, end = shard_contracts(n, world, rank)
    per = -(-n // world)  # slices are padded to one length: all_gather_into_tensor wants equal parts
    cdtype = _cabi.complex_dtype(engine._dtype)
    device = contracts.device if ops is not None else engine._device
    mine = torch.zeros((per, width), dtype=cdtype, device=device)
    if end > begin:
        part = contracts[begin:end].contiguous()
        args = engine.fused_args(part, end - begin, matrix_offset=begin)
        if ops is None:
            out = _cabi.cf_fused(args, engine._device, engine._dtype)  # RAW or NORMALIZE: a contract's mean is local
        elif engine._cfg.normalization is ForwardNormalization.RAW:
            out = ops.cf_fused(args)
        else:
            terminal, tsum = ops.fused_terminal(args)
            out = ops.cf_from_terminal(args, terminal, tsum)
        mine[: end - begin] = out.to(cdtype)
    gathered = torch.empty((world * per, width), dtype=cdtype, device=device)  # rank q's slice at rows [q per, (q + 1) per)
    dist.all_gather_into_tensor(torch.view_as_real(gathered), torch.view_as_real(mine), group=group)
    engine.consume(n)
    pieces = [gathered[q * per : q * per + shard_contracts(n, world, q)[1] - shard_contracts(n, world, q)[0]] for q in range(world)]
    return torch.cat(pieces, dim=0)


def _all_reduce_sum(t: torch.Tensor, group) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t, op=dist.ReduceOp.SUM, group=group)


def sharded_cf_targets(
    engine: BlackScholes,
    contracts: torch.Tensor,
    *,
    group=None,
    ops: DeviceOps | None = None,
    max_staging_bytes: int = 8 << 30,
    exchange: PeerExchange | None = None,
    shard: str = "batches",
) -> torch.Tensor:
    """CF targets ``[C, N]`` of ``contracts`` (``[C, 6]`` float64 on the engine's device) computed by all ranks of
    ``group``; every rank returns the full result.

    ``shard="batches"`` (default): every rank simulates batch rows ``[begin, end)`` of ALL contracts and the partial
    sums are exchanged (one all-reduce, or the exchange fused into the kernels over peer memory).
    ``shard="contracts"``: the alternative of SURVEY.md §8e for small path counts — every rank simulates ALL batch rows
    of its contiguous slice of the contracts and one all-gather assembles ``[C, N]``; no reduction crosses ranks (so
    NORMALIZE needs no second exchange) and the result is bit-identical to the single-GPU one.
    ``shard="auto"``: contracts when a batch-row shard would leave a rank fewer than CONTRACT_SHARD_MAX_PATHS paths
    per contract and there are at least as many contracts as ranks, else batches."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if shard not in ("batches", "contracts", "auto"):
        raise ValueError(f"shard must be 'batches', 'contracts' or 'auto'; got {shard!r}")
    if shard == "auto":
        per_rank_paths = engine._sp.total_paths() // max(world, 1)
        shard = "contracts" if world > 1 and per_rank_paths < CONTRACT_SHARD_MAX_PATHS and contracts.shape[0] >= world else "batches"
    if shard == "contracts" and world > 1:
        return _contract_sharded(engine, contracts, world, rank, group, ops)
    if ops is None:
        contracts = engine.contract_rows(contracts)  # [C, 6] float64 contiguous on the engine's device, as cf_targets does
    else:  # a stand-in for the device calls (CPU tests) reads the rows where they are
        if contracts.dim() != 2 or contracts.shape[1] != 6:
            raise ValueError(f"contracts must have shape [C, 6]; got {tuple(contracts.shape)}")
        contracts = contracts.to(torch.float64).contiguous()
    sp = engine._sp
    shard = shard_batches(sp.batches_per_mc_run, world, rank)
    ops = ops or _CabiOps(engine._device, engine._dtype)
    n = contracts.shape[0]
    if engine._cfg.normalization is ForwardNormalization.RAW and exchange is not None and world > 1:
        # all-reduce fused into the finalise kernel over peer memory: no collective call on the data path
        args = engine.fused_args(contracts, n, batch_begin=shard.begin, batch_end=shard.end)
        out = exchange.cf_fused(args, engine._device, engine._dtype)
        engine.consume(n)
        return out
    if engine._cfg.normalization is ForwardNormalization.RAW:
        args = engine.fused_args(contracts, n, batch_begin=shard.begin, batch_end=shard.end)
        out = ops.cf_fused(args)
        _all_reduce_sum(out, group)
        engine.consume(n)
        return out
    # NORMALIZE: stage terminals for a chunk of contracts, allreduce their sums, then the payoff pass
    per_contract = shard.rows * sp.network_size * (4 if sp.dtype.value == "float32" else 8)
    chunk = max(1, min(n, max_staging_bytes // max(per_contract, 1)))
    outs = []
    for c0 in range(0, n, chunk):
        part = contracts[c0 : c0 + chunk].contiguous()
        args = engine.fused_args(part, part.shape[0], batch_begin=shard.begin, batch_end=shard.end)
        terminal, tsum = ops.fused_terminal(args)
        if exchange is not None and world > 1:  # both exchanges of the step through peer memory, one epoch
            pg = exchange.next_group()
            _cabi.p2p_allreduce_sum_f64(tsum, pg)
            out = _cabi.cf_from_terminal_p2p(args, pg, terminal, tsum, engine._dtype)
        else:
            _all_reduce_sum(tsum, group)
            out = ops.cf_from_terminal(args, terminal, tsum)
            _all_reduce_sum(out, group)
        engine.consume(part.shape[0])
        outs.append(out)
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
