"""The N > 1 path on CPU: world_size-2 (and 3) ``gloo`` runs of the batch-sharding + all-reduce
logic in spectralmc_b200.distributed, with the device calls replaced by an oracle stand-in that
honours the same ``smc_fused_args`` (batch range, global path counters, partial scaling).
What is verified is the host logic: shard ranges, which collectives run on what, NORMALIZE's
two-phase protocol, and the engine's skip bookkeeping — the sum over ranks must equal the
unsharded oracle result."""

from __future__ import annotations

import ctypes
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gbm as ogbm
from oracle import philox
from spectralmc_b200 import _cabi
from spectralmc_b200.distributed import shard_batches, sharded_cf_targets
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import BlackScholes
from spectralmc_b200.numerical import Precision
from tests.helpers import make_black_scholes_config, make_simulation_params

ROWS = [(100.0, 100.0, 1.0, 0.05, 0.0, 0.2), (37.5, 41.0, 2.5, -0.01, 0.03, 0.65), (5.0, 4.0, 0.7, 0.1, 0.0, 1.1)]
T, N, B, SEED, SKIP = 6, 8, 21, 11, 4


def test_shard_batches_partitions_the_rows() -> None:
    for total in (8, 21, 65536, 7):
        for world in (1, 2, 3, 7):
            if total < world:
                continue
            shards = [shard_batches(total, world, r) for r in range(world)]
            assert shards[0].begin == 0 and shards[-1].end == total
            assert all(a.end == b.begin for a, b in zip(shards[:-1], shards[1:]))
            assert max(s.rows for s in shards) - min(s.rows for s in shards) <= 1
    with pytest.raises(ValueError):
        shard_batches(2, 3, 0)
    with pytest.raises(ValueError):
        shard_batches(8, 2, 2)


class OracleOps:
    """CPU stand-in for the three device calls, driven by the SAME smc_fused_args."""

    def _contracts(self, a: _cabi.FusedArgs) -> np.ndarray:
        buf = (ctypes.c_double * (6 * a.n_contracts)).from_address(a.contracts)
        return np.ctypeslib.as_array(buf).reshape(a.n_contracts, 6).copy()

    def _terminal(self, a, row, k):
        lo, hi = a.batch_begin * a.network_size, a.batch_end * a.network_size
        z = philox.normals_matrix(a.timesteps, a.batches_total * a.network_size, np.float64, a.seed, k, col_begin=lo, col_end=hi)
        X0, K, Tm, r, d, v = row
        ogbm.simulate_paths_inplace(z, a.timesteps, Tm / a.timesteps, X0, r, d, v, a.scheme == _cabi.SMC_LOG_EULER)
        return z[-1]

    def _cf(self, a, row, terminal, scale):
        X0, K, Tm, r, d, v = row
        put = np.exp(-r * Tm) * np.maximum(K - terminal * scale, 0.0)
        mat = put.reshape(a.batch_end - a.batch_begin, a.network_size)
        return np.fft.fft(mat, axis=1).sum(axis=0) / a.batches_total  # PARTIAL mean

    def cf_fused(self, a):
        rows = self._contracts(a)
        return torch.from_numpy(np.stack([self._cf(a, r, self._terminal(a, r, a.first_matrix_index + i), 1.0) for i, r in enumerate(rows)]))

    def fused_terminal(self, a):
        rows = self._contracts(a)
        term = np.stack([self._terminal(a, r, a.first_matrix_index + i) for i, r in enumerate(rows)])
        return torch.from_numpy(term), torch.from_numpy(term.sum(axis=1))

    def cf_from_terminal(self, a, terminal, tsum):
        rows = self._contracts(a)
        out = []
        for i, r in enumerate(rows):
            X0, K, Tm, rr, d, v = r
            mean = float(tsum[i]) / (a.batches_total * a.network_size)
            out.append(self._cf(a, r, terminal[i].numpy(), X0 * np.exp((rr - d) * Tm) / mean))
        return torch.from_numpy(np.stack(out))


def _engine(norm: ForwardNormalization) -> BlackScholes:
    sp = make_simulation_params(timesteps=T, network_size=N, batches_per_mc_run=B, mc_seed=SEED, skip=SKIP, dtype=Precision.float64)
    return BlackScholes(make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=norm))


def _worker(rank: int, world: int, port: int, norm_value: str, chunk_bytes: int, queue, shard: str = "batches") -> None:
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        engine = _engine(ForwardNormalization(norm_value))
        contracts = torch.tensor(ROWS, dtype=torch.float64)
        out = sharded_cf_targets(engine, contracts, ops=OracleOps(), max_staging_bytes=chunk_bytes, shard=shard)
        queue.put((rank, out.numpy(), engine.snapshot().unwrap().sim_params.skip))
    finally:
        dist.destroy_process_group()


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize(
    "world,norm,chunk_bytes,shard",
    [
        (2, ForwardNormalization.RAW, 8 << 30, "batches"),
        (2, ForwardNormalization.NORMALIZE, 8 << 30, "batches"),
        (2, ForwardNormalization.NORMALIZE, 1, "batches"),  # 1 byte: one contract per NORMALIZE chunk
        (3, ForwardNormalization.NORMALIZE, 8 << 30, "batches"),
        # whole contracts dealt out + one all-gather (the alternative for small path counts); with more
        # ranks than contracts divide evenly, slices are ragged and padded; "auto" picks contracts at this size
        (2, ForwardNormalization.RAW, 8 << 30, "contracts"),
        (2, ForwardNormalization.NORMALIZE, 8 << 30, "contracts"),
        (3, ForwardNormalization.NORMALIZE, 8 << 30, "auto"),
    ],
)
def test_sharded_targets_equal_the_unsharded_result(world, norm, chunk_bytes, shard) -> None:
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, norm.value, chunk_bytes, queue, shard)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the unsharded reference: the oracle's own end-to-end function on the full matrices
    ref = []
    for i, row in enumerate(ROWS):
        z = philox.normals_matrix(T, N * B, np.float64, SEED, SKIP + i)
        cf, _ = ogbm.simulate_fft(ogbm.Contract(*row), z, N, normalization=norm.value)
        ref.append(cf)
    ref = np.stack(ref)
    for rank, out, skip in results:
        assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref)), rank
        assert skip == SKIP + len(ROWS)  # every rank advances the stream by one matrix per contract
    assert np.array_equal(results[0][1], results[1][1])  # all ranks hold the same all-reduced tensor


def _exchange_worker(rank: int, world: int, port: int, queue) -> None:
    from spectralmc_b200.distributed import PeerExchange

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        try:
            PeerExchange(4, 16)
            queue.put((rank, "created"))
        except RuntimeError as exc:
            queue.put((rank, str(exc)))
        dist.barrier()  # still in step with the other ranks after the failure
    finally:
        dist.destroy_process_group()


def test_peer_exchange_setup_fails_on_every_rank_together() -> None:
    """Without a device every rank's exchange-buffer allocation fails; the set-up is collective, so all ranks
    learn about it and raise (nobody is left blocked in a handle exchange), and the process group stays usable."""
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        assert results[rank].startswith("peer exchange unavailable") and "rank 0" in results[rank] and "rank 1" in results[rank]
