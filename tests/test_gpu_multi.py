"""Two real GPUs over NCCL: the sharded fused path (batch rows split across ranks, one
all-reduce of the complex partial sums; NORMALIZE adds the terminal-sum all-reduce) reproduces the
single-GPU result.  Skipped unless the box exposes >= 2 devices (run with `gpurun --gpus 2`)."""

from __future__ import annotations

import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROWS = [(100.0, 100.0, 1.0, 0.05, 0.0, 0.2), (37.5, 41.0, 2.5, -0.01, 0.03, 0.65), (5.0, 4.0, 0.7, 0.1, 0.0, 1.1)]


def _worker(rank: int, world: int, port: int, prec: str, norm: str, queue) -> None:
    import torch.distributed as dist

    from spectralmc_b200.distributed import sharded_cf_targets
    from spectralmc_b200.effects import ForwardNormalization, PathScheme
    from spectralmc_b200.gbm import BlackScholes
    from spectralmc_b200.numerical import Precision
    from tests.helpers import expect_success, make_black_scholes_config, make_simulation_params

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        sp = make_simulation_params(timesteps=20, network_size=64, batches_per_mc_run=1001, mc_seed=5, skip=2, dtype=Precision(prec))
        cfg = make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization(norm))
        contracts = torch.tensor(ROWS, dtype=torch.float64, device="cuda")
        sharded = sharded_cf_targets(BlackScholes(cfg), contracts)
        whole = expect_success(BlackScholes(cfg).cf_targets(contracts))  # unsharded, on this rank alone
        torch.cuda.synchronize()
        queue.put((rank, sharded.cpu().numpy(), whole.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("prec", ["float32", "float64"])
@pytest.mark.parametrize("norm", ["raw_paths", "normalize_forwards"])
def test_two_gpu_sharding_matches_single_gpu(prec, norm) -> None:
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, prec, norm, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tol = 2e-6 if prec == "float32" else 1e-13
    for rank, sharded, whole in results:
        assert np.max(np.abs(sharded - whole)) <= tol * np.max(np.abs(whole)), rank
    assert np.array_equal(results[0][1], results[1][1])


def _train_worker(rank: int, world: int, port: int, queue) -> None:
    import torch.distributed as dist

    from spectralmc_b200.cvnn import make_cvnn
    from spectralmc_b200.effects import ForwardNormalization, PathScheme
    from spectralmc_b200.gbm_trainer import GbmCVNNPricer, build_gbm_cvnn_pricer_config, build_training_config
    from spectralmc_b200.numerical import Precision
    from tests.helpers import expect_success, make_black_scholes_config, make_domain_bounds, make_simulation_params

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        sp = make_simulation_params(timesteps=4, network_size=16, batches_per_mc_run=1024, mc_seed=9, dtype=Precision.float32)
        cfg = make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
        net = make_cvnn(6, 16, seed=9, device=f"cuda:{rank}")
        pc = expect_success(build_gbm_cvnn_pricer_config(cfg=cfg, domain_bounds=make_domain_bounds(), cvnn=net))
        trainer = expect_success(GbmCVNNPricer.create(pc, process_group=dist.group.WORLD, peer_exchange=True))
        result = expect_success(trainer.train(expect_success(build_training_config(num_batches=3, batch_size=16, learning_rate=1e-2))))
        single = expect_success(GbmCVNNPricer.create(expect_success(build_gbm_cvnn_pricer_config(
            cfg=cfg, domain_bounds=make_domain_bounds(), cvnn=make_cvnn(6, 16, seed=9, device=f"cuda:{rank}")))))
        alone = expect_success(single.train(expect_success(build_training_config(num_batches=3, batch_size=16, learning_rate=1e-2))))
        torch.cuda.synchronize()
        queue.put((rank, result.losses, alone.losses, [p.detach().cpu().numpy() for p in net.parameters()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_training_replicas_stay_identical() -> None:
    """Batch rows sharded over two ranks, partial sums exchanged inside the finalise kernel over peer
    memory, each rank steps its own CVNN replica through the fused step: the replicas stay
    bit-identical and track the single-GPU run."""
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((queue.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, a0, p0), (_, l1, a1, p1) = results
    assert l0 == l1 and all(np.array_equal(x, y) for x, y in zip(p0, p1))
    assert max(abs(x - y) / abs(y) for x, y in zip(l0, a0)) <= 1e-4  # sharded sums differ from single-GPU sums in the last bits only


def _p2p_worker(rank: int, world: int, port: int, prec: str, queue, norm: str = "raw_paths") -> None:
    import torch.distributed as dist

    from spectralmc_b200.distributed import PeerExchange, sharded_cf_targets
    from spectralmc_b200.effects import ForwardNormalization, PathScheme
    from spectralmc_b200.gbm import BlackScholes
    from spectralmc_b200.numerical import Precision
    from tests.helpers import expect_success, make_black_scholes_config, make_simulation_params

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        out = {}
        for N, B, C in ((64, 1001, 3), (48, 37, 700), (16, 8, 3000)):  # 3000 contracts: several per CTA of the persistent exchange grid
            sp = make_simulation_params(timesteps=20, network_size=N, batches_per_mc_run=B, mc_seed=5, skip=2, dtype=Precision(prec))
            cfg = make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization(norm))
            rows = np.tile(np.asarray(ROWS), (C // 3 + 1, 1))[:C]
            rows[:, 1] *= np.linspace(0.8, 1.2, C)
            contracts = torch.tensor(rows, dtype=torch.float64, device="cuda")
            exchange = PeerExchange(C, N)
            engine = BlackScholes(cfg)
            fused = [sharded_cf_targets(engine, contracts, exchange=exchange) for _ in range(3)]  # three epochs: both slots reused
            nccl_engine = BlackScholes(cfg)
            nccl = [sharded_cf_targets(nccl_engine, contracts) for _ in range(3)]
            whole = expect_success(BlackScholes(cfg).cf_targets(contracts))
            torch.cuda.synchronize()
            assert expect_success(engine.snapshot()).sim_params.skip == 2 + 3 * C
            out[(N, C)] = ([t.cpu().numpy() for t in fused], [t.cpu().numpy() for t in nccl], whole.cpu().numpy())
            exchange.close()
        queue.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("norm", ["raw_paths", "normalize_forwards"])
@pytest.mark.parametrize("prec", ["float32", "float64"])
def test_two_gpu_fused_peer_exchange_matches_nccl_and_single_gpu(prec, norm) -> None:
    """The all-reduce fused into the finalise kernel over peer memory (smc_cf_fused_p2p; for NORMALIZE
    smc_p2p_allreduce_sum_f64 + smc_cf_from_terminal_p2p): complete targets on every rank, bit-identical
    across ranks, equal to the NCCL route and to one GPU."""
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, prec, queue, norm)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((queue.get(timeout=240) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tol = 2e-6 if prec == "float32" else 1e-13
    for key in results[0][1]:
        (f0, n0, w0), (f1, n1, w1) = results[0][1][key], results[1][1][key]
        for step in range(3):
            assert np.array_equal(f0[step], f1[step]), (key, step)  # every rank forms the identical sum
            scale = np.max(np.abs(n0[step]))
            assert np.max(np.abs(f0[step] - n0[step])) <= tol * scale, (key, step)
        assert np.max(np.abs(f0[0] - w0)) <= tol * np.max(np.abs(w0)), key
        assert not np.array_equal(f0[0], f0[1])  # later calls consume later normal matrices


def _contract_shard_worker(rank: int, world: int, port: int, norm: str, queue) -> None:
    import torch.distributed as dist

    from spectralmc_b200.distributed import sharded_cf_targets
    from spectralmc_b200.effects import ForwardNormalization, PathScheme
    from spectralmc_b200.gbm import BlackScholes
    from spectralmc_b200.numerical import Precision
    from tests.helpers import expect_success, make_black_scholes_config, make_simulation_params

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        # the reference's own test size: one timestep, 16 x 4096 paths per contract (tests/test_gbm_trainer.py:127-136)
        sp = make_simulation_params(timesteps=1, network_size=16, batches_per_mc_run=4096, mc_seed=5, skip=7, dtype=Precision.float32)
        cfg = make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization(norm))
        rows = np.tile(np.asarray(ROWS[0]), (37, 1))
        rows[:, 1] = np.linspace(70.0, 130.0, 37)  # 37 contracts: ragged slices (19 + 18)
        contracts = torch.tensor(rows, dtype=torch.float64, device="cuda")
        engine = BlackScholes(cfg)
        dealt = sharded_cf_targets(engine, contracts, shard="contracts")
        auto = sharded_cf_targets(BlackScholes(cfg), contracts, shard="auto")  # picks contracts at this size
        whole = expect_success(BlackScholes(cfg).cf_targets(contracts))
        torch.cuda.synchronize()
        queue.put((rank, dealt.cpu().numpy(), auto.cpu().numpy(), whole.cpu().numpy(), expect_success(engine.snapshot()).sim_params.skip))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("norm", ["raw_paths", "normalize_forwards"])
def test_two_gpu_contract_sharding_is_bit_identical_to_one_gpu(norm) -> None:
    """Whole contracts dealt out + one all-gather (the alternative for small path counts): no sum
    crosses ranks, so every rank holds exactly the single-GPU bits, and the stream advances by all 37 matrices."""
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_contract_shard_worker, args=(r, 2, port, norm, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, dealt, auto, whole, skip in results:
        assert np.array_equal(dealt, whole) and np.array_equal(auto, whole), rank
        assert skip == 7 + 37
