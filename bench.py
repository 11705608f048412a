#!/usr/bin/env python
"""Headline benchmark: GBM path-steps/s (and CF estimates/s) of the batch-generation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic contracts: Philox normals ->
GBM stepping -> payoff -> FFT/mean, i.e. one CF training target per contract
(reference: gbm_trainer.py:1546-1553 for one `_run_batch`).

Workload (BASELINE.json configs[1], "c2"): fp32, 1 contract (X0=100,K=100,T=1,r=.05,d=0,v=.2),
252 timesteps, network_size=128, batches_per_mc_run=65 536 per GPU, RAW, LOG_EULER.
Multi-GPU `value`: weak scaling — every rank simulates 65 536 batch rows of the same contract batch
(global path counters, B_total = 65 536 * N), then ONE exchange of the partial CF sums.

The JSON line carries
  value / ms_per_step   exactly K steps, device-resident inputs, CUDA events, max over ranks
  e2e                   the same K steps through the C-ABI host entry point with HOST buffers (pinned H2D of the
                        contracts, D2H of the targets, inside the timed region)
  sustained             the same step back to back for >= 2 s (what the clock sampler can actually see)
  clocks                nvidia-smi samples taken DURING the timed regions and the sustained leg
  roofline              the dominant kernel (fused simulator: XU / FP32-issue pipes, calibrated live)
  roofline_materialised the HBM-bound generator / steppers of the materialised-normals mode vs MEASURED_PEAKS.json
  workloads             N = 1: c2 NORMALIZE, c2 SIMPLE_EULER, c4 (fp64, FP64-pipe roofline from a live DFMA
                        calibration), c3 (the trainer-test training step through GbmCVNNPricer.train)
  gpu_reference_baseline  N = 1: the reference's own Numba kernel body (taken from its source by AST) + its
                        per-contract pipeline shape on this GPU, RAW and NORMALIZE — the ">= 10x" comparison
  cpu_baseline          N = 1: the oracle port on the host cores, bounded sample
  parity_check          N > 1: fused peer exchange == NCCL route == unsharded result, bit-identical across ranks;
                        a failure exits non-zero
  strong                N > 1: FIXED problems (c2 itself; 64 contracts x 2^20 rows) split over the N ranks, with the
                        one-GPU time of the same problem measured in the same run
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CANON = (100.0, 100.0, 1.0, 0.05, 0.0, 0.2)  # reference tests/test_gbm.py:146
ODD = (37.5, 41.0, 2.5, -0.01, 0.03, 0.65)
WORKLOADS = {
    # name: (contracts, timesteps, network_size, batches per GPU, dtype)
    "c2": dict(contracts=1, T=252, N=128, B=65536, dtype="float32"),
    "c2x8": dict(contracts=8, T=252, N=128, B=65536, dtype="float32"),
    "c3_trainer": dict(contracts=1024, T=1, N=16, B=4096, dtype="float32"),
    "c4": dict(contracts=512, T=365, N=256, B=4096, dtype="float64"),
    # BASELINE configs[4] on 8 GPUs: 4096 contracts x 2^20 batch rows (131 072 per GPU), 1.39e14 path-steps per step
    "c5": dict(contracts=4096, T=252, N=128, B=131072, dtype="float32"),
}
# SASS-counted work of the fused kernels' inner loops; tests/test_bench_contract.py re-counts them from the built
# library (cuobjdump) so that a kernel edit cannot silently invalidate the rooflines.
F32_LOOP = {"kernel": "_ZN3smc11step_kernelIfLi0ELi0ELi0ELi0E", "instructions": 137, "mufu": 18, "normals": 12}
F64_LOOP = {"kernel": "_ZN3smc11step_kernelIdLi0ELi0ELi0ELi1E", "instructions": 150, "fp64": 76, "normals": 4}
ISSUE_SLOTS_PER_STEP = F32_LOOP["instructions"] / F32_LOOP["normals"]  # warp-instructions per path-step per lane
XU_OPS_PER_STEP = F32_LOOP["mufu"] / F32_LOOP["normals"]              # log-Euler sum loop: LG2 + SQRT + ONE SIN per Box-Muller pair = 1.5 per normal
XU_OPS_PER_STEP_ALL_NORMALS = 2.0                                     # loops that need every normal (SIMPLE_EULER): LG2 + SQRT + SIN + COS per pair
FP64_OPS_PER_STEP = F64_LOOP["fp64"] / F64_LOOP["normals"]            # DFMA + DMUL + DADD per float64 path-step


def measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self) -> dict:
        sm, mx, power, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        under_load = sum(1 for w in power if w > 0.5 * max(power))
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "samples_under_load": under_load, "power_w_max": max(power),
                "window": "the K timed steps (device and e2e) and the sustained leg"}


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------
def cpu_pass(w: dict, sample_B: int, threads: int, matrix_index: int) -> float:
    """One pass of the oracle's C restatement (oracle/gbm_oracle.c via oracle/cport.py) over
    `sample_B` batch rows of the workload: Philox normals -> path stepping (float64 arithmetic, as
    the reference kernel) -> payoff -> CF, on `threads` host threads.  Returns seconds."""
    import numpy as np

    from oracle import cport

    t0 = time.perf_counter()
    cport.simulate_fft(CANON, w["T"], w["N"], sample_B, np.dtype(w["dtype"]), True, False, 7, matrix_index, threads=threads)
    return time.perf_counter() - t0


def cpu_sample_rows(w: dict, args, cores: int) -> int:
    # ~0.25-0.5 s of CPU work per pass at ~1e7 path-steps/s/core
    return args.cpu_sample_batches or max(64, min(w["B"], int(cores * 4e6 / (w["T"] * w["N"]))))


def run_reference(args) -> None:
    """--impl reference: the reference's algorithm on the host cores.  The reference itself is
    GPU-only Python (Numba + CuPy, import-time CUDA asserts) and cannot run on a CPU, so this arm
    times the oracle port with every host thread on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    sample_B = cpu_sample_rows(w, args, cores)
    for i in range(max(args.warmup, 1)):
        cpu_pass(w, sample_B, cores, i)
    times = [cpu_pass(w, sample_B, cores, 100 + i) for i in range(args.steps)]
    total = sum(times)
    value = args.steps * sample_B * w["N"] * w["T"] / total
    sample = (f"each step = {sample_B} of {w['B']} batch rows ({sample_B * w['N']} paths x {w['T']} steps): normals + stepping + payoff + CF; "
              f"oracle/gbm_oracle.c on {cores} threads")
    line = {
        "impl": "reference", "metric": "gbm_path_steps_per_sec", "value": value, "unit": "path-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if w["dtype"] == "float32" else "f64",
        "data": "synthetic", "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "path-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cf_estimates_per_sec": value / (w["B"] * w["N"] * w["T"]),
    }
    print(json.dumps(line), flush=True)


def workload_config(name: str, gpus: int) -> dict:
    """The same dictionary for both arms (the exchange route a multi-GPU run took is reported beside it, as
    `collective`, because only the b200 arm knows it at run time)."""
    w = WORKLOADS[name]
    return {
        "workload": f"{name}: {w['contracts']} contract(s) {w['dtype']}, T={w['T']}, N={w['N']}, B={w['B']} per GPU, RAW, LOG_EULER, fused Philox normals",
        "contracts_per_step": w["contracts"], "timesteps": w["T"], "network_size": w["N"],
        "batches_per_gpu": w["B"], "batches_total": w["B"] * gpus, "parallelism": f"batch-shard x{gpus}, one exchange of the partial sums per step" if gpus > 1 else "single GPU",
        "l2": "fused mode has no HBM-resident inputs (normals are drawn in registers); a 256 MiB L2 flush runs between timed steps anyway; materialised-mode inputs (8.5 GB) exceed L2",
    }


# ------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    from spectralmc_b200 import _cabi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = WORKLOADS[args.workload]
    dtype = torch.float32 if w["dtype"] == "float32" else torch.float64
    C, T, N, B = w["contracts"], w["T"], w["N"], w["B"]
    B_total = B * world
    rows = np.tile(np.asarray(CANON, dtype=np.float64), (C, 1))
    if C > 1:  # vary the strike so contracts differ
        rows[:, 1] = np.linspace(80.0, 120.0, C)
    contracts_dev = torch.tensor(rows, device=dev)
    contracts_pin = torch.tensor(rows).pin_memory()
    out_pin = torch.empty((C, N), dtype=_cabi.complex_dtype(dtype)).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def fused_args(contracts, n_contracts, t, n, b_total, lo, hi, dt, first=0, scheme=_cabi.SMC_LOG_EULER, norm=_cabi.SMC_RAW):
        return _cabi.make_fused_args(contracts, n_contracts, t, n, b_total, dt, scheme, norm, 7, first, batch_begin=lo, batch_end=hi)

    def make_args(step: int, contracts):
        return fused_args(contracts, C, T, N, B_total, rank * B, (rank + 1) * B, dtype, first=step * C)

    ws = torch.empty(_cabi.cf_fused_host_workspace_bytes(make_args(0, None)) + 4096, dtype=torch.uint8, device=dev)
    # multi-GPU exchange of the partial sums: fused into the step kernel over peer memory
    # (smc_cf_fused_p2p) when the IPC set-up succeeds, else one ncclAllReduce after the kernel
    exchange, collective = None, "none"
    if world > 1:
        collective = "nccl allreduce"
        if args.collective in ("auto", "p2p"):
            try:
                from spectralmc_b200.distributed import PeerExchange

                exchange = PeerExchange(max(C, 64), N)  # 64: room for the strong-scaling problem's contracts
                collective = "fused peer-memory exchange (smc_cf_fused_p2p)"
            except Exception as exc:  # noqa: BLE001 - any set-up failure falls back to the NCCL route, visibly
                if args.collective == "p2p":
                    raise
                collective = f"nccl allreduce (peer exchange unavailable: {type(exc).__name__}: {exc})"
        ok = torch.tensor([1 if exchange is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # every rank must take the same route
        if int(ok.item()) == 0 and exchange is not None:
            exchange, collective = None, "nccl allreduce (peer exchange unavailable on another rank)"

    def run_sharded(a, work, out=None, route="auto"):
        """One sharded step of arguments `a`: complete targets on every rank."""
        if exchange is not None and route in ("auto", "p2p"):
            return exchange.cf_fused(a, dev, dtype, work, out=out)
        res = _cabi.cf_fused(a, dev, dtype, work)
        if world > 1:
            dist.all_reduce(torch.view_as_real(res))
        return res

    launches = {"n": 0}
    plan_launches = int(_cabi.LIB.smc_cf_fused_launch_count(_cabi.byref(make_args(0, None)))) + (1 if exchange is not None else 0)
    zero_copy_out = exchange is not None and out_pin.numel() * out_pin.element_size() <= 65536

    def step_device(step: int):
        out = run_sharded(make_args(step, contracts_dev), ws)
        launches["n"] += plan_launches
        return out

    def step_e2e(step: int):
        if world == 1:
            _cabi.cf_fused_host(make_args(step, None), contracts_pin, out_pin, ws)  # H2D + kernel + D2H + sync
        elif zero_copy_out:
            # as smc_cf_fused_host does on one GPU: contracts copied H2D, targets written by the collecting CTAs
            # straight into the small pinned buffer through its device alias (no D2H copy)
            cdev = contracts_pin.to(dev, non_blocking=True)
            run_sharded(make_args(step, cdev), ws, out=out_pin)
            torch.cuda.current_stream().synchronize()
        else:
            cdev = contracts_pin.to(dev, non_blocking=True)
            out = run_sharded(make_args(step, cdev), ws)
            out_pin.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        launches["n"] += plan_launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed(fn, steps: int, first_step: int, flush_l2: bool = True) -> float:
        """Sum of per-step CUDA-event durations (ms); L2 flushed between steps, outside the events."""
        pairs = []
        barrier()
        for s in range(steps):
            if flush_l2:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(first_step + s)
            b.record()
            pairs.append((a, b))
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in pairs))

    for s in range(args.warmup):
        step_device(s)
        step_e2e(s)
    barrier()
    launches["n"] = 0  # count only what the timed regions launch

    with ClockSampler(local) as clocks:
        ms_dev = timed(step_device, args.steps, 1000)
        ms_e2e = timed(step_e2e, args.steps, 2000)
        counted = launches["n"]
        # sustained leg: the same step back to back for >= args.sustain_s seconds, one event pair around the whole
        # run (what a 50 ms clock sampler can see; the K-step regions above last a few tens of milliseconds)
        sustain_steps = max(args.steps, int(args.sustain_s * 1e3 / max(ms_dev / args.steps, 1e-3)))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(sustain_steps):
            step_device(5000 + s)
        e1.record()
        barrier()
        ms_sustain = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.summary()

    path_steps_per_step = float(C) * B_total * N * T
    value = path_steps_per_step * args.steps / (ms_dev * 1e-3)
    e2e_value = path_steps_per_step * args.steps / (ms_e2e * 1e-3)

    line = {
        "metric": "gbm_path_steps_per_sec", "value": value, "unit": "path-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dtype == torch.float32 else "f64", "data": "synthetic",
        "config": workload_config(args.workload, world), "collective": collective,
        "cf_estimates_per_sec": C * args.steps / (ms_dev * 1e-3),
        "e2e": {"value": e2e_value, "unit": "path-steps/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(contracts_pin.numel() * 8), "d2h_bytes_per_step": int(out_pin.numel() * out_pin.element_size()),
                "api": "smc_cf_fused_host (C ABI, pinned host buffers)" if world == 1 else
                       (f"H2D of the contracts + smc_cf_fused_p2p writing the targets into the pinned host buffer through its device alias [{collective}]"
                        if zero_copy_out else f"H2D + sharded fused path [{collective}] + D2H")},
        "sustained": {"value": path_steps_per_step * sustain_steps / (ms_sustain * 1e-3), "unit": "path-steps/s",
                      "ms_per_step": ms_sustain / sustain_steps, "steps": sustain_steps, "seconds": ms_sustain * 1e-3,
                      "note": "back-to-back steps, no L2 flush in between (the fused path has no HBM-resident input)"},
        "gpu_launches": counted, "kernels_per_step": plan_launches, "clocks": clk,
    }

    failed = False
    if world > 1:
        line["parity_check"] = parity_check(_cabi, torch, dist, dev, exchange, world, rank, fused_args, run_sharded, dtype)
        failed = not line["parity_check"]["ok"]
        line["strong"] = strong_scaling(_cabi, torch, dev, world, rank, fused_args, run_sharded, barrier, max_over_ranks, args)

    if rank == 0:
        # ---- rooflines (rank 0, after the timed region) ----
        peaks = measured_peaks()
        calib = {}
        for kind, name in ((0, "ffma"), (1, "mufu"), (2, "philox"), (3, "dfma")):
            ops, ms = _cabi.pipe_calibrate(kind, 1 << 15, dev)
            calib[name] = ops / (ms * 1e-3)
        line["roofline"] = fused_f32_roofline(value / world, calib)
        line["roofline"]["kernel_share_of_step"] = "the step is this one kernel (+ a memset of its control words)" if world == 1 else \
            "step = this kernel + the exchange collect kernel"
        line["roofline_materialised"] = materialised_roofline(_cabi, torch, dev, peaks, T, N, dtype)
        if world == 1 and args.workload == "c2" and not args.no_workloads:
            line["workloads"] = secondary_workloads(_cabi, torch, dev, calib, args)
            line["gpu_reference_baseline"] = numba_reference_leg(torch, value, args)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(w, args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if exchange is not None:
            exchange.close()
        dist.destroy_process_group()
    if failed:
        sys.exit(3)


def fused_f32_roofline(per_gpu_rate: float, calib: dict, xu_ops: float = XU_OPS_PER_STEP, issue_slots: float = ISSUE_SLOTS_PER_STEP) -> dict:
    """Roofline of smc::step_kernel<float, FUSED, LOG_EULER, COLSUM>: bound by the XU (MUFU) and FP32-issue
    pipes, not by HBM or tensor cores.  Algorithmic work per path-step from the SASS of the inner loop
    (F32_LOOP), peaks from the live microbenchmarks (MEASURED_PEAKS.json holds only HBM and bf16)."""
    issue_ach, xu_ach = per_gpu_rate * issue_slots, per_gpu_rate * xu_ops
    fr_issue, fr_xu = issue_ach / calib["ffma"], xu_ach / calib["mufu"]
    bound = "fp32_issue" if fr_issue >= fr_xu else "xu"
    return {
        "kernel": "smc::step_kernel<float, SRC_FUSED, LOG_EULER, OUT_COLSUM>",
        "bound": bound,
        "achieved": (issue_ach if bound == "fp32_issue" else xu_ach) / 1e12,
        "peak": (calib["ffma"] if bound == "fp32_issue" else calib["mufu"]) / 1e12,
        "unit": "Tlane-op/s", "frac": max(fr_issue, fr_xu),
        # not measured in this run; the figure of the committed ncu capture is in profiles/r2_ncu_summary.txt
        # (a few tens of KB per launch against 48 B of algorithmic input: the kernel has no HBM-resident operand)
        "traffic": None,
        "peak_source": "calibrated live: smc_pipe_calibrate (FFMA issue-rate and MUFU.EX2 microbenchmarks); MEASURED_PEAKS.json holds only HBM/bf16",
        "detail": {"issue_slots_per_path_step": issue_slots, "xu_ops_per_path_step": xu_ops,
                   "fp32_issue": {"achieved": issue_ach / 1e12, "peak": calib["ffma"] / 1e12, "frac": fr_issue},
                   "xu": {"achieved": xu_ach / 1e12, "peak": calib["mufu"] / 1e12, "frac": fr_xu},
                   "philox_blocks_per_s_peak": calib["philox"]},
    }


def parity_check(_cabi, torch, dist, dev, exchange, world, rank, fused_args, run_sharded, dtype) -> dict:
    """Correctness of the multi-GPU path in the driver's own record: on a small shape (3 contracts, radix-2 N)
    and on the full c2 shape, the fused peer exchange (cf_exchange push in finish_contract + exchange_collect)
    must equal the NCCL route, both must equal the UNSHARDED single-GPU result, and every rank must hold
    identical bits."""
    import numpy as np

    res = {"ok": True, "tolerance": 2e-6, "cases": []}

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())

    for name, (rows, t, n, b_total) in {"small": ([CANON, ODD, (90.0, 100.0, 2.0, 0.03, 0.01, 0.3)], 24, 128, 96 * world),
                                        "c2": ([CANON], 252, 128, 65536)}.items():
        contracts = torch.tensor(np.asarray(rows, dtype=np.float64), device=dev)
        per = b_total // world
        lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else b_total
        shard = fused_args(contracts, len(rows), t, n, b_total, lo, hi, dtype, first=77)
        full = fused_args(contracts, len(rows), t, n, b_total, 0, b_total, dtype, first=77)
        unsharded = _cabi.cf_fused(full, dev, dtype)
        nccl = run_sharded(shard, None, route="nccl")
        case = {"case": name, "nccl_vs_unsharded": rel(nccl, unsharded)}
        outs = [nccl]
        if exchange is not None:
            p2p = run_sharded(shard, None, route="p2p")
            case["p2p_vs_nccl"] = rel(p2p, nccl)
            case["p2p_vs_unsharded"] = rel(p2p, unsharded)
            outs.append(p2p)
            exchange.check()
        identical = True
        for o in outs:  # bit-identical on every rank
            ref = o.clone()
            dist.broadcast(ref, src=0)
            same = torch.tensor([1 if torch.equal(torch.view_as_real(ref), torch.view_as_real(o)) else 0], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            identical = identical and bool(same.item())
        case["bit_identical_across_ranks"] = identical
        ok = identical and all(v <= res["tolerance"] for k, v in case.items() if k.endswith(("_vs_nccl", "_vs_unsharded")))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        case["ok"] = bool(flag.item())
        res["ok"] = res["ok"] and case["ok"]
        res["cases"].append(case)
    return res


def strong_scaling(_cabi, torch, dev, world, rank, fused_args, run_sharded, barrier, max_over_ranks, args) -> dict:
    """FIXED problems split over the ranks (batch rows sharded, one exchange per step), beside the one-GPU time of
    the same problem measured in the same run (every rank runs the whole problem alone, at the same time; max
    over ranks).  c2 is BASELINE configs[1] itself; c5cut is configs[4]'s shape with the contracts cut to 64
    (4096 contracts x 2^20 rows is 84 s per step on one GPU)."""
    import numpy as np

    dtype = torch.float32
    out = {}
    for name, (c, t, n, b_total, steps) in {"c2": (1, 252, 128, 65536, max(args.steps, 10)), "c5cut": (64, 252, 128, 1 << 20, 2)}.items():
        if name == "c5cut" and args.no_c5cut:
            continue
        rows = np.tile(np.asarray(CANON, dtype=np.float64), (c, 1))
        rows[:, 1] = np.linspace(80.0, 120.0, c) if c > 1 else rows[:, 1]
        contracts = torch.tensor(rows, device=dev)
        per = b_total // world
        lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else b_total
        shard = fused_args(contracts, c, t, n, b_total, lo, hi, dtype)
        full = fused_args(contracts, c, t, n, b_total, 0, b_total, dtype)
        need = max(_cabi.LIB.smc_cf_fused_workspace_bytes(_cabi.byref(shard)), _cabi.LIB.smc_cf_fused_workspace_bytes(_cabi.byref(full)))
        work = torch.empty(int(need) + 4096, dtype=torch.uint8, device=dev)

        def time_it(fn):
            fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            barrier()
            return max_over_ranks(a.elapsed_time(b)) / steps

        ms_1 = time_it(lambda: _cabi.cf_fused(full, dev, dtype, work))
        ms_n = time_it(lambda: run_sharded(shard, work))
        total = float(c) * b_total * n * t
        out[name] = {"problem": f"{c} contract(s), T={t}, N={n}, B_total={b_total} (fixed), float32 RAW LOG_EULER", "steps": steps,
                     "ms_per_step_1gpu": ms_1, "ms_per_step": ms_n, "speedup": ms_1 / ms_n, "efficiency": ms_1 / ms_n / world,
                     "value": total / (ms_n * 1e-3), "unit": "path-steps/s"}
    return out


def secondary_workloads(_cabi, torch, dev, calib: dict, args) -> dict:
    """The other BASELINE.json configurations and variants on this GPU (N = 1 only), each with its own roofline."""
    import numpy as np

    def time_fused(c, t, n, b, dtype, scheme, norm, steps, contracts_rows=None, stream_version=0):
        rows = np.tile(np.asarray(CANON, dtype=np.float64), (c, 1)) if contracts_rows is None else contracts_rows
        contracts = torch.tensor(rows, device=dev)
        a = _cabi.make_fused_args(contracts, c, t, n, b, dtype, scheme, norm, 7, 0, stream_version=stream_version)
        work = torch.empty(int(_cabi.LIB.smc_cf_fused_workspace_bytes(_cabi.byref(a))) + 4096, dtype=torch.uint8, device=dev)
        _cabi.cf_fused(a, dev, dtype, work)
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.cf_fused(a, dev, dtype, work)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, int(_cabi.LIB.smc_cf_fused_launch_count(_cabi.byref(a)))

    out = {}
    c2 = WORKLOADS["c2"]
    steps2 = float(c2["T"]) * c2["N"] * c2["B"]
    ms, k = time_fused(1, c2["T"], c2["N"], c2["B"], torch.float32, _cabi.SMC_LOG_EULER, _cabi.SMC_NORMALIZE, 10)
    out["c2_normalize"] = {"ms": ms, "path_steps_per_sec": steps2 / (ms * 1e-3), "kernels": k,
                           "note": "two kernels: simulate + stage terminals (33.5 MB written), then payoff + CF from the staged terminals",
                           "roofline": {"bound": "xu", "frac": steps2 / (ms * 1e-3) * XU_OPS_PER_STEP / calib["mufu"], "unit": "of calibrated MUFU peak, whole step"}}
    ms, k = time_fused(1, c2["T"], c2["N"], c2["B"], torch.float32, _cabi.SMC_SIMPLE_EULER, _cabi.SMC_RAW, 10)
    out["c2_simple_euler"] = {"ms": ms, "path_steps_per_sec": steps2 / (ms * 1e-3), "kernels": k,
                              "roofline": {"bound": "xu", "frac": steps2 / (ms * 1e-3) * XU_OPS_PER_STEP_ALL_NORMALS / calib["mufu"],
                                           "unit": "of calibrated MUFU peak, whole step (2 XU ops per path-step: this scheme needs every normal, so COS + SIN stay)"}}
    # the opt-in Philox4x32-7 stream (smc_stream_version 1): NOT the stream `value` is measured on — a different sample set
    ms, k = time_fused(1, c2["T"], c2["N"], c2["B"], torch.float32, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, 10, stream_version=_cabi.SMC_STREAM_PHILOX7)
    out["c2_philox7_opt_in"] = {"ms": ms, "path_steps_per_sec": steps2 / (ms * 1e-3), "kernels": k,
                                "note": "opt-in stream_version=1 (seven Philox rounds; passes the same battery, tests/test_gpu_stream_battery.py); every other number in this line is the default Philox4x32-10 stream",
                                "roofline": {"bound": "xu", "frac": steps2 / (ms * 1e-3) * XU_OPS_PER_STEP / calib["mufu"], "unit": "of calibrated MUFU peak, whole step"}}
    # c4: 512 Sobol contracts, float64 (BASELINE configs[3]); FP64-pipe roofline from the live DFMA calibration
    c4 = WORKLOADS["c4"]
    from spectralmc_b200.gbm import BlackScholes
    from spectralmc_b200.sobol_sampler import BoundSpec, SobolConfig, SobolSampler, build_domain_bounds

    bounds = build_domain_bounds(BlackScholes.Inputs, {k_: BoundSpec(*b) for k_, b in dict(
        X0=(0.001, 10_000.0), K=(0.001, 20_000.0), T=(0.0, 10.0), r=(-0.2, 0.2), d=(-0.2, 0.2), v=(0.0, 2.0)).items()}).unwrap()
    sampler = SobolSampler.create(BlackScholes.Inputs, bounds, config=SobolConfig(seed=31, skip=0)).unwrap()
    rows4 = np.ascontiguousarray(sampler.sample_array(c4["contracts"]).unwrap() if hasattr(sampler, "sample_array") else
                                 np.asarray([[x.X0, x.K, x.T, x.r, x.d, x.v] for x in sampler.sample(c4["contracts"]).unwrap()]), dtype=np.float64)
    steps4 = float(c4["contracts"]) * c4["T"] * c4["N"] * c4["B"]
    ms, k = time_fused(c4["contracts"], c4["T"], c4["N"], c4["B"], torch.float64, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, 2, contracts_rows=rows4)
    out["c4"] = {"ms": ms, "path_steps_per_sec": steps4 / (ms * 1e-3), "cf_estimates_per_sec": c4["contracts"] / (ms * 1e-3), "kernels": k,
                 "workload": "512 Sobol contracts (seed 31), float64, T=365, N=256, B=4096, RAW, LOG_EULER",
                 "roofline": {"kernel": "smc::step_kernel<double, SRC_FUSED, LOG_EULER, OUT_COLSUM>", "bound": "fp64",
                              "achieved": steps4 / (ms * 1e-3) * FP64_OPS_PER_STEP / 1e12, "peak": calib["dfma"] / 1e12, "unit": "Tlane-op/s (DFMA+DMUL+DADD)",
                              "frac": steps4 / (ms * 1e-3) * FP64_OPS_PER_STEP / calib["dfma"], "fp64_ops_per_path_step": FP64_OPS_PER_STEP,
                              "peak_source": "calibrated live: smc_pipe_calibrate kind 3 (DFMA)"}}
    # c3: the trainer-test training step (1024 Sobol contracts, T=1, N=16, B=4096) through GbmCVNNPricer.train
    c3 = WORKLOADS["c3_trainer"]
    ms_sim, k = time_fused(c3["contracts"], c3["T"], c3["N"], c3["B"], torch.float32, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, 20)
    out["c3_trainer"] = {"simulation_ms": ms_sim, "simulation_kernels": k,
                         "simulation_path_steps_per_sec": float(c3["contracts"]) * c3["T"] * c3["N"] * c3["B"] / (ms_sim * 1e-3)}
    try:
        out["c3_trainer"].update(trainer_step(torch, c3))
    except Exception as exc:  # noqa: BLE001 - reported, not fatal for the headline
        out["c3_trainer"]["trainer_error"] = f"{type(exc).__name__}: {exc}"
    return out


def trainer_step(torch, c3: dict) -> dict:
    """Wall-clock per training step of GbmCVNNPricer.train (Sobol + H2D + targets + CVNN fwd/bwd/Adam)."""
    from spectralmc_b200.cvnn import make_cvnn
    from spectralmc_b200.effects import ForwardNormalization, PathScheme
    from spectralmc_b200.gbm import BlackScholes, BlackScholesConfig, SimulationParams
    from spectralmc_b200.gbm_trainer import GbmCVNNPricer, TrainingConfig
    from spectralmc_b200.numerical import Precision
    from spectralmc_b200.sobol_sampler import BoundSpec, build_domain_bounds

    bounds = build_domain_bounds(BlackScholes.Inputs, {k: BoundSpec(*b) for k, b in dict(
        X0=(0.001, 10_000.0), K=(0.001, 20_000.0), T=(0.0, 10.0), r=(-0.2, 0.2), d=(-0.2, 0.2), v=(0.0, 2.0)).items()}).unwrap()
    sp = SimulationParams(timesteps=c3["T"], network_size=c3["N"], batches_per_mc_run=c3["B"], threads_per_block=256, mc_seed=42,
                          buffer_size=1, dtype=Precision.float32)
    cfg = BlackScholesConfig(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
    pricer = GbmCVNNPricer(cfg, bounds, make_cvnn(6, c3["N"], seed=42))
    pricer.train(TrainingConfig(num_batches=3, batch_size=c3["contracts"])).unwrap()
    torch.cuda.synchronize()
    steps, best = 100, float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        losses = pricer.train(TrainingConfig(num_batches=steps, batch_size=c3["contracts"])).unwrap().losses
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / steps)
    return {"training_step_ms": best * 1e3, "cf_estimates_per_sec": c3["contracts"] / best, "last_loss": float(losses[-1]),
            "api": "GbmCVNNPricer.train (Sobol batch + pinned H2D + smc_cf_fused + CVNN forward/backward/Adam as one CUDA graph), wall clock"}


def numba_reference_leg(torch, value: float, args) -> dict:
    """The ">= 10x the reference" comparison in the driver's own record: the reference's Numba kernel, TAKEN FROM ITS
    SOURCE (baseline/_ref/spectralmc_gbm.py is a copy of /root/reference/src/spectralmc/gbm.py made by
    __graft_entry__.build(); the function `SimulateBlackScholes`, gbm.py:224-257, is cut out by AST and compiled
    with numba.cuda.jit as the reference does), launched as gbm.py:413-426 does (total_blocks = ceil(P/256), 256
    threads, in place), followed by the reference's per-contract pipeline with its host synchronisations
    (gbm.py:434,442,476; gbm_trainer.py:814-817).  CuPy is not installable here, so its array calls are replaced
    1:1 by torch ops: torch.randn (cuRAND Philox) for cp.random...standard_normal (XORWOW), torch.fft.fft (cuFFT,
    the library CuPy calls) and torch.mean."""
    src_path = os.path.join(ROOT, "baseline", "_ref", "spectralmc_gbm.py")
    if args.no_gpu_reference:
        return {"skipped": "--no-gpu-reference"}
    if not os.path.exists(src_path):
        return {"skipped": f"{os.path.relpath(src_path, ROOT)} is missing (made by __graft_entry__.build() where /root/reference exists)"}
    try:
        import ast
        from math import exp, sqrt  # noqa: F401 - names the kernel body uses

        from numba import cuda
    except Exception as exc:  # noqa: BLE001
        return {"skipped": f"numba.cuda not importable: {type(exc).__name__}: {exc}"}
    try:
        fn = next(n for n in ast.parse(open(src_path).read()).body if isinstance(n, ast.FunctionDef) and n.name == "SimulateBlackScholes")
        fn.decorator_list, fn.returns = [], None
        for a in fn.args.args:
            a.annotation = None
        ns: dict = {"cuda": cuda}
        exec("from math import exp, sqrt\n" + ast.unparse(fn), ns)  # noqa: S102 - the reference's own kernel body
        kernel = cuda.jit(ns["SimulateBlackScholes"])
        w = WORKLOADS["c2"]
        T, N, B = w["T"], w["N"], w["B"]
        P = N * B
        X0, K, Tm, r, d, v = CANON
        dt = Tm / T

        def one_contract(normalize: bool):
            sims = torch.randn((T, P), dtype=torch.float32, device="cuda")
            torch.cuda.synchronize()  # the generator's stream.synchronize() before hand-out (async_normals.py:230)
            kernel[(P + 255) // 256, 256](cuda.as_cuda_array(sims), T, dt, X0, r, d, v, True)  # gbm.py:413-426
            times = torch.linspace(dt, Tm, T, dtype=torch.float32, device="cuda")              # gbm.py:429
            forwards = X0 * torch.exp((r - d) * times)                                         # gbm.py:430
            df = torch.exp(-r * times)                                                         # gbm.py:431
            cuda.synchronize()                                                                 # gbm.py:434
            if normalize:
                row_means = sims.mean(dim=1, keepdim=True).squeeze()                           # gbm.py:437
                sims *= (forwards / row_means).unsqueeze(1)                                    # gbm.py:438
            torch.cuda.synchronize()                                                           # gbm.py:442
            terminal = sims[-1]
            put = df[-1] * torch.clamp(K - terminal, min=0)                                    # gbm.py:473
            call = df[-1] * torch.clamp(terminal - K, min=0)  # noqa: F841                       gbm.py:474
            torch.cuda.synchronize()                                                           # gbm.py:476
            cf = torch.fft.fft(put.reshape(B, N), dim=1).mean(dim=0)                           # gbm_trainer.py:814-817
            torch.cuda.synchronize()                                                           # gbm_trainer.py:1553
            return cf

        out = {"kernel_source": "baseline/_ref/spectralmc_gbm.py: SimulateBlackScholes (reference gbm.py:224-257), cut out by AST, numba.cuda.jit",
               "launch": "grid ceil(P/256) x 256 threads, in place (gbm.py:413-426); host syncs of gbm.py:434,442,476",
               "normals": "torch.randn (cuRAND Philox) stands in for CuPy's XORWOW generator; torch.fft.fft is cuFFT as in CuPy",
               "workload": "c2: one contract per pass, float32, T=252, N=128, B=65536", "reps": 5}
        for name, normalize in (("raw", False), ("normalize", True)):
            for _ in range(2):
                cf = one_contract(normalize)
            times = []
            for _ in range(5):
                t0 = time.perf_counter()
                cf = one_contract(normalize)
                times.append(time.perf_counter() - t0)
            best = min(times)
            out[name] = {"ms_per_contract": best * 1e3, "ms_per_contract_median": sorted(times)[2] * 1e3,
                         "path_steps_per_sec": T * P / best, "put_price": float(cf[0].real) / N}
        out["ms_per_contract"] = out["raw"]["ms_per_contract"]
        out["path_steps_per_sec"] = out["raw"]["path_steps_per_sec"]
        out["ratio"] = value / out["raw"]["path_steps_per_sec"]
        out["ratio_note"] = "this run's `value` (RAW, device-timed) / the reference pipeline's RAW rate (best of 5, wall clock)"
        return out
    except Exception as exc:  # noqa: BLE001 - a baseline that cannot run is reported, it does not fail the bench
        return {"skipped": f"{type(exc).__name__}: {exc}"}


def materialised_roofline(_cabi, torch, dev, peaks, T, N, dtype) -> dict:
    """HBM-bound kernels of the materialised-normals mode, timed alone with CUDA events.
    Algorithmic bytes: generator sizeof(real)/normal written; in-place stepper 2*sizeof(real)
    per path-step; terminal-only stepper sizeof(real) per path-step."""
    B = 16384  # 2.1M paths x 252 steps x 4 B = 2.1 GB per matrix: >> L2
    P = N * B
    z = torch.empty((T, P), dtype=dtype, device=dev)
    size = z.element_size()
    hbm = peaks.get("hbm_gbs")
    out = {"peak": hbm, "unit": "GB/s", "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm else "unavailable", "matrix_bytes": T * P * size}

    def time_it(fn, reps=5):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    ms = time_it(lambda: _cabi.philox_normals(z, 7, 0))
    out["philox_normals"] = {"bound": "hbm", "achieved": T * P * size / ms / 1e6, "ms": ms, "bytes": T * P * size}
    ms = time_it(lambda: _cabi.gbm_terminal_from_normals(z, 1.0 / T, *CANON[:1], CANON[3], CANON[4], CANON[5], _cabi.SMC_LOG_EULER))
    out["gbm_terminal_from_normals"] = {"bound": "hbm", "achieved": T * P * size / ms / 1e6, "ms": ms, "bytes": T * P * size}
    ms = time_it(lambda: _cabi.gbm_paths_inplace(z, 1.0 / T, CANON[0], CANON[3], CANON[4], CANON[5], _cabi.SMC_LOG_EULER, 256))
    out["gbm_paths_inplace"] = {"bound": "hbm", "achieved": 2 * T * P * size / ms / 1e6, "ms": ms, "bytes": 2 * T * P * size}
    if hbm:
        for k in ("philox_normals", "gbm_terminal_from_normals", "gbm_paths_inplace"):
            out[k]["frac"] = out[k]["achieved"] / hbm
    del z
    return out


def cpu_baseline(w: dict, args) -> dict:
    cores = os.cpu_count() or 1
    sample_B = cpu_sample_rows(w, args, cores) * 4
    cpu_pass(w, sample_B, cores, 0)
    reps, total = 0, 0.0
    while total < 10.0 and reps < 40:
        total += cpu_pass(w, sample_B, cores, 1 + reps)
        reps += 1
    value = reps * sample_B * w["N"] * w["T"] / total
    return {"value": value, "unit": "path-steps/s", "cores": cores, "kind": "port",
            "sample": f"{reps} passes over {sample_B} of {w['B']} batch rows ({sample_B * w['N']} paths x {w['T']} steps each), {total:.1f} s; "
                      f"oracle/gbm_oracle.c (plain C, float64 path arithmetic as the reference kernel) on {cores} threads"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--collective", default="auto", choices=("auto", "p2p", "nccl"),
                    help="multi-GPU exchange: fused peer-memory exchange, NCCL all-reduce, or the former with fallback to the latter")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained leg in seconds")
    ap.add_argument("--cpu-sample-batches", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the secondary workloads and the GPU reference baseline")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-c5cut", action="store_true", help="N > 1: skip the 64-contract x 2^20-row strong-scaling problem")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
