#!/usr/bin/env python
"""Headline benchmark: GBM path-steps/s (and CF estimates/s) of the batch-generation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic contracts: Philox normals ->
GBM stepping -> payoff -> FFT/mean, i.e. one CF training target per contract
(reference: gbm_trainer.py:1546-1553 for one `_run_batch`).

Workload (BASELINE.json configs[1], "c2"): fp32, 1 contract (X0=100,K=100,T=1,r=.05,d=0,v=.2),
252 timesteps, network_size=128, batches_per_mc_run=65 536 per GPU, RAW, LOG_EULER.
Multi-GPU: weak scaling — every rank simulates 65 536 batch rows of the same contract batch
(global path counters, B_total = 65 536 * N), then ONE NCCL all-reduce of the complex partial CF
sums.

The JSON line carries: `value` (device-resident inputs, kernel-only), `e2e` (host buffers through
the C-ABI host entry point: pinned H2D of the contracts, D2H of the targets, inside the timed
region), `roofline` for the dominant kernel (the fused simulator: bound by FP32-issue/XU pipes, not
HBM or tensor — peaks calibrated live by two microbenchmarks), `roofline_materialised` (the
HBM-bound generator and stepper of the materialised-normals mode against MEASURED_PEAKS.json),
`cpu_baseline` (the oracle port on the host cores, bounded sample) and `clocks`.
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
WORKLOADS = {
    # name: (contracts, timesteps, network_size, batches per GPU, dtype)
    "c2": dict(contracts=1, T=252, N=128, B=65536, dtype="float32"),
    "c2x8": dict(contracts=8, T=252, N=128, B=65536, dtype="float32"),
    "c3_trainer": dict(contracts=1024, T=1, N=16, B=4096, dtype="float32"),
    "c4": dict(contracts=512, T=365, N=256, B=4096, dtype="float64"),
    # BASELINE configs[4] on 8 GPUs: 4096 contracts x 2^20 batch rows (131 072 per GPU), 1.39e14 path-steps per step
    "c5": dict(contracts=4096, T=252, N=128, B=131072, dtype="float32"),
}
# SASS-counted work per fp32 path-step of the fused log-Euler kernel (profiles/ has the listing):
ISSUE_SLOTS_PER_STEP = 163.0 / 12.0  # warp-instructions issued per path-step per lane (163 per two 6-normal blocks, SASS: profiles/r1_fused_f32_sass_inner_loop.txt)
XU_OPS_PER_STEP = 2.0  # (LG2 + SQRT + SIN + COS) per Box–Muller pair / 2 normals; log-sum variant


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
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


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

    def make_args(step: int, contracts):
        return _cabi.make_fused_args(contracts, C, T, N, B_total, dtype, _cabi.SMC_LOG_EULER, _cabi.SMC_RAW, 7, step * C,
                                     batch_begin=rank * B, batch_end=(rank + 1) * B)

    ws = torch.empty(_cabi.cf_fused_host_workspace_bytes(make_args(0, None)) + 4096, dtype=torch.uint8, device=dev)
    # multi-GPU exchange of the partial sums: fused into the finalise kernel over peer memory
    # (smc_cf_fused_p2p) when the IPC set-up succeeds, else one ncclAllReduce after the kernels
    exchange, collective = None, "none"
    if world > 1:
        collective = "nccl allreduce"
        if args.collective in ("auto", "p2p"):
            try:
                from spectralmc_b200.distributed import PeerExchange

                exchange = PeerExchange(C, N)
                collective = "fused peer-memory exchange (smc_cf_fused_p2p)"
            except Exception as exc:  # noqa: BLE001 - any set-up failure falls back to the NCCL route, visibly
                if args.collective == "p2p":
                    raise
                collective = f"nccl allreduce (peer exchange unavailable: {type(exc).__name__}: {exc})"
        ok = torch.tensor([1 if exchange is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # every rank must take the same route
        if int(ok.item()) == 0 and exchange is not None:
            exchange, collective = None, "nccl allreduce (peer exchange unavailable on another rank)"

    def sharded(step: int, contracts):
        a = make_args(step, contracts)
        if exchange is not None:
            return _cabi.cf_fused_p2p(a, exchange.next_group(), dev, dtype, ws)
        out = _cabi.cf_fused(a, dev, dtype, ws)
        if world > 1:
            dist.all_reduce(torch.view_as_real(out))
        return out

    launches = {"n": 0}
    plan_launches = int(_cabi.LIB.smc_cf_fused_launch_count(_cabi.byref(make_args(0, None))))

    def step_device(step: int):
        out = sharded(step, contracts_dev)
        launches["n"] += plan_launches
        return out

    def step_e2e(step: int):
        if world == 1:
            _cabi.cf_fused_host(make_args(step, None), contracts_pin, out_pin, ws)  # H2D + kernels + D2H + sync
        elif exchange is not None and contracts_pin.numel() * 8 <= 65536 and out_pin.numel() * out_pin.element_size() <= 65536:
            # small pinned buffers: the kernels read the contracts and write the targets through the buffers'
            # device aliases (as smc_cf_fused_host does on one GPU) — no staging copies
            _cabi.cf_fused_p2p(make_args(step, contracts_pin), exchange.next_group(), dev, dtype, ws, out=out_pin)
            torch.cuda.current_stream().synchronize()
        else:
            cdev = contracts_pin.to(dev, non_blocking=True)
            out = sharded(step, cdev)
            out_pin.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        launches["n"] += plan_launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps: int, first_step: int) -> float:
        """Sum of per-step CUDA-event durations (ms); L2 flushed between steps, outside the events."""
        pairs = []
        barrier()
        for s in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(first_step + s)
            b.record()
            pairs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in pairs)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for s in range(args.warmup):
        step_device(s)
        step_e2e(s)
    barrier()
    launches["n"] = 0  # count only what the timed regions launch

    with ClockSampler(local) as clocks:
        ms_dev = timed(step_device, args.steps, 1000)
        ms_e2e = timed(step_e2e, args.steps, 2000)
    clk = clocks.summary()

    path_steps_per_step = float(C) * B_total * N * T
    value = path_steps_per_step * args.steps / (ms_dev * 1e-3)
    e2e_value = path_steps_per_step * args.steps / (ms_e2e * 1e-3)

    line = {
        "metric": "gbm_path_steps_per_sec", "value": value, "unit": "path-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dtype == torch.float32 else "f64", "data": "synthetic",
        "config": dict(workload_config(args.workload, world), collective=collective),
        "cf_estimates_per_sec": C * args.steps / (ms_dev * 1e-3),
        "e2e": {"value": e2e_value, "unit": "path-steps/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(contracts_pin.numel() * 8), "d2h_bytes_per_step": int(out_pin.numel() * out_pin.element_size()),
                "api": "smc_cf_fused_host (C ABI, pinned host buffers)" if world == 1 else
                       (f"smc_cf_fused_p2p on pinned host buffers (contracts read and targets written through their device aliases) [{collective}]"
                        if exchange is not None and contracts_pin.numel() * 8 <= 65536 and out_pin.numel() * out_pin.element_size() <= 65536
                        else f"H2D + sharded fused path [{collective}] + D2H")},
        "gpu_launches": launches["n"], "clocks": clk,
    }

    if rank == 0:
        # ---- rooflines (rank 0, after the timed region) ----
        peaks = measured_peaks()
        calib = {}
        for kind, name in ((0, "ffma"), (1, "mufu"), (2, "philox")):
            ops, ms = _cabi.pipe_calibrate(kind, 1 << 15, dev)
            calib[name] = ops / (ms * 1e-3)
        per_gpu_rate = value / world
        issue_ach = per_gpu_rate * ISSUE_SLOTS_PER_STEP
        xu_ach = per_gpu_rate * XU_OPS_PER_STEP
        fr_issue, fr_xu = issue_ach / calib["ffma"], xu_ach / calib["mufu"]
        bound = "fp32_issue" if fr_issue >= fr_xu else "xu"
        line["roofline"] = {
            "kernel": "smc::tile_kernel<float, SRC_FUSED, LOG_EULER, OUT_COLSUM>",
            "bound": bound,
            "achieved": (issue_ach if bound == "fp32_issue" else xu_ach) / 1e12,
            "peak": (calib["ffma"] if bound == "fp32_issue" else calib["mufu"]) / 1e12,
            "unit": "Tlane-op/s", "frac": max(fr_issue, fr_xu),
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full` launch at this
            # workload (profiles/r1_fused_f32_ncu_raw.csv): 23 808 B read, 4 096 B written; algorithmic input 48 B
            "traffic": 27904 if args.workload == "c2" else None,
            "peak_source": "calibrated live: smc_pipe_calibrate (FFMA issue-rate and MUFU.EX2 microbenchmarks); MEASURED_PEAKS.json holds only HBM/bf16",
            "detail": {"issue_slots_per_path_step": ISSUE_SLOTS_PER_STEP, "xu_ops_per_path_step": XU_OPS_PER_STEP,
                       "fp32_issue": {"achieved": issue_ach / 1e12, "peak": calib["ffma"] / 1e12, "frac": fr_issue},
                       "xu": {"achieved": xu_ach / 1e12, "peak": calib["mufu"] / 1e12, "frac": fr_xu},
                       "philox_blocks_per_s_peak": calib["philox"]},
        }
        line["roofline_materialised"] = materialised_roofline(_cabi, torch, dev, peaks, T, N, dtype)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(w, args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if exchange is not None:
            exchange.close()
        dist.destroy_process_group()


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
    ap.add_argument("--cpu-sample-batches", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
