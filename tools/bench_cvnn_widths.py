#!/usr/bin/env python
"""CVNN training step at realistic widths: the C-ABI step (smc_cvnn_train_step: SIMT FP32 complex GEMM with fused
epilogues, hand-derived backward, fused MSE and Adam) against torch autograd + smc_adam_step (cuBLAS SGEMM, TF32 off
as the reference trains, runtime/torch_runtime.py:76-77), BOTH replayed as one CUDA graph — the two routes
GbmCVNNPricer takes (gbm_trainer.py: _StepGraph / _TorchStepGraph).

    python tools/bench_cvnn_widths.py > profiles/r2_cvnn_widths.jsonl

Network: 6 -> hidden (modReLU) -> hidden (modReLU) -> N=128, float32.  Prints one JSON line per (hidden, rows).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

from spectralmc_b200.cvnn import ComplexLinear, ComplexSequential, FlatAdam, FusedCVNN, modReLU
from spectralmc_b200.gbm_trainer import _StepGraph, _TorchStepGraph

dev = torch.device("cuda", 0)
N = 128


def build(hidden: int, seed: int = 3) -> ComplexSequential:
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        net = ComplexSequential(ComplexSequential(ComplexLinear(6, hidden), modReLU(hidden)),
                                ComplexSequential(ComplexLinear(hidden, hidden), modReLU(hidden)), ComplexLinear(hidden, N))
    return net.to(device=dev, dtype=torch.float32)


def time_graph(g, reps: int = 50) -> float:
    for _ in range(5):
        g.graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.graph.replay()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


for hidden in (32, 128, 256, 512, 1024):
    for rows in (1024, 8192):
        real = torch.randn(rows, 6, device=dev)
        targets = torch.randn(rows, N, dtype=torch.complex64, device=dev)
        fused = FusedCVNN(build(hidden), lr=1e-3)
        g1 = _StepGraph(fused, rows)
        g1.real_in.copy_(real); g1.targets.copy_(targets)
        ms_fused = time_graph(g1)
        net = build(hidden)
        adam = FlatAdam(list(net.parameters()), lr=1e-3)
        g2 = _TorchStepGraph(net, adam, rows, 6, N, torch.float32)
        g2.real_in.copy_(real); g2.targets.copy_(targets)
        ms_torch = time_graph(g2)
        # complex GEMM work: 8 real flops per complex MAC, forward + two backward GEMMs per linear layer
        flops = 3 * 8.0 * rows * (6 * hidden + hidden * hidden + hidden * N)
        print(json.dumps({"hidden": hidden, "rows": rows, "fused_ms": round(ms_fused, 4), "torch_graph_ms": round(ms_torch, 4),
                          "fused_over_torch": round(ms_fused / ms_torch, 3), "gemm_gflop_per_step": round(flops / 1e9, 3),
                          "fused_tflops": round(flops / ms_fused / 1e9, 2), "torch_tflops": round(flops / ms_torch / 1e9, 2),
                          "loss_fused": float(g1.loss), "loss_torch": float(g2.loss)}), flush=True)
        del fused, g1, g2, net, adam
