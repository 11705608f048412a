#!/usr/bin/env python
"""Generate tests/golden/cvnn_*.npz by RUNNING THE REFERENCE's CVNN classes (build container only).

    python tests/golden/make_golden_cvnn.py

What runs: ``spectralmc.cvnn.{ComplexLinear, modReLU, zReLU, ComplexSequential}`` imported
unmodified from /root/reference/src on CPU torch (the module asserts CUDA at import through
``spectralmc.runtime.get_torch_handle``; the handle cache is pre-seeded), driven
by the five statements of ``GbmCVNNPricer._torch_step`` (gbm_trainer.py:828-834: two MSE terms,
``zero_grad``, ``backward``, ``optimizer.step``) with ``optim.Adam(params, lr)`` as the trainer
builds it (gbm_trainer.py:1513).  Networks are seeded inside ``torch.random.fork_rng`` as
``build_model`` does (cvnn_factory.py:343-368) and nested the way the factory nests them
(``_maybe_activate``/``_maybe_project``, cvnn_factory.py:185-190).

Each fixture holds: the layer list, the initial parameters (``parameters()`` order), three
batches of inputs/targets, the step-0 prediction and gradients, the three losses, and the
parameters after three Adam steps.  Nothing here is imported by tests or the product.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
STEPS = 3
LR = 1e-2


def describe(module, cvnn) -> list:
    """Flatten nested ComplexSequential containers into the oracle's layer list."""
    if isinstance(module, cvnn.ComplexSequential):
        return [d for m in module.layers for d in describe(m, cvnn)]
    if isinstance(module, cvnn.ComplexLinear):
        return [["linear", module.in_features, module.out_features, module.real_bias is not None]]
    if isinstance(module, cvnn.modReLU):
        return [["modrelu", int(module.bias.numel())]]
    if isinstance(module, cvnn.zReLU):
        return [["zrelu"]]
    raise TypeError(type(module))


def run_case(name: str, make_net, make_batch, dtype: torch.dtype, cvnn) -> None:
    torch.set_default_dtype(dtype)
    with torch.random.fork_rng():
        torch.manual_seed(11)
        net = make_net()
    net = net.to(dtype)
    layers = describe(net, cvnn)
    # biases start at zero in the reference; perturb them so the bias paths are exercised
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.add_(0.3 * torch.randn(p.shape, generator=gen, dtype=dtype))
    params0 = [p.detach().clone().numpy() for p in net.parameters()]
    adam = torch.optim.Adam(net.parameters(), lr=LR)  # gbm_trainer.py:1513
    out: dict[str, np.ndarray] = {"layers": np.array(json.dumps(layers)), "lr": np.array(LR)}
    for i, p in enumerate(params0):
        out[f"param0_{i}"] = p
    losses = []
    for s in range(STEPS):
        real_in, imag_in, targets = make_batch(s, dtype)
        pred_r, pred_i = net(real_in, imag_in)
        loss = torch.nn.functional.mse_loss(pred_r, torch.real(targets)) + torch.nn.functional.mse_loss(
            pred_i, torch.imag(targets)
        )  # gbm_trainer.py:828-830
        adam.zero_grad(set_to_none=True)
        loss.backward()
        if s == 0:
            out["pred_r"], out["pred_i"] = pred_r.detach().numpy(), pred_i.detach().numpy()
            for i, p in enumerate(net.parameters()):
                out[f"grad0_{i}"] = p.grad.detach().clone().numpy()
        adam.step()
        losses.append(float(loss.detach()))
        out[f"real_in_{s}"], out[f"imag_in_{s}"] = real_in.numpy(), imag_in.numpy()
        out[f"targets_{s}"] = targets.numpy()
    out["losses"] = np.array(losses, dtype=np.float64)
    for i, p in enumerate(net.parameters()):
        out[f"param{STEPS}_{i}"] = p.detach().clone().numpy()
    tag = "float32" if dtype == torch.float32 else "float64"
    np.savez_compressed(os.path.join(HERE, f"cvnn_{name}_{tag}.npz"), **out)
    print(f"cvnn_{name}_{tag}: layers={layers} losses={losses}")


def main() -> None:
    sys.path.insert(0, REF_SRC)
    sys.dont_write_bytecode = True
    import spectralmc.runtime.torch_runtime as tr

    tr._TORCH_HANDLE = torch  # skip the import-time CUDA assert (runtime/torch_runtime.py:86-99)
    import spectralmc.cvnn as cvnn
    from spectralmc.sobol_sampler import BoundSpec  # noqa: F401  (import check only)

    # ---- case "pricer": the test network of the reference (tests/helpers/factories.py:69-105):
    # Seq(Seq(Linear(6,32), modReLU(32)), Linear(32,16)); inputs = contract rows (imag = 0,
    # gbm_trainer.py:1775-1783) drawn from the default test domain, targets of CF magnitude.
    lower = np.array([1e-3, 1e-3, 0.0, -0.2, -0.2, 0.0])
    upper = np.array([1e4, 2e4, 10.0, 0.2, 0.2, 2.0])

    def pricer_net():
        return cvnn.ComplexSequential(
            cvnn.ComplexSequential(cvnn.ComplexLinear(6, 32), cvnn.modReLU(32)), cvnn.ComplexLinear(32, 16)
        )

    def pricer_batch(s, dtype):
        from scipy.stats.qmc import Sobol

        raw = Sobol(d=6, scramble=True, seed=42 + s).random(64)
        rows = lower + (upper - lower) * raw
        rng = np.random.default_rng(100 + s)
        t = rng.standard_normal((64, 16)) * 50.0 + 1j * rng.standard_normal((64, 16)) * 50.0
        cd = torch.complex64 if dtype == torch.float32 else torch.complex128
        real = torch.tensor(rows, dtype=dtype)
        return real, torch.zeros_like(real), torch.tensor(t, dtype=cd)

    # ---- case "deep": two hidden layers, zReLU and modReLU, a bias-free layer, complex inputs,
    # ragged widths and a batch that is not a multiple of any tile size.
    def deep_net():
        return cvnn.ComplexSequential(
            cvnn.ComplexSequential(cvnn.ComplexLinear(5, 19), cvnn.modReLU(19)),
            cvnn.ComplexSequential(cvnn.ComplexLinear(19, 33, bias=False), cvnn.zReLU()),
            cvnn.ComplexLinear(33, 7),
        )

    def deep_batch(s, dtype):
        rng = np.random.default_rng(200 + s)
        cd = torch.complex64 if dtype == torch.float32 else torch.complex128
        real = torch.tensor(rng.standard_normal((77, 5)), dtype=dtype)
        imag = torch.tensor(rng.standard_normal((77, 5)), dtype=dtype)
        t = rng.standard_normal((77, 7)) + 1j * rng.standard_normal((77, 7))
        return real, imag, torch.tensor(t, dtype=cd)

    for dtype in (torch.float64, torch.float32):
        run_case("pricer", pricer_net, pricer_batch, dtype, cvnn)
        run_case("deep", deep_net, deep_batch, dtype, cvnn)


if __name__ == "__main__":
    main()
