#!/usr/bin/env python
"""Generate tests/golden/cvnnfactory_*.npz by RUNNING THE REFERENCE's ``spectralmc.cvnn_factory``.

    python tests/golden/make_golden_cvnn_factory.py          (build container only)

``build_model`` (cvnn_factory.py:343-368) and every block of ``spectralmc.cvnn`` it can emit
(ComplexLinear, zReLU, modReLU, NaiveComplexBatchNorm, CovarianceComplexBatchNorm, ComplexSequential,
ComplexResidual) run unmodified on CPU torch; the third-party stand-ins are those of make_golden.py
(a NumPy-backed ``cupy`` so that ``spectralmc.models`` imports, and the pre-seeded torch handle).

The fixture records, for one configuration that uses every layer kind, in float64 and float32:
the ``state_dict`` (names, shapes, seeded initial values, buffers), a training-mode forward on one
batch, the running statistics after it, and an eval-mode forward.  Nothing here is imported by tests
or the product; tests read only the .npz files.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True


def main() -> None:
    import make_golden as mg

    mg.install_shims()
    import torch

    import spectralmc.cvnn_factory as f
    from spectralmc.models.torch import FullPrecisionDType

    def config(dtype):
        act = lambda k: f.ActivationCfg(kind=k)  # noqa: E731
        layers = [
            f.LinearCfg(width=f.ExplicitWidth(value=8), activation=act(f.ActivationKind.MOD_RELU)),
            f.NaiveBNCfg(),
            f.ResidualCfg(
                body=f.SequentialCfg(layers=[f.LinearCfg(width=f.ExplicitWidth(value=8), activation=act(f.ActivationKind.Z_RELU)),
                                             f.CovBNCfg(momentum=0.2)]),
                activation=act(f.ActivationKind.MOD_RELU),
            ),
            f.ResidualCfg(body=f.SequentialCfg(layers=[f.LinearCfg(width=f.ExplicitWidth(value=5), bias=False)])),  # auto projection 8 -> 5
            f.SequentialCfg(layers=[f.LinearCfg(), f.CovBNCfg(affine=False)], activation=act(f.ActivationKind.Z_RELU)),
        ]
        return f.build_cvnn_config(dtype=dtype, layers=layers, seed=17, final_activation=act(f.ActivationKind.MOD_RELU)).unwrap()

    for tag, enum, tdt in (("float64", FullPrecisionDType.float64, torch.float64), ("float32", FullPrecisionDType.float32, torch.float32)):
        net = f.build_model(n_inputs=6, n_outputs=4, cfg=config(enum)).unwrap()
        out: dict[str, np.ndarray] = {}
        names = list(net.state_dict().keys())
        out["names"] = np.array(names)
        out["param_names"] = np.array([n for n, _ in net.named_parameters()])
        for i, (k, v) in enumerate(net.state_dict().items()):
            out[f"init_{i}"] = v.detach().clone().numpy()
        rng = np.random.default_rng(8)
        real = torch.tensor(rng.standard_normal((37, 6)) * 3.0 + 1.0, dtype=tdt)
        imag = torch.tensor(rng.standard_normal((37, 6)) * 0.5 - 2.0, dtype=tdt)
        net.train()
        with torch.no_grad():
            tr, ti = net(real, imag)
        out["real_in"], out["imag_in"] = real.numpy(), imag.numpy()
        out["train_r"], out["train_i"] = tr.numpy(), ti.numpy()
        for i, (k, v) in enumerate(net.state_dict().items()):
            out[f"after_{i}"] = v.detach().clone().numpy()
        net.eval()
        with torch.no_grad():
            er, ei = net(real[:5], imag[:5])
        out["eval_r"], out["eval_i"] = er.numpy(), ei.numpy()
        out["repr"] = np.array(repr(net))
        np.savez_compressed(os.path.join(HERE, f"cvnnfactory_all_layers_{tag}.npz"), **out)
        print(tag, len(names), "state entries;", "output", tuple(tr.shape))
        if tag == "float64":
            print(repr(net))


if __name__ == "__main__":
    main()
