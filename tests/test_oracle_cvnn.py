"""The CVNN-step oracle (oracle/cvnn.py) against outputs of the reference's own ``spectralmc.cvnn``
classes (tests/golden/cvnn_*.npz, produced by tests/golden/make_golden_cvnn.py), plus the host
side of the C ABI's CVNN descriptor (no device needed)."""

from __future__ import annotations

import ctypes
import json
import os

import numpy as np
import pytest
import torch

from oracle import cvnn as ocvnn
from tests.conftest import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = [(n, t) for n in ("pricer", "deep") for t in ("float64", "float32")]
# float64: the two sides differ only in summation order; float32: same, at float32 width
TOL = {"float64": 1e-12, "float32": 1e-5}


def load_case(name: str, tag: str):
    z = np.load(os.path.join(GOLDEN, f"cvnn_{name}_{tag}.npz"))
    layers = [tuple(layer) for layer in json.loads(str(z["layers"]))]
    n = len([k for k in z.files if k.startswith("param0_")])
    return z, layers, n


def nw(a, b) -> float:
    """Norm-wise relative error max|a - b| / max|b|."""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(float(np.abs(b).max()), 1e-300))


@pytest.mark.parametrize("name,tag", CASES)
def test_forward_loss_and_gradients_match_the_reference(name, tag) -> None:
    z, layers, n = load_case(name, tag)
    p0 = [z[f"param0_{i}"] for i in range(n)]
    loss, grads, (pr, pi) = ocvnn.loss_and_grads(layers, p0, z["real_in_0"], z["imag_in_0"], z["targets_0"])
    assert nw(pr, z["pred_r"]) <= TOL[tag] and nw(pi, z["pred_i"]) <= TOL[tag]
    assert abs(loss - z["losses"][0]) <= TOL[tag] * z["losses"][0]
    for i, g in enumerate(grads):
        assert g.dtype == p0[i].dtype and nw(g, z[f"grad0_{i}"]) <= TOL[tag], i


@pytest.mark.parametrize("name,tag", CASES)
def test_three_adam_steps_match_the_reference(name, tag) -> None:
    z, layers, n = load_case(name, tag)
    p0 = [z[f"param0_{i}"] for i in range(n)]
    xs = [z[f"real_in_{s}"] for s in range(3)], [z[f"imag_in_{s}"] for s in range(3)], [z[f"targets_{s}"] for s in range(3)]
    losses, params = ocvnn.train_steps(layers, p0, *xs, lr=float(z["lr"]), steps=3)
    for got, ref in zip(losses, z["losses"]):
        assert abs(got - ref) <= 10 * TOL[tag] * ref
    for i, p in enumerate(params):
        assert nw(p, z[f"param3_{i}"]) <= 10 * TOL[tag], i


def test_golden_layer_lists_are_what_the_product_describes() -> None:
    """``spectralmc_b200.cvnn.describe`` flattens nested containers into the golden's layer list."""
    from spectralmc_b200 import cvnn

    net = cvnn.ComplexSequential(cvnn.ComplexSequential(cvnn.ComplexLinear(5, 19), cvnn.modReLU(19)),
                                 cvnn.ComplexSequential(cvnn.ComplexLinear(19, 33, bias=False), cvnn.zReLU()), cvnn.ComplexLinear(33, 7))
    _, layers, n = load_case("deep", "float64")
    assert [tuple(x) for x in cvnn.describe(net)] == layers
    assert len(list(net.parameters())) == n
    assert [tuple(p.shape) for p in net.parameters()] == [np.load(os.path.join(GOLDEN, "cvnn_deep_float64.npz"))[f"param0_{i}"].shape for i in range(n)]
    assert cvnn.describe(torch.nn.Linear(3, 3)) is None
    assert cvnn.describe(cvnn.ComplexSequential(cvnn.ComplexLinear(2, 2), torch.nn.ReLU())) is None


def test_product_layers_reproduce_the_reference_forward_on_cpu() -> None:
    """The torch route of the product's layer classes == the reference's classes (float64, CPU)."""
    from spectralmc_b200 import cvnn

    z, layers, n = load_case("deep", "float64")
    net = cvnn.ComplexSequential(cvnn.ComplexSequential(cvnn.ComplexLinear(5, 19), cvnn.modReLU(19)),
                                 cvnn.ComplexSequential(cvnn.ComplexLinear(19, 33, bias=False), cvnn.zReLU()), cvnn.ComplexLinear(33, 7)).double()
    with torch.no_grad():
        for i, p in enumerate(net.parameters()):
            p.copy_(torch.from_numpy(z[f"param0_{i}"]))
        pr, pi = net(torch.from_numpy(z["real_in_0"]), torch.from_numpy(z["imag_in_0"]))
    assert nw(pr.numpy(), z["pred_r"]) <= 1e-13 and nw(pi.numpy(), z["pred_i"]) <= 1e-13


def test_xavier_initialisation_matches_the_reference_stream() -> None:
    """Same seed, same layer sizes -> the reference's initial weights (cvnn.py:107-114 draws
    real_weight then imag_weight with xavier_uniform_)."""
    from spectralmc_b200 import cvnn

    z, _, n = load_case("pricer", "float64")
    torch.set_default_dtype(torch.float64)
    try:
        with torch.random.fork_rng():
            torch.manual_seed(11)
            net = cvnn.ComplexSequential(cvnn.ComplexSequential(cvnn.ComplexLinear(6, 32), cvnn.modReLU(32)), cvnn.ComplexLinear(32, 16))
    finally:
        torch.set_default_dtype(torch.float32)
    for i, p in enumerate(net.parameters()):
        if p.dim() == 2:  # the golden perturbs the (zero-initialised) biases afterwards
            assert np.array_equal(p.detach().numpy(), z[f"param0_{i}"]), i


# ----------------------------------------------------------------------------- descriptor (host only)
def test_descriptor_offsets_follow_parameters_order() -> None:
    from spectralmc_b200 import _cabi

    layers = [("linear", 6, 32, True), ("modrelu", 32), ("linear", 32, 16, False), ("zrelu",)]
    net, n = _cabi.make_cvnn_net(layers, 6, torch.float32)
    assert n == 2 * 6 * 32 + 64 + 32 + 2 * 32 * 16
    assert [net.layers[i].param_offset for i in range(4)] == [0, 448, 480, 1504]
    assert _cabi.cvnn_output_width(net) == 16
    small, big = _cabi.cvnn_workspace_bytes(net, 64, False), _cabi.cvnn_workspace_bytes(net, 64, True)
    assert 0 < small < big


def test_descriptor_validation_needs_no_device() -> None:
    from spectralmc_b200 import _cabi

    bad, _ = _cabi.make_cvnn_net([("linear", 6, 32, True), ("linear", 31, 4, True)], 6, torch.float32)
    assert _cabi.LIB.smc_cvnn_workspace_bytes(ctypes.byref(bad), 8, 1) == 0
    assert b"expects 31 inputs" in _cabi.LIB.smc_last_error()
    with pytest.raises(_cabi.SmcError):
        _cabi.cvnn_output_width(bad)
    net, n = _cabi.make_cvnn_net([("linear", 6, 4, True)], 6, torch.float64)
    net.n_params = n - 1
    assert _cabi.LIB.smc_cvnn_output_width(ctypes.byref(net)) == -1 and b"outside the buffer" in _cabi.LIB.smc_last_error()
    net.n_params = n
    rc = _cabi.LIB.smc_cvnn_forward(ctypes.byref(net), None, None, None, 4, None, None, None, 0, None)
    assert rc == 1 and b"NULL" in _cabi.LIB.smc_last_error()
    rc = _cabi.LIB.smc_cvnn_forward(ctypes.byref(net), 16, 16, 16, 4, 16, 16, 16, 8, None)
    assert rc == 3 and b"workspace" in _cabi.LIB.smc_last_error()
    hyper = _cabi.AdamArgs(1e-2, 1.5, 0.999, 1e-8)
    rc = _cabi.LIB.smc_adam_step(16, 16, 16, 16, 4, 0, 16, ctypes.byref(hyper), None)
    assert rc == 1 and b"hyper" in _cabi.LIB.smc_last_error()
    with pytest.raises(ValueError):
        _cabi.make_cvnn_net([("batchnorm", 4)], 4, torch.float32)
