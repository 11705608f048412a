"""The fused CVNN step behind the C ABI against the golden outputs of the
reference's own ``spectralmc.cvnn`` classes, the oracle, and the product's torch route.

Tolerances (norm-wise, max|a-b| / max|b|): float64 1e-12, float32 1e-5 — the two sides differ in
summation order only (the reference's matmuls are cuBLAS/MKL, ours a fixed k-order FMA chain)."""

from __future__ import annotations

import json
import os

import numpy as np
import pytest
import torch

from oracle import cvnn as ocvnn
from spectralmc_b200 import _cabi, cvnn
from spectralmc_b200.effects import ForwardNormalization, PathScheme
from spectralmc_b200.gbm import BlackScholes
from spectralmc_b200.gbm_trainer import GbmCVNNPricer, TrainingConfig, build_gbm_cvnn_pricer_config
from spectralmc_b200.numerical import Precision
from tests.conftest import ROOT
from tests.helpers import expect_success, make_black_scholes_config, make_domain_bounds, make_simulation_params

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = [(n, t) for n in ("pricer", "deep") for t in ("float64", "float32")]
TOL = {"float64": 1e-12, "float32": 1e-5}
DEV = torch.device("cuda")


def nw(a, b) -> float:
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-300))


def build(layers, dtype):
    mods = []
    for layer in layers:
        if layer[0] == "linear":
            mods.append(cvnn.ComplexLinear(layer[1], layer[2], bias=bool(layer[3])))
        elif layer[0] == "modrelu":
            mods.append(cvnn.modReLU(layer[1]))
        else:
            mods.append(cvnn.zReLU())
    return cvnn.ComplexSequential(*mods).to(device=DEV, dtype=dtype)


def golden_net(name, tag):
    z = np.load(os.path.join(GOLDEN, f"cvnn_{name}_{tag}.npz"))
    layers = [tuple(layer) for layer in json.loads(str(z["layers"]))]
    dtype = torch.float64 if tag == "float64" else torch.float32
    net = build(layers, dtype)
    n = len(list(net.parameters()))
    with torch.no_grad():
        for i, p in enumerate(net.parameters()):
            p.copy_(torch.from_numpy(z[f"param0_{i}"]))
    batches = [tuple(torch.from_numpy(z[f"{k}_{s}"]).to(DEV) for k in ("real_in", "imag_in", "targets")) for s in range(3)]
    return z, net, n, batches


@pytest.mark.parametrize("name,tag", CASES)
def test_forward_and_gradients_match_the_reference(name, tag) -> None:
    z, net, n, batches = golden_net(name, tag)
    fused = cvnn.FusedCVNN(net, lr=float(z["lr"]))
    pr, pi = fused.forward(*batches[0][:2])
    assert nw(pr, z["pred_r"]) <= TOL[tag] and nw(pi, z["pred_i"]) <= TOL[tag]
    loss = fused.loss_backward(*batches[0])
    assert abs(float(loss) - z["losses"][0]) <= TOL[tag] * z["losses"][0]
    for i, p in enumerate(net.parameters()):
        assert p.grad is not None and nw(p.grad, z[f"grad0_{i}"]) <= TOL[tag], i
    # parameters untouched by forward / loss_backward
    for i, p in enumerate(net.parameters()):
        assert np.array_equal(p.detach().cpu().numpy(), z[f"param0_{i}"])


@pytest.mark.parametrize("name,tag", CASES)
def test_three_training_steps_match_the_reference(name, tag) -> None:
    z, net, n, batches = golden_net(name, tag)
    fused = cvnn.FusedCVNN(net, lr=float(z["lr"]))
    losses = [float(fused.train_step(*b)) for b in batches]
    for got, ref in zip(losses, z["losses"]):
        assert abs(got - ref) <= 10 * TOL[tag] * ref
    for i, p in enumerate(net.parameters()):
        assert nw(p, z[f"param3_{i}"]) <= 10 * TOL[tag], i
    assert int(fused.step.item()) == 3
    sd = fused.optimizer_state_dict()
    assert set(sd["state"]) == set(range(n)) and float(sd["state"][0]["step"]) == 3.0


SHAPES = [
    # (layers, rows): ragged widths, one row, several weight-gradient splits / column-sum slabs
    ([("linear", 6, 32, True), ("modrelu", 32), ("linear", 32, 128, True)], 1024),
    ([("linear", 3, 5, True), ("zrelu",), ("linear", 5, 2, False)], 1),
    ([("modrelu", 4), ("linear", 4, 70, True), ("zrelu",), ("modrelu", 70), ("linear", 70, 33, True), ("modrelu", 33)], 300),
    ([("linear", 6, 64, True), ("modrelu", 64), ("linear", 64, 64, True), ("modrelu", 64), ("linear", 64, 16, True)], 3000),
    ([("linear", 17, 1, True)], 513),
]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("layers,rows", SHAPES)
def test_fused_step_matches_torch_autograd_and_oracle(layers, rows, dtype) -> None:
    tol = 1e-11 if dtype == torch.float64 else 2e-5
    torch.manual_seed(rows)
    net = build(layers, dtype)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.add_(0.2 * torch.randn_like(p))
    n_in = layers[0][1]
    real = torch.randn(rows, n_in, dtype=dtype, device=DEV)
    imag = torch.randn(rows, n_in, dtype=dtype, device=DEV)
    ref_r, ref_i = net(real, imag)
    targets = torch.complex(torch.randn_like(ref_r), torch.randn_like(ref_r))
    ref_loss = torch.nn.functional.mse_loss(ref_r, targets.real) + torch.nn.functional.mse_loss(ref_i, targets.imag)
    ref_grads = torch.autograd.grad(ref_loss, list(net.parameters()))
    params0 = [p.detach().cpu().numpy().copy() for p in net.parameters()]

    fused = cvnn.FusedCVNN(net, lr=1e-2)
    pr, pi = fused.forward(real, imag)
    assert nw(pr, ref_r) <= tol and nw(pi, ref_i) <= tol
    loss = fused.loss_backward(real, imag, targets)
    assert abs(float(loss) - float(ref_loss.detach())) <= tol * float(ref_loss.detach())
    for p, g in zip(net.parameters(), ref_grads):
        assert nw(p.grad, g) <= tol * 5

    o_loss, o_grads, _ = ocvnn.loss_and_grads(layers, params0, real.cpu().numpy(), imag.cpu().numpy(), targets.cpu().numpy())
    assert abs(float(loss) - o_loss) <= tol * o_loss
    for p, g in zip(net.parameters(), o_grads):
        assert nw(p.grad, g) <= tol * 5

    # bit reproducibility of the whole backward (fixed-order reductions)
    first = fused.grads.clone()
    fused.grads.fill_(float("nan"))
    fused.loss_backward(real, imag, targets)
    assert torch.equal(first, fused.grads)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_adam_matches_torch_optimizer_over_many_steps(dtype) -> None:
    """Same gradients into smc_adam_step and torch.optim.Adam for 25 steps."""
    torch.manual_seed(3)
    n = 1000
    p_ref = torch.nn.Parameter(torch.randn(n, dtype=dtype, device=DEV))
    opt = torch.optim.Adam([p_ref], lr=3e-3)
    p = p_ref.detach().clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int64, device=DEV)
    hyper = _cabi.AdamArgs(3e-3, 0.9, 0.999, 1e-8)
    for _ in range(25):
        g = torch.randn(n, dtype=dtype, device=DEV) * torch.rand(1, device=DEV).to(dtype)
        p_ref.grad = g.clone()
        opt.step()
        _cabi.adam_step(p, g, m, v, step, hyper)
    assert int(step.item()) == 25
    assert nw(p, p_ref) <= (1e-13 if dtype == torch.float64 else 2e-6)
    assert nw(m, opt.state[p_ref]["exp_avg"]) <= (1e-13 if dtype == torch.float64 else 2e-6)
    assert nw(v, opt.state[p_ref]["exp_avg_sq"]) <= (1e-13 if dtype == torch.float64 else 2e-6)


def test_workspace_guard_bands_survive_a_training_step() -> None:
    layers = [("linear", 7, 19, True), ("modrelu", 19), ("linear", 19, 5, True)]
    net, n = _cabi.make_cvnn_net(layers, 7, torch.float32)
    rows, band = 333, 4096
    need = _cabi.cvnn_workspace_bytes(net, rows, True)
    raw = torch.full((need + 2 * band,), 0xA5, dtype=torch.uint8, device=DEV)
    ws = raw[band : band + need]
    assert ws.data_ptr() % 16 == 0
    bufs = [torch.full((n + 64,), 7.0, dtype=torch.float32, device=DEV) for _ in range(4)]
    params, grads, m, v = [b[32 : 32 + n] for b in bufs]
    params.normal_()
    m.zero_()
    v.zero_()
    real = torch.randn(rows, 7, device=DEV)
    targets = torch.randn(rows, 5, 2, device=DEV)
    loss = torch.zeros(3, dtype=torch.float64, device=DEV)
    step = torch.zeros(3, dtype=torch.int64, device=DEV)
    _cabi.cvnn_train_step(net, params, grads, m, v, step[1:2], _cabi.AdamArgs(1e-2, 0.9, 0.999, 1e-8), real, real.clone(), targets,
                          loss[1:2], ws)
    torch.cuda.synchronize()
    assert bool((raw[:band] == 0xA5).all()) and bool((raw[band + need :] == 0xA5).all())
    for b in bufs[1:]:  # grads, m, v: bands intact (params was re-drawn in place)
        assert bool((b[:32] == 7.0).all()) and bool((b[32 + n :] == 7.0).all())
    assert step.tolist() == [0, 1, 0] and loss[0] == 0 and loss[2] == 0 and float(loss[1]) > 0
    with pytest.raises(_cabi.SmcError) as err:
        _cabi.cvnn_train_step(net, params, grads, m, v, step[1:2], _cabi.AdamArgs(1e-2, 0.9, 0.999, 1e-8), real, real.clone(), targets,
                              loss[1:2], ws[: need - 512])
    assert err.value.code == 3


# ----------------------------------------------------------------------------- trainer wiring
def _pricer(precision=Precision.float32, *, seed=42, N=16, B=2**10, T=2, **kw):
    sp = make_simulation_params(timesteps=T, network_size=N, batches_per_mc_run=B, threads_per_block=256, mc_seed=seed,
                                buffer_size=1, dtype=precision)
    cfg = make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
    return GbmCVNNPricer(cfg, make_domain_bounds(), cvnn.make_cvnn(6, N, seed=seed, dtype=precision.to_torch()), **kw)


def _params(p):
    return [q.detach().clone() for q in p._cvnn.parameters()]


@pytest.mark.parametrize("precision", [Precision.float32, Precision.float64])
def test_graph_replay_is_bit_identical_to_eager_fused_steps(precision) -> None:
    a, b = _pricer(precision, cuda_graph=True), _pricer(precision, cuda_graph=False)
    assert a._use_fused and b._use_fused
    la = expect_success(a.train(TrainingConfig(num_batches=4, batch_size=16))).losses
    lb = expect_success(b.train(TrainingConfig(num_batches=4, batch_size=16))).losses
    assert la == lb and len(a._graphs) == 1 and not b._graphs
    assert all(torch.equal(x, y) for x, y in zip(_params(a), _params(b)))
    # a second batch size captures a second graph; the first keeps working
    expect_success(a.train(TrainingConfig(num_batches=1, batch_size=8)))
    expect_success(a.train(TrainingConfig(num_batches=1, batch_size=16)))
    assert sorted(a._graphs) == [8, 16] and int(a._fused.step.item()) == 6


@pytest.mark.parametrize("precision", [Precision.float32, Precision.float64])
def test_fused_training_tracks_the_torch_route(precision) -> None:
    """Same targets, same initial weights: the C-ABI step and the torch op-by-op step
    (GbmCVNNPricer._torch_step, reference gbm_trainer.py:819-835) stay together."""
    a, b = _pricer(precision), _pricer(precision, fused_step=False)
    ra = expect_success(a.train(TrainingConfig(num_batches=5, batch_size=32)))
    rb = expect_success(b.train(TrainingConfig(num_batches=5, batch_size=32)))
    la, lb = ra.losses, rb.losses
    assert abs(ra.final_grad_norm - rb.final_grad_norm) <= (1e-9 if precision is Precision.float64 else 2e-3) * rb.final_grad_norm
    tol = 1e-9 if precision is Precision.float64 else 2e-3
    assert max(abs(x - y) / abs(y) for x, y in zip(la, lb)) <= tol
    assert max(nw(x, y) for x, y in zip(_params(a), _params(b))) <= tol
    # snapshots of the two routes are interchangeable (torch Adam state_dict layout)
    sa, sb = expect_success(a.snapshot()), expect_success(b.snapshot())
    assert set(sa.optimizer_state["state"]) == set(sb.optimizer_state["state"])
    import copy

    c = expect_success(GbmCVNNPricer.create(sb.model_copy(update={"cvnn": copy.deepcopy(b._cvnn)})))  # torch-route snapshot -> fused trainer
    assert c._use_fused
    lc = expect_success(c.train(TrainingConfig(num_batches=1, batch_size=32))).losses
    ld = expect_success(b.train(TrainingConfig(num_batches=1, batch_size=32))).losses
    assert int(c._fused.step.item()) == 6 and abs(lc[0] - ld[0]) / abs(ld[0]) <= tol


def test_learning_rate_change_recaptures_the_graph() -> None:
    a, b = _pricer(), _pricer(cuda_graph=False)
    for lr in (1e-2, 1e-3):
        expect_success(a.train(TrainingConfig(num_batches=2, batch_size=8, learning_rate=lr)))
        expect_success(b.train(TrainingConfig(num_batches=2, batch_size=8, learning_rate=lr)))
    assert all(torch.equal(x, y) for x, y in zip(_params(a), _params(b)))


def test_predict_price_is_the_dc_bin_over_n() -> None:
    """mean_n ifft(S)[n] == S[0] / N (reference gbm_trainer.py:1729-1730)."""
    p = _pricer(Precision.float64)
    inputs = [BlackScholes.Inputs(X0=100, K=100, T=1.0, r=0.05, d=0.0, v=0.2), BlackScholes.Inputs(X0=50, K=60, T=0.5, r=0.0, d=0.0, v=0.4)]
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        fused = expect_success(p.predict_price(inputs))
        generic = expect_success(_pricer(Precision.float64, fused_step=False).predict_price(inputs))
    for x, y in zip(fused, generic):
        assert abs(x.put_price - y.put_price) <= 1e-12 * abs(y.put_price) and abs(x.call_price - y.call_price) <= 1e-12 * abs(y.call_price)


def test_unsupported_networks_take_the_torch_route() -> None:
    class Odd(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.inner = cvnn.make_cvnn(6, 16, seed=1)

        def forward(self, real, imag):
            return self.inner(real, imag)

    sp = make_simulation_params(timesteps=1, network_size=16, batches_per_mc_run=256, mc_seed=5, dtype=Precision.float32)
    cfg = make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
    p = GbmCVNNPricer(cfg, make_domain_bounds(), Odd())
    assert not p._use_fused
    expect_success(p.train(TrainingConfig(num_batches=1, batch_size=8)))
    with pytest.raises(ValueError):
        GbmCVNNPricer(cfg, make_domain_bounds(), Odd(), fused_step=True)


def test_wide_network_and_large_batch() -> None:
    """Sizes beyond the reference's tests: 20 000 rows (157 column-sum slabs, 256 weight-gradient splits
    capped), hidden width 256, 128 outputs."""
    layers = [("linear", 6, 256, True), ("modrelu", 256), ("linear", 256, 128, True)]
    torch.manual_seed(5)
    net = build(layers, torch.float32)
    rows = 20000
    real, imag = torch.randn(rows, 6, device=DEV), torch.randn(rows, 6, device=DEV)
    ref_r, ref_i = net(real, imag)
    targets = torch.complex(torch.randn_like(ref_r), torch.randn_like(ref_r))
    ref_loss = torch.nn.functional.mse_loss(ref_r, targets.real) + torch.nn.functional.mse_loss(ref_i, targets.imag)
    ref_grads = torch.autograd.grad(ref_loss, list(net.parameters()))
    fused = cvnn.FusedCVNN(net)
    loss = fused.loss_backward(real, imag, targets)
    assert abs(float(loss) - float(ref_loss.detach())) <= 1e-5 * float(ref_loss.detach())
    for p, g in zip(net.parameters(), ref_grads):
        assert nw(p.grad, g) <= 1e-4


def test_factory_networks_with_batch_norm_and_residuals_train_on_the_torch_route() -> None:
    """A network using every block of the reference's factory (batch norms, residuals) is outside the
    fused step: the trainer takes the torch route for it — autograd + smc_adam_step, captured once and
    replayed as one CUDA graph — deterministically, and in step with the eager torch route."""
    from spectralmc_b200 import cvnn_factory as f

    def act(kind):
        return f.ActivationCfg(kind=kind)

    def trainer(**kw):
        layers = [
            f.LinearCfg(width=f.ExplicitWidth(value=24), activation=act(f.ActivationKind.MOD_RELU)),
            f.CovBNCfg(),
            f.ResidualCfg(body=f.SequentialCfg(layers=[f.LinearCfg(activation=act(f.ActivationKind.Z_RELU)), f.NaiveBNCfg()]),
                          activation=act(f.ActivationKind.MOD_RELU)),
        ]
        cfg = expect_success(f.build_cvnn_config(dtype=Precision.float32, layers=layers, seed=21))
        net = expect_success(f.build_model(n_inputs=6, n_outputs=16, cfg=cfg)).to("cuda", torch.float32)
        sp = make_simulation_params(timesteps=2, network_size=16, batches_per_mc_run=512, mc_seed=21, dtype=Precision.float32)
        bs = make_black_scholes_config(sim_params=sp, path_scheme=PathScheme.LOG_EULER, normalization=ForwardNormalization.RAW)
        return expect_success(GbmCVNNPricer.create(expect_success(build_gbm_cvnn_pricer_config(cfg=bs, domain_bounds=make_domain_bounds(), cvnn=net)), **kw))

    a, b = trainer(), trainer()
    assert not a._use_fused
    ra = expect_success(a.train(TrainingConfig(num_batches=3, batch_size=32)))
    rb = expect_success(b.train(TrainingConfig(num_batches=3, batch_size=32)))
    assert ra.losses == rb.losses and all(np.isfinite(ra.losses))
    assert all(torch.equal(x, y) for x, y in zip(_params(a), _params(b)))
    assert len(a._torch_graphs) == 1 and a._flat_adam is not None and int(a._flat_adam.step.item()) == 3  # one captured step, replayed
    # ... and the captured step follows the eager torch step (torch.optim.Adam, op by op) of the reference
    c = trainer(cuda_graph=False)
    rc = expect_success(c.train(TrainingConfig(num_batches=3, batch_size=32)))
    assert not c._torch_graphs and max(abs(x - y) / abs(y) for x, y in zip(ra.losses, rc.losses)) <= 1e-4
    # Adam moves a parameter whose gradient is rounding noise by +-lr per step in either direction: bound the
    # divergence of the two float32 runs by the three steps taken, and require agreement where gradients are real
    assert max(float((x - y).abs().max()) for x, y in zip(_params(a), _params(c))) <= 3 * 1e-2 * 1.01
    assert nw(_params(a)[0], _params(c)[0]) <= 1e-3  # first layer's weights: large, well-defined gradients
    sa, sc = expect_success(a.snapshot()), expect_success(c.snapshot())
    assert set(sa.optimizer_state["state"]) == set(sc.optimizer_state["state"]) and float(sa.optimizer_state["state"][0]["step"]) == 3.0
    assert float(a._cvnn.state_dict()["layers.0.layers.1.running_C_rr"].sub(0.5).abs().max()) > 0  # statistics were tracked
