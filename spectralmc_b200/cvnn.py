"""Complex-valued network blocks consuming the hot path's ``[C, N]`` targets, and their fused step.

Two things live here:

* the layer classes of the reference's ``spectralmc.cvnn`` that its pricer networks are made of —
  ``ComplexLinear`` (cvnn.py:65-146: four real matmuls, Xavier-uniform weights, zero biases),
  ``zReLU`` (:149-162), ``modReLU`` (:168-210, epsilon 1e-9 under the root) and
  ``ComplexSequential`` (:439-452) — same constructor arguments, parameter names and forward
  semantics, models mapping ``(real, imag) -> (real, imag)``.  Their ``forward`` is plain torch:
  it is the generic route for networks the fused step does not cover.  The remaining blocks the
  reference's factory can emit — ``NaiveComplexBatchNorm`` (:213-274), ``CovarianceComplexBatchNorm``
  (:277-433) and ``ComplexResidual`` (:454-493) — are provided with the same parameter / buffer
  names (state dicts interchange with the reference's) and stay on the torch route.
* ``FusedCVNN``: the same network driven through the C ABI (``smc_cvnn_forward``,
  ``smc_cvnn_train_step``): one complex GEMM per ``ComplexLinear`` with bias and activation in
  the epilogue, hand-derived backward, MSE loss and Adam (``GbmCVNNPricer._torch_step``,
  gbm_trainer.py:819-835) as a fixed launch sequence with no host round trip, so a whole
  training step replays as ONE CUDA graph.
"""

from __future__ import annotations

from functools import reduce

import torch
from torch import nn

from spectralmc_b200 import _cabi


class ComplexLinear(nn.Module):
    """``W z + b`` with ``W = A + iB`` held as two real matrices (reference cvnn.py:65-146)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True) -> None:
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.real_weight = nn.Parameter(torch.empty(out_features, in_features))
        self.imag_weight = nn.Parameter(torch.empty(out_features, in_features))
        if bias:
            self.real_bias = nn.Parameter(torch.empty(out_features))
            self.imag_bias = nn.Parameter(torch.empty(out_features))
        else:
            self.real_bias = None
            self.imag_bias = None
        self.reset_parameters()

    def reset_parameters(self) -> None:
        nn.init.xavier_uniform_(self.real_weight)
        nn.init.xavier_uniform_(self.imag_weight)
        if self.real_bias is not None:
            nn.init.zeros_(self.real_bias)
            nn.init.zeros_(self.imag_bias)

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        out_r = real @ self.real_weight.T - imag @ self.imag_weight.T
        out_i = real @ self.imag_weight.T + imag @ self.real_weight.T
        if self.real_bias is not None:
            out_r, out_i = out_r + self.real_bias, out_i + self.imag_bias
        return out_r, out_i


class zReLU(nn.Module):
    """Pass ``z`` where ``Re z >= 0`` and ``Im z >= 0``, else 0 (reference cvnn.py:149-162)."""

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        mask = (real >= 0) & (imag >= 0)
        return real * mask, imag * mask


class modReLU(nn.Module):
    """``z -> relu(|z| + b) * z / |z|`` with ``|z| = sqrt(x^2 + y^2 + 1e-9)`` (reference cvnn.py:168-210)."""

    def __init__(self, num_features: int) -> None:
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(num_features))

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        mod = torch.sqrt(real * real + imag * imag + 1e-9)
        scale = torch.relu(mod + self.bias.unsqueeze(0)) / mod
        return scale * real, scale * imag


class ComplexSequential(nn.Module):
    """Left fold of complex blocks (reference cvnn.py:439-452)."""

    def __init__(self, *layers: nn.Module) -> None:
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        return reduce(lambda state, layer: layer(*state), self.layers, (real, imag))


class NaiveComplexBatchNorm(nn.Module):
    """``BatchNorm1d`` applied to the real and to the imaginary plane independently (reference
    cvnn.py:213-274; sub-module names ``bn_real`` / ``bn_imag`` as there)."""

    def __init__(self, num_features: int, *, eps: float = 1e-5, momentum: float = 0.1, affine: bool = True,
                 track_running_stats: bool = True) -> None:
        super().__init__()
        kw = dict(eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats)
        self.bn_real = nn.BatchNorm1d(num_features, **kw)
        self.bn_imag = nn.BatchNorm1d(num_features, **kw)

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        return self.bn_real(real), self.bn_imag(imag)


class CovarianceComplexBatchNorm(nn.Module):
    """Whitening batch norm of Trabelsi et al. (reference cvnn.py:277-433): each feature's (re, im)
    pair is centred and multiplied by ``V^{-1/2}`` of its 2x2 covariance ``V`` (+ eps on the
    diagonal), then mapped by a learnable symmetric ``Gamma`` and shift ``beta``.

    The reference forms ``V^{-1/2}`` with a batched ``eigh``; this class uses the closed form for a
    symmetric positive-definite 2x2 matrix, ``V^{-1/2} = [[c + s, -b], [-b, a + s]] / (s t)`` with
    ``s = sqrt(det V)``, ``t = sqrt(a + c + 2 s)`` — the same matrix (eigenvalues are >= eps, so the
    reference's clamp never acts), without a LAPACK-style call per step.  Buffers and parameters keep
    the reference's names, so state dicts interchange.
    """

    def __init__(self, num_features: int, *, eps: float = 1e-5, momentum: float = 0.1, affine: bool = True,
                 track_running_stats: bool = True) -> None:
        super().__init__()
        self.eps, self.momentum, self.affine, self.track_running_stats = eps, momentum, affine, track_running_stats
        for name, fill in (("running_mean_real", 0.0), ("running_mean_imag", 0.0), ("running_C_rr", 0.5), ("running_C_ri", 0.0),
                           ("running_C_ii", 0.5)):
            self.register_buffer(name, torch.full((num_features,), fill))
        names = ("beta_real", "beta_imag", "gamma_rr", "gamma_ri", "gamma_ii")
        init = (0.0, 0.0, 1.0, 0.0, 1.0)
        for name, fill in zip(names, init):
            if affine:
                setattr(self, name, nn.Parameter(torch.full((num_features,), fill)))
            else:
                self.register_parameter(name, None)

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        if self.training or not self.track_running_stats:
            mu_r, mu_i = real.mean(dim=0), imag.mean(dim=0)
            dr, di = real - mu_r, imag - mu_i
            a, b, c = (dr * dr).mean(dim=0), (dr * di).mean(dim=0), (di * di).mean(dim=0)
            if self.track_running_stats:
                with torch.no_grad():
                    for buf, val in ((self.running_mean_real, mu_r), (self.running_mean_imag, mu_i), (self.running_C_rr, a),
                                     (self.running_C_ri, b), (self.running_C_ii, c)):
                        buf.mul_(1 - self.momentum).add_(val * self.momentum)
        else:
            dr, di = real - self.running_mean_real, imag - self.running_mean_imag
            a, b, c = self.running_C_rr, self.running_C_ri, self.running_C_ii
        a, c = a + self.eps, c + self.eps
        s = torch.sqrt(a * c - b * b)
        inv = 1.0 / (s * torch.sqrt(a + c + 2.0 * s))
        w_rr, w_ri, w_ii = (c + s) * inv, -b * inv, (a + s) * inv
        white_r, white_i = w_rr * dr + w_ri * di, w_ri * dr + w_ii * di
        if not self.affine:
            return white_r, white_i
        return (self.gamma_rr * white_r + self.gamma_ri * white_i + self.beta_real,
                self.gamma_ri * white_r + self.gamma_ii * white_i + self.beta_imag)


class ComplexResidual(nn.Module):
    """``post_act(proj(x) + body(x))`` with optional projection and activation (reference cvnn.py:454-493)."""

    def __init__(self, body: nn.Module, proj: nn.Module | None = None, post_act: nn.Module | None = None) -> None:
        super().__init__()
        self.body, self.proj, self.post_act = body, proj, post_act

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        body_r, body_i = self.body(real, imag)
        skip_r, skip_i = (real, imag) if self.proj is None else self.proj(real, imag)
        out = (body_r + skip_r, body_i + skip_i)
        return out if self.post_act is None else self.post_act(*out)


def make_cvnn(n_inputs: int, n_outputs: int, *, hidden_width: int = 32, seed: int = 0,
              dtype: torch.dtype = torch.float32, device: torch.device | str = "cuda") -> ComplexSequential:
    """6 -> hidden (modReLU) -> N, the shape of the reference's test network
    (tests/helpers/factories.py:69-105), nested as the factory nests it (cvnn_factory.py:185-190);
    weights seeded inside a forked RNG as ``build_model`` does (cvnn_factory.py:343-368)."""
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        net = ComplexSequential(ComplexSequential(ComplexLinear(n_inputs, hidden_width), modReLU(hidden_width)),
                                ComplexLinear(hidden_width, n_outputs))
    return net.to(device=device, dtype=dtype)


# ----------------------------------------------------------------------------- fused step
def describe(module: nn.Module) -> list[tuple] | None:
    """Flatten ``module`` into the C ABI's layer list, or ``None`` if it holds an unsupported block."""
    if isinstance(module, ComplexSequential):
        out: list[tuple] = []
        for child in module.layers:
            sub = describe(child)
            if sub is None:
                return None
            out += sub
        return out
    if isinstance(module, ComplexLinear):
        return [("linear", module.in_features, module.out_features, module.real_bias is not None)]
    if isinstance(module, modReLU):
        return [("modrelu", int(module.bias.numel()))]
    if isinstance(module, zReLU):
        return [("zrelu",)]
    return None


class FlatAdam:
    """``torch.optim.Adam`` (defaults: no weight decay, no amsgrad) over ONE flat buffer, stepped by
    ``smc_adam_step``: the step counter lives on the device and the bias corrections are formed in float64
    inside the kernel, so a step is capturable in a CUDA graph AND follows the eager optimiser's arithmetic
    (torch's own capturable Adam keeps ``step`` in float32 and drifts by ~1e-7 in early steps).

    Construction re-points every parameter at a view of the flat parameter buffer (values unchanged) and
    every ``.grad`` at a view of the flat gradient buffer; moving the module afterwards (``.to()``) detaches
    it from the buffers and is not supported.  State is exchanged in ``torch.optim.Adam.state_dict()`` layout.
    """

    def __init__(self, params: list[nn.Parameter], *, lr: float = 1e-2, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8) -> None:
        if not params:
            raise ValueError("no parameters")
        self._params = params  # ALL parameters, so that state-dict indices line up with torch.optim.Adam(params)
        live = [p for p in params if p.requires_grad]
        if not live:
            raise ValueError("no trainable parameters")
        self.dtype, self.device = live[0].dtype, live[0].device
        if self.device.type != "cuda":
            raise ValueError("FlatAdam needs CUDA parameters (spectralmc_b200 has no CPU path)")
        if any(p.dtype != self.dtype or p.device != self.device for p in live):
            raise ValueError("parameters must share one device and dtype")
        n = sum(p.numel() for p in live)
        self.params = torch.empty(n, dtype=self.dtype, device=self.device)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.step = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.hyper = _cabi.AdamArgs(lr, betas[0], betas[1], eps)
        # frozen parameters (requires_grad=False) are neither flattened nor updated, as torch.optim.Adam skips
        # them; their slot stays None.  Trainable parameters are updated DENSELY every step: one that receives no
        # gradient in some step sees a zero gradient (its moments decay and residual momentum still moves it),
        # whereas eager Adam with zero_grad(set_to_none=True) skips it for that step.  The two agree whenever
        # every trainable parameter gets a gradient in every step — true for every network cvnn_factory builds.
        self._slices: list[tuple[int, int] | None] = []
        offset = 0
        with torch.no_grad():
            for p in params:
                if not p.requires_grad:
                    self._slices.append(None)
                    continue
                k = p.numel()
                self.params[offset : offset + k].copy_(p.detach().reshape(-1))
                p.data = self.params[offset : offset + k].view(p.shape)
                p.grad = self.grads[offset : offset + k].view(p.shape)
                self._slices.append((offset, k))
                offset += k

    def apply(self) -> None:
        """One Adam update of the whole buffer from ``self.grads``; no host synchronisation."""
        _cabi.adam_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.step, self.hyper)

    def state_dict(self) -> dict:
        step = float(self.step.item())
        state = {}
        if step > 0:
            for i, (p, sl) in enumerate(zip(self._params, self._slices)):
                if sl is None:
                    continue
                o, k = sl
                state[i] = {"step": torch.tensor(step), "exp_avg": self.exp_avg[o : o + k].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o : o + k].view(p.shape).clone()}
        group = {"lr": self.hyper.lr, "betas": (self.hyper.beta1, self.hyper.beta2), "eps": self.hyper.eps, "weight_decay": 0,
                 "amsgrad": False, "params": list(range(len(self._slices)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        group = sd["param_groups"][0]
        self.hyper = _cabi.AdamArgs(float(group["lr"]), group["betas"][0], group["betas"][1], group["eps"])
        steps = {float(s["step"]) for s in sd["state"].values()}
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ")
        self.step.fill_(int(steps.pop()) if steps else 0)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, sl in enumerate(self._slices):
            if sl is not None and i in sd["state"]:
                o, k = sl
                self.exp_avg[o : o + k].copy_(sd["state"][i]["exp_avg"].reshape(-1))
                self.exp_avg_sq[o : o + k].copy_(sd["state"][i]["exp_avg_sq"].reshape(-1))


class FusedCVNN:
    """A supported network + Adam state in flat device buffers (``FlatAdam``), stepped through the C ABI.

    The module keeps working as a torch module (``state_dict``, torch ``forward``, gradient inspection): its
    parameters and gradients are views of the flat buffers.
    """

    def __init__(self, net: nn.Module, *, lr: float = 1e-2, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8) -> None:
        layers = describe(net)
        if not layers:
            raise ValueError("network is not a ComplexSequential of ComplexLinear / modReLU / zReLU")
        params = list(net.parameters())
        if not params:
            raise ValueError("network has no parameters")
        if params[0].device.type != "cuda":
            raise ValueError("FusedCVNN needs the network on a CUDA device (spectralmc_b200 has no CPU path)")
        if layers[0][0] == "zrelu":
            raise ValueError("cannot infer the input width of a network that starts with zReLU")
        self.n_inputs = layers[0][1]
        self.net, self.layers = net, layers
        self.adam = FlatAdam(params, lr=lr, betas=betas, eps=eps)
        self.dtype, self.device = self.adam.dtype, self.adam.device
        self.desc, n = _cabi.make_cvnn_net(layers, self.n_inputs, self.dtype)
        if n != self.adam.params.numel():
            raise AssertionError("descriptor / parameter count mismatch")
        self.n_outputs = _cabi.cvnn_output_width(self.desc)
        self._workspaces: dict[tuple[int, bool], torch.Tensor] = {}

    # the flat buffers, under the names the C ABI uses
    params = property(lambda self: self.adam.params)
    grads = property(lambda self: self.adam.grads)
    exp_avg = property(lambda self: self.adam.exp_avg)
    exp_avg_sq = property(lambda self: self.adam.exp_avg_sq)
    step = property(lambda self: self.adam.step)

    def complex_macs(self, rows: int) -> int:
        """Complex multiply-adds of one forward pass over ``rows`` inputs (the step does three GEMMs per layer)."""
        return rows * sum(layer[1] * layer[2] for layer in self.layers if layer[0] == "linear")

    @property
    def hyper(self) -> "_cabi.AdamArgs":
        return self.adam.hyper

    @hyper.setter
    def hyper(self, value: "_cabi.AdamArgs") -> None:
        self.adam.hyper = value

    def _workspace(self, rows: int, training: bool) -> torch.Tensor:
        key = (rows, training)
        ws = self._workspaces.get(key)
        if ws is None:
            ws = torch.empty(_cabi.cvnn_workspace_bytes(self.desc, rows, training), dtype=torch.uint8, device=self.device)
            self._workspaces[key] = ws
        return ws

    def _check_inputs(self, real: torch.Tensor, imag: torch.Tensor) -> int:
        if real.shape != imag.shape or real.dim() != 2 or real.shape[1] != self.n_inputs:
            raise ValueError(f"inputs must be two [rows, {self.n_inputs}] tensors")
        if real.dtype != self.dtype or imag.dtype != self.dtype:
            raise TypeError(f"inputs must be {self.dtype}")
        return real.shape[0]

    def forward(self, real: torch.Tensor, imag: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        rows = self._check_inputs(real, imag)
        out_r = torch.empty((rows, self.n_outputs), dtype=self.dtype, device=self.device)
        out_i = torch.empty_like(out_r)
        if rows:
            _cabi.cvnn_forward(self.desc, self.params, real.contiguous(), imag.contiguous(), out_r, out_i, self._workspace(rows, False))
        return out_r, out_i

    __call__ = forward

    def _check_targets(self, rows: int, targets: torch.Tensor) -> torch.Tensor:
        if targets.shape != (rows, self.n_outputs) or targets.dtype != _cabi.complex_dtype(self.dtype):
            raise ValueError(f"targets must be [{rows}, {self.n_outputs}] {_cabi.complex_dtype(self.dtype)}")
        return torch.view_as_real(targets.contiguous())

    def loss_backward(self, real: torch.Tensor, imag: torch.Tensor, targets: torch.Tensor, loss: torch.Tensor | None = None) -> torch.Tensor:
        """Forward, ``mse(real) + mse(imag)`` and every gradient (``p.grad`` views); returns the device loss."""
        rows = self._check_inputs(real, imag)
        loss = torch.empty(1, dtype=torch.float64, device=self.device) if loss is None else loss
        _cabi.cvnn_loss_backward(self.desc, self.params, real.contiguous(), imag.contiguous(), self._check_targets(rows, targets),
                                 self.grads, loss, self._workspace(rows, True))
        return loss

    def train_step(self, real: torch.Tensor, imag: torch.Tensor, targets: torch.Tensor, loss: torch.Tensor | None = None) -> torch.Tensor:
        """One ``_torch_step`` (gbm_trainer.py:819-835): loss, backward, Adam.  No host synchronisation."""
        rows = self._check_inputs(real, imag)
        loss = torch.empty(1, dtype=torch.float64, device=self.device) if loss is None else loss
        _cabi.cvnn_train_step(self.desc, self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.step, self.hyper,
                              real.contiguous(), imag.contiguous(), self._check_targets(rows, targets), loss,
                              self._workspace(rows, True))
        return loss

    def warm_up(self, rows: int) -> None:
        """Launch every kernel of a step once WITHOUT touching parameters or optimiser state, so
        module loading cannot fall inside a graph capture."""
        real = torch.zeros((rows, self.n_inputs), dtype=self.dtype, device=self.device)
        targets = torch.zeros((rows, self.n_outputs), dtype=_cabi.complex_dtype(self.dtype), device=self.device)
        scratch = [torch.zeros(1, dtype=self.dtype, device=self.device) for _ in range(4)]
        grads = torch.empty_like(self.grads)
        loss = torch.empty(1, dtype=torch.float64, device=self.device)
        _cabi.cvnn_loss_backward(self.desc, self.params, real, real, torch.view_as_real(targets), grads, loss, self._workspace(rows, True))
        _cabi.adam_step(scratch[0], scratch[1], scratch[2], scratch[3], torch.zeros(1, dtype=torch.int64, device=self.device), self.hyper)

    # ---- optimiser state in torch.optim.Adam's state_dict layout (snapshot interchange) -----
    def optimizer_state_dict(self) -> dict:
        return self.adam.state_dict()

    def load_optimizer_state_dict(self, sd: dict) -> None:
        self.adam.load_state_dict(sd)
