"""CPU oracle for the CVNN training step that consumes the CF targets.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): a NumPy restatement, with the backward
pass written out by hand, of

* ``ComplexLinear.forward``  /root/reference/src/spectralmc/cvnn.py:120-146
      out_r = x_r A^T - x_i B^T + b_r ;  out_i = x_r B^T + x_i A^T + b_i      (W = A + iB)
* ``zReLU.forward``          cvnn.py:160-162   (pass where Re >= 0 and Im >= 0)
* ``modReLU.forward``        cvnn.py:193-210   (|z| = sqrt(x^2 + y^2 + 1e-9); relu(|z| + b) / |z| * z)
* ``ComplexSequential``      cvnn.py:439-452   (left fold over layers; nested containers flatten)
* the trainer's loss and step ``GbmCVNNPricer._torch_step`` gbm_trainer.py:819-835
      loss = mse(pred_r, Re t) + mse(pred_i, Im t); zero_grad; backward; Adam step
* ``torch.optim.Adam`` with its defaults (betas 0.9/0.999, eps 1e-8, no weight decay, no
  amsgrad), which is what the trainer constructs (gbm_trainer.py:1513).

Pinned against outputs of the reference's own ``spectralmc.cvnn`` classes run under CPU torch
(``tests/golden/make_golden_cvnn.py`` -> ``tests/golden/cvnn_*.npz``), float64 to ~1e-13 and
float32 to ~1e-5 norm-wise (summation order inside the matmuls is the library's).

A network is described by ``layers``: a list of ``("linear", in_features, out_features, bias)``,
``("modrelu", features)`` or ``("zrelu",)`` and ``params``: the arrays of
``module.parameters()`` in order (linear: real_weight, imag_weight[, real_bias, imag_bias];
modrelu: bias).
"""

from __future__ import annotations

import numpy as np

MODRELU_EPS = 1e-9  # cvnn.py:205


def n_layer_params(layer) -> int:
    if layer[0] == "linear":
        return 4 if layer[3] else 2
    return 1 if layer[0] == "modrelu" else 0


def forward(layers, params, xr, xi, *, keep=False):
    """Left fold of the layers over (xr, xi).  With ``keep`` also returns what backward needs."""
    saved = []
    it = iter(params)
    for layer in layers:
        kind = layer[0]
        if kind == "linear":
            a, b = next(it), next(it)
            br, bi = (next(it), next(it)) if layer[3] else (None, None)
            out_r = xr @ a.T - xi @ b.T  # cvnn.py:137
            out_i = xr @ b.T + xi @ a.T  # cvnn.py:138
            if br is not None:
                out_r, out_i = out_r + br, out_i + bi
            saved.append((xr, xi, a, b))
            xr, xi = out_r, out_i
        elif kind == "modrelu":
            bias = next(it)
            mag = np.sqrt(xr * xr + xi * xi + xr.dtype.type(MODRELU_EPS))  # cvnn.py:205
            thr = np.maximum(mag + bias[None, :], 0)  # cvnn.py:206
            scale = thr / mag
            saved.append((xr, xi, mag, thr, bias))
            xr, xi = scale * xr, scale * xi
        elif kind == "zrelu":
            mask = (xr >= 0) & (xi >= 0)  # cvnn.py:161
            saved.append((mask,))
            xr, xi = xr * mask, xi * mask
        else:
            raise ValueError(kind)
    return (xr, xi, saved) if keep else (xr, xi)


def loss_and_grads(layers, params, xr, xi, target):
    """MSE(real) + MSE(imag) (gbm_trainer.py:828-830) and d loss / d params, in params order."""
    pr, pi, saved = forward(layers, params, xr, xi, keep=True)
    tr, ti = target.real.astype(pr.dtype), target.imag.astype(pr.dtype)
    dr, di = pr - tr, pi - ti
    count = dr.size
    loss = (dr * dr).sum(dtype=np.float64) / count + (di * di).sum(dtype=np.float64) / count
    gr, gi = 2.0 * dr / count, 2.0 * di / count
    grads_rev = []
    for layer, s in zip(reversed(layers), reversed(saved)):
        kind = layer[0]
        if kind == "linear":
            in_r, in_i, a, b = s
            g_a = gr.T @ in_r + gi.T @ in_i
            g_b = -gr.T @ in_i + gi.T @ in_r
            these = [g_a, g_b]
            if layer[3]:
                these += [gr.sum(axis=0), gi.sum(axis=0)]
            grads_rev.append(these)
            gr, gi = gr @ a + gi @ b, -gr @ b + gi @ a
        elif kind == "modrelu":
            zr, zi, mag, thr, bias = s
            on = (mag + bias[None, :]) > 0
            dot = gr * zr + gi * zi
            scale = thr / mag
            # d scale / d z = on * (-bias) * z / mag^3 ; d scale / d bias = on / mag
            corr = np.where(on, -bias[None, :] * dot / (mag * mag * mag), 0)
            grads_rev.append([np.where(on, dot / mag, 0).sum(axis=0)])
            gr, gi = gr * scale + corr * zr, gi * scale + corr * zi
        else:
            (mask,) = s
            grads_rev.append([])
            gr, gi = gr * mask, gi * mask
    grads = [g for these in reversed(grads_rev) for g in these]
    return float(loss), [g.astype(p.dtype) for g, p in zip(grads, params)], (pr, pi)


def adam_step(params, grads, exp_avg, exp_avg_sq, step, *, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update; ``step`` is the 1-based step being taken."""
    bc1 = 1.0 - beta1**step
    bc2_sqrt = np.sqrt(1.0 - beta2**step)
    out = []
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        t = p.dtype.type
        m += (g - m) * t(1.0 - beta1)  # exp_avg.lerp_(grad, 1 - beta1)
        v *= t(beta2)
        v += t(1.0 - beta2) * g * g  # addcmul_
        denom = np.sqrt(v) / t(bc2_sqrt) + t(eps)
        out.append((p - t(lr / bc1) * (m / denom)).astype(p.dtype))
    return out


def train_steps(layers, params, xr, xi, targets, *, lr, steps):
    """``steps`` trainer steps on the same batch list; returns (losses, params)."""
    params = [p.copy() for p in params]
    m = [np.zeros_like(p) for p in params]
    v = [np.zeros_like(p) for p in params]
    losses = []
    for s in range(steps):
        loss, grads, _ = loss_and_grads(layers, params, xr[s], xi[s], targets[s])
        params = adam_step(params, grads, m, v, s + 1, lr=lr)
        losses.append(loss)
    return losses, params
