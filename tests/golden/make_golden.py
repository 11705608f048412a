#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE in the build container.

Usage (build container only; needs /root/reference):
    python tests/golden/make_golden.py

What runs: the reference's own ``spectralmc.gbm.BlackScholes._simulate / price /
get_host_price`` and the trainer's ``cp.mean(cp.fft.fft(mat, axis=1), axis=0)``
(gbm_trainer.py:814-817), imported unmodified from /root/reference/src.  The container
has no GPU and no CuPy, so the harness supplies (and only supplies) stand-ins for the
third-party layers underneath the reference code:

* ``cupy``  -> a NumPy-backed module exposing exactly the calls the path makes
  (``linspace/exp/mean/expand_dims/asarray/maximum/fft.fft``, ``random.default_rng``,
  ``cuda.Stream/Event``).  The normal matrices it hands out are recorded in the
  fixture — they are the "reference's own normal draws" injected into both paths.
* ``numba.cuda`` -> Numba's own CUDA simulator (``NUMBA_ENABLE_CUDASIM=1``), which
  executes the reference kernel body ``SimulateBlackScholes`` verbatim.
* ``spectralmc.effects`` package ``__init__`` is bypassed (it drags in the S3 store,
  TensorBoard and a CUDA assert); the submodules the path needs are loaded as-is.

float32 caveat: compiled for a GPU, the kernel widens each float32
element to float64 (its scalar arguments are float64) and narrows on store; the
simulator, under NumPy-2 weak-scalar promotion, would instead stay in float32.  For
float32 cases the harness therefore runs the same kernel body on a float64 copy of the
float32 matrix and narrows the result — exactly the compiled data flow.  The PTX opcode
histogram that justifies this is printed by ``--ptx`` (run without the simulator).

Nothing here is imported by tests or by the product; tests read only the .npz files.
"""

from __future__ import annotations

import os
import sys
import types

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def ptx_histogram() -> None:
    """Evidence for the float32 data flow: compile the reference kernel body to PTX."""
    import ast
    import collections
    import re

    from numba import boolean, cuda, float32, float64, int64, void

    src = open(f"{REF_SRC}/spectralmc/gbm.py").read()
    fn = next(
        n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "SimulateBlackScholes"
    )
    fn.decorator_list = []
    ns: dict[str, object] = {"cuda": cuda}
    exec("from math import exp, sqrt\n" + ast.unparse(fn), ns)  # noqa: S102 - reference code, read-only
    sig = void(float32[:, ::1], int64, float64, float64, float64, float64, float64, boolean)
    ptx, _ = cuda.compile_ptx(ns["SimulateBlackScholes"], sig, cc=(9, 0))
    ops = collections.Counter(re.findall(r"^\s+([a-z][a-z0-9_.]+)\s", ptx, flags=re.M))
    keep = {k: v for k, v in ops.items() if re.match(r"(ld\.global|st\.global|cvt|fma|mul\.f|add\.f|ex2|sqrt)", k)}
    for k in sorted(keep):
        print(f"{k:28s} {keep[k]}")


def install_shims():
    os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
    import numpy as np

    recorded: list[np.ndarray] = []

    # ---------------- cupy stand-in (NumPy) ----------------
    cp = types.ModuleType("cupy")
    for name in (
        "float32 float64 complex64 complex128 dtype ndarray linspace exp mean expand_dims "
        "asarray maximum allclose zeros empty"
    ).split():
        setattr(cp, name, getattr(np, name))
    cp.fft = np.fft

    # CuPy indexing/ufuncs give 0-d arrays where NumPy gives scalars; accept both wherever the
    # reference's Pydantic result models check ``isinstance(x, cp.ndarray)`` (gbm.py:279-292).
    class _NdMeta(type):
        def __instancecheck__(cls, obj):
            return isinstance(obj, (np.ndarray, np.generic))

    cp.ndarray = _NdMeta("ndarray", (), {})

    class _Rng:
        def __init__(self, seed):
            self._g = np.random.default_rng(seed)

        def standard_normal(self, shape, dtype=np.float64):
            z = self._g.standard_normal(shape, dtype=dtype)
            recorded.append(z.copy())
            return z

    cp.random = types.SimpleNamespace(default_rng=_Rng)

    class _Stream:
        def __init__(self, non_blocking=False):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def synchronize(self):
            pass

    class _Event:
        ptr = 0

        def __init__(self, disable_timing=False):
            pass

        def record(self):
            pass

    cp.cuda = types.SimpleNamespace(
        Stream=_Stream, Event=_Event, runtime=types.SimpleNamespace(eventQuery=lambda p: 0)
    )
    sys.modules["cupy"] = cp

    # ---------------- numba simulator gaps ----------------
    from numba import cuda

    import numba.cuda.cudadrv.devicearray as dev  # the simulator's module; lacks the real class name

    dev.DeviceNDArray = getattr(dev, "DeviceNDArray", dev.FakeCUDAArray)
    cuda.as_cuda_array = lambda a: a

    # ---------------- torch handle: skip the import-time CUDA assert ----------------
    sys.path.insert(0, REF_SRC)
    import torch

    import spectralmc.runtime.torch_runtime as tr

    tr._TORCH_HANDLE = torch

    # ---------------- effects package without its __init__ ----------------
    eff = types.ModuleType("spectralmc.effects")
    eff.__path__ = [f"{REF_SRC}/spectralmc/effects"]
    sys.modules["spectralmc.effects"] = eff
    import importlib

    mc = importlib.import_module("spectralmc.effects.montecarlo")
    rng = importlib.import_module("spectralmc.effects.rng")
    for n in ("PathScheme", "ForwardNormalization", "GenerateNormals", "SimulatePaths", "ComputeFFT"):
        setattr(eff, n, getattr(mc, n))
    eff.CaptureRNGState = rng.CaptureRNGState
    # descriptive-only ADTs the data path never executes (gbm.py:342-397): light placeholders
    eff.EffectSequence = type("EffectSequence", (), {"__class_getitem__": classmethod(lambda c, i: c)})
    eff.StreamSync = type("StreamSync", (), {"__init__": lambda self, **k: None})
    eff.sequence_effects = lambda *a: a
    return cp, recorded


def main() -> None:
    if "--ptx" in sys.argv:
        ptx_histogram()
        return
    cp, recorded = install_shims()
    import numpy as np

    import spectralmc.gbm as ref_gbm
    from spectralmc.effects import ForwardNormalization, PathScheme
    from spectralmc.models.numerical import Precision
    from spectralmc.result import Success

    # compiled-typing adapter for float32 matrices (see module docstring)
    raw_kernel = ref_gbm.SimulateBlackScholes

    class _Launch:
        def __init__(self, cfg):
            self.cfg = cfg

        def __call__(self, io, *args):
            if io.dtype == np.float32:
                wide = io.astype(np.float64)
                raw_kernel[self.cfg](wide, *args)
                io[...] = wide.astype(np.float32)
            else:
                raw_kernel[self.cfg](io, *args)

    class _Kernel:
        def __getitem__(self, cfg):
            return _Launch(cfg)

    ref_gbm.SimulateBlackScholes = _Kernel()

    canonical = dict(X0=100.0, K=100.0, T=1.0, r=0.05, d=0.0, v=0.2)  # tests/test_gbm.py:146
    cases = []
    for prec in ("float64", "float32"):
        for scheme in (PathScheme.LOG_EULER, PathScheme.SIMPLE_EULER):
            for norm in (ForwardNormalization.RAW, ForwardNormalization.NORMALIZE):
                # BASELINE.json configs[0]: T=12, N=16, B=64
                cases.append(dict(name=f"c1_{prec}_{scheme.value}_{norm.value}", prec=prec, scheme=scheme,
                                  norm=norm, T=12, N=16, B=64, tpb=256, seed=42, contract=canonical))
    # ragged sizes: N not a power of two, P not a multiple of the block, T not a multiple of 4
    odd = dict(X0=37.5, K=41.0, T=2.5, r=-0.01, d=0.03, v=0.65)
    for prec in ("float64", "float32"):
        cases.append(dict(name=f"ragged_{prec}", prec=prec, scheme=PathScheme.LOG_EULER,
                          norm=ForwardNormalization.NORMALIZE, T=7, N=12, B=11, tpb=32, seed=5, contract=odd))
        cases.append(dict(name=f"single_step_{prec}", prec=prec, scheme=PathScheme.LOG_EULER,
                          norm=ForwardNormalization.RAW, T=1, N=16, B=32, tpb=64, seed=7, contract=canonical))
    # degenerate contracts that are legal (gbm.py:272-275): T = 0 and v = 0
    cases.append(dict(name="expiry_zero_float64", prec="float64", scheme=PathScheme.LOG_EULER,
                      norm=ForwardNormalization.RAW, T=3, N=8, B=4, tpb=32, seed=9,
                      contract=dict(X0=90.0, K=100.0, T=0.0, r=0.03, d=0.01, v=0.3)))
    cases.append(dict(name="vol_zero_float64", prec="float64", scheme=PathScheme.LOG_EULER,
                      norm=ForwardNormalization.NORMALIZE, T=3, N=8, B=4, tpb=32, seed=9,
                      contract=dict(X0=90.0, K=100.0, T=2.0, r=0.03, d=0.01, v=0.0)))

    # the headline depths: 252 steps (configs c2/c5) and 365 steps (c4) on a narrow matrix, so that the reference's
    # own kernel body pins the error accumulation over a full-length path (the reference never tests T > 16)
    hot = dict(X0=120.0, K=95.0, T=3.0, r=0.03, d=0.01, v=1.2)
    cases.append(dict(name="deep_c2_float32", prec="float32", scheme=PathScheme.LOG_EULER, norm=ForwardNormalization.RAW,
                      T=252, N=16, B=4, tpb=64, seed=7, contract=canonical))
    cases.append(dict(name="deep_c4_float64", prec="float64", scheme=PathScheme.LOG_EULER, norm=ForwardNormalization.NORMALIZE,
                      T=365, N=16, B=4, tpb=64, seed=31, contract=hot))
    cases.append(dict(name="deep_simple_euler_float32", prec="float32", scheme=PathScheme.SIMPLE_EULER,
                      norm=ForwardNormalization.NORMALIZE, T=252, N=8, B=8, tpb=32, seed=11, contract=hot))
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    if only:
        cases = [c for c in cases if any(c["name"].startswith(o) for o in only)]

    for case in cases:
        recorded.clear()
        sp = ref_gbm.SimulationParams(
            timesteps=case["T"], network_size=case["N"], batches_per_mc_run=case["B"],
            threads_per_block=case["tpb"], mc_seed=case["seed"], buffer_size=1, skip=0,
            dtype=Precision(case["prec"]),
        )
        cfg = ref_gbm.BlackScholesConfig(sim_params=sp, path_scheme=case["scheme"], normalization=case["norm"])
        engine = ref_gbm.BlackScholes(cfg)
        inputs = ref_gbm.BlackScholes.Inputs(**case["contract"])
        normals = recorded[0].copy()  # the matrix the engine is about to consume (buffer_size = 1)
        sr = engine._simulate(inputs).unwrap()
        pr = engine.price(inputs=inputs, sr_result=Success(sr)).unwrap()
        host = engine.get_host_price(pr)
        mat = pr.put_price.reshape(sp.batches_per_mc_run, sp.network_size)
        cf = cp.mean(cp.fft.fft(mat, axis=1), axis=0)  # gbm_trainer.py:814-817
        snap = engine.snapshot().unwrap()
        assert snap.sim_params.skip == 1
        out = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(
            out,
            contract=np.array([case["contract"][k] for k in ("X0", "K", "T", "r", "d", "v")], dtype=np.float64),
            scheme=case["scheme"].value, normalization=case["norm"].value,
            timesteps=case["T"], network_size=case["N"], batches=case["B"], threads_per_block=case["tpb"],
            normals=normals, times=sr.times, sims=sr.sims, forwards=sr.forwards, df=sr.df,
            put_price=pr.put_price, call_price=pr.call_price, underlying=pr.underlying,
            put_price_intrinsic=np.asarray(pr.put_price_intrinsic),
            call_price_intrinsic=np.asarray(pr.call_price_intrinsic),
            cf=cf,
            host=np.array([host.put_price_intrinsic, host.call_price_intrinsic, host.underlying,
                           host.put_convexity, host.call_convexity, host.put_price, host.call_price]),
        )
        print(f"{case['name']:48s} cf[0]={cf[0]:.6g} dtype={cf.dtype} -> {os.path.basename(out)}")

    if only:
        return
    # Sobol contracts from the reference's own sampler (sobol_sampler.py imports cleanly)
    from spectralmc.sobol_sampler import BoundSpec, SobolConfig, SobolSampler, build_domain_bounds

    bounds = dict(X0=(0.001, 10_000.0), K=(0.001, 20_000.0), T=(0.0, 10.0), r=(-0.2, 0.2), d=(-0.2, 0.2), v=(0.0, 2.0))
    dims = build_domain_bounds(ref_gbm.BlackScholes.Inputs, {k: BoundSpec(*b) for k, b in bounds.items()}).unwrap()
    rows = {}
    for seed, skip, n in ((42, 0, 16), (42, 5, 8), (31, 0, 64)):
        s = SobolSampler.create(ref_gbm.BlackScholes.Inputs, dims, config=SobolConfig(seed=seed, skip=skip)).unwrap()
        pts = s.sample(n).unwrap()
        rows[f"seed{seed}_skip{skip}_n{n}"] = np.array([[p.X0, p.K, p.T, p.r, p.d, p.v] for p in pts])
    np.savez_compressed(os.path.join(HERE, "sobol_contracts.npz"), **rows)
    print("sobol_contracts.npz", {k: v.shape for k, v in rows.items()})


if __name__ == "__main__":
    main()
