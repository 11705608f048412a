"""The C-ABI library loads and exports every symbol include/spectralmc_b200.h declares.
No compute call is made here (there is no GPU on the CPU test box)."""

from __future__ import annotations

import ctypes
import os
import re

from tests.conftest import ROOT


def _declared() -> set[str]:
    text = open(os.path.join(ROOT, "include", "spectralmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(smc_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol() -> None:
    from spectralmc_b200 import _cabi

    lib = ctypes.CDLL(_cabi.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 20
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)


def test_version_and_error_string() -> None:
    from spectralmc_b200 import _cabi

    assert _cabi.version() == 103
    assert isinstance(_cabi.LIB.smc_last_error(), bytes)


def test_workspace_queries_are_pure_host_functions() -> None:
    from spectralmc_b200 import _cabi

    assert _cabi.LIB.smc_means3_workspace_bytes(10_000) > 0
    assert _cabi.LIB.smc_normalize_rows_workspace_bytes(12, 1024) > 0
    assert _cabi.LIB.smc_cf_fft_mean_workspace_bytes(64, 16, 0) > 0
    args = _cabi.make_fused_args(None, 4, 12, 16, 64, __import__("torch").float32, 0, 1, 42, 0)
    raw = _cabi.LIB.smc_cf_fused_workspace_bytes(ctypes.byref(args))
    args.normalization = 0
    assert _cabi.LIB.smc_cf_fused_workspace_bytes(ctypes.byref(args)) > raw > 0


def test_argument_validation_needs_no_device() -> None:
    """Bad arguments are rejected before any CUDA call, with a message."""
    from spectralmc_b200 import _cabi

    rc = _cabi.LIB.smc_philox_normals(None, 4, 4, 0, 1, 0, None)
    assert rc == 1 and b"NULL" in _cabi.LIB.smc_last_error()
    rc = _cabi.LIB.smc_philox_normals(None, 0, 4, 0, 1, 0, None)
    assert rc == 1 and b"shape" in _cabi.LIB.smc_last_error()
    rc = _cabi.LIB.smc_gbm_paths_inplace(ctypes.c_void_p(16), 4, 4, 0, 0.1, 1.0, 0.0, 0.0, 0.2, 0, 48, None)
    assert rc == 1 and b"threads_per_block" in _cabi.LIB.smc_last_error()
    rc = _cabi.LIB.smc_payoff(ctypes.c_void_p(16), 4, 7, 1.0, 1.0, None, None, None)
    assert rc == 1 and b"dtype" in _cabi.LIB.smc_last_error()


def test_product_never_imports_the_oracle() -> None:
    pkg = os.path.join(ROOT, "spectralmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_missing_library_fails_loudly() -> None:
    """No fallback: importing the package without the shared library is an ImportError."""
    import subprocess
    import sys

    env = dict(os.environ, SPECTRALMC_B200_LIB="/nonexistent/libspectralmc_b200.so", PYTHONPATH=ROOT)
    proc = subprocess.run([sys.executable, "-c", "import spectralmc_b200"], env=env, capture_output=True, text=True)
    assert proc.returncode != 0
    assert "ImportError" in proc.stderr and "no fallback" in proc.stderr


def test_peer_exchange_argument_validation_needs_no_device() -> None:
    from spectralmc_b200 import _cabi

    torch = __import__("torch")
    # data + per-contract flags + small all-reduce region + its flags + the status word, in 8-byte cells
    assert _cabi.LIB.smc_p2p_buffer_bytes(4, 16, 2) == (2 * 2 * 4 * 16 + 2 * 2 * 4 + 2 * 2 * 4 + 2 * 2 + 1) * 8
    assert _cabi.LIB.smc_p2p_buffer_bytes(4, 16, 17) == 0  # at most 16 peers
    args = _cabi.make_fused_args(None, 4, 12, 16, 64, torch.float32, 0, _cabi.SMC_RAW, 42, 0, batch_begin=0, batch_end=32)
    args.contracts = 16  # any non-NULL value: validation happens before the first device access
    group = _cabi.P2PGroup()
    group.rank, group.world, group.capacity_contracts, group.network_size, group.epoch = 0, 2, 4, 16, 0
    call = lambda: _cabi.LIB.smc_cf_fused_p2p(ctypes.byref(args), ctypes.byref(group), 16, 16, 1 << 20, None)  # noqa: E731
    assert call() == 1 and b"epoch" in _cabi.LIB.smc_last_error()
    group.epoch = 1
    assert call() == 1 and b"buffer of rank 0 is NULL" in _cabi.LIB.smc_last_error()
    group.buffers[0], group.buffers[1] = 16, 16
    group.capacity_contracts = 3
    assert call() == 1 and b"sized for 3 contracts" in _cabi.LIB.smc_last_error()
    group.capacity_contracts, group.rank = 4, 2
    assert call() == 1 and b"bad rank" in _cabi.LIB.smc_last_error()
    group.rank = 0
    # the device-free pre-flight check callers run BEFORE they advance the epoch sees the same errors
    check = lambda ws: _cabi.LIB.smc_cf_fused_p2p_check(ctypes.byref(args), ctypes.byref(group), ws)  # noqa: E731
    assert check(1 << 20) == 0
    assert check(16) == 3 and b"workspace" in _cabi.LIB.smc_last_error()
    args.normalization = _cabi.SMC_NORMALIZE
    assert call() == 1 and b"NORMALIZE" in _cabi.LIB.smc_last_error()
    assert check(1 << 20) == 1 and b"NORMALIZE" in _cabi.LIB.smc_last_error()


def test_tile_plan_depends_on_the_problem_shape_only() -> None:
    """How a simulation is cut into CTAs (smc_cf_fused_plan, no device): tiles cover the local rows exactly, are
    multiples of the row-lane count, never smaller than 16 384 path-steps, and the ticket tree ends in a root of at
    most 16 vectors — at BASELINE's shapes and at ragged ones."""
    from spectralmc_b200 import _cabi

    torch = __import__("torch")
    shapes = {"c1": (1, 12, 16, 64, torch.float64), "c2": (1, 252, 128, 65536, torch.float32), "c3": (1024, 1, 16, 4096, torch.float32),
              "c4": (512, 365, 256, 4096, torch.float64), "c5 share": (4096, 252, 128, 131072, torch.float32),
              "ragged": (3, 7, 24, 1001, torch.float32), "wide": (2, 5, 1000, 77, torch.float64)}
    for name, (C, T, N, B, dtype) in shapes.items():
        for lo, hi in ((0, B), (B // 3, B - B // 5)):
            args = _cabi.make_fused_args(None, C, T, N, B, dtype, 0, _cabi.SMC_RAW, 7, 0, batch_begin=lo, batch_end=hi)
            args.contracts = 16
            plan = _cabi.cf_fused_plan(args)
            rows = hi - lo
            assert plan["row_lanes"] == max(1, 256 // min(N, 256)), name
            assert plan["tile_rows"] % plan["row_lanes"] == 0 and plan["tile_rows"] >= 1, name
            assert plan["tail_tile_rows"] % plan["row_lanes"] == 0 and 1 <= plan["tail_tile_rows"] <= plan["tile_rows"], name
            # main tiles of tile_rows rows, then (single-contract launches) a fine tail of tail_tile_rows rows: together exactly `rows`
            main, tail = plan["main_tiles"], plan["tiles"] - plan["main_tiles"]
            covered_max = main * plan["tile_rows"] + tail * plan["tail_tile_rows"]
            assert covered_max >= rows and covered_max - rows < plan["tile_rows"] + plan["tail_tile_rows"], (name, plan, rows)
            assert tail == 0 or (C == 1 and tail * plan["tail_tile_rows"] <= rows // 4 + plan["tail_tile_rows"]), (name, plan)
            assert plan["tail_tile_rows"] * N * T >= 16384 or plan["tiles"] == 1 or plan["tail_tile_rows"] == plan["row_lanes"], (name, plan)
            assert 1 <= plan["root_fan_in"] <= 16, (name, plan)
            width, level = plan["tiles"], 0
            while width > 16:
                width, level = -(-width // 16), level + 1
            assert (level, width) == (plan["levels"], plan["root_fan_in"]), (name, plan)
            assert _cabi.LIB.smc_cf_fused_workspace_bytes(ctypes.byref(args)) >= plan["tiles"] * C * N * 8
    c2 = _cabi.make_fused_args(None, 1, 252, 128, 65536, torch.float32, 0, _cabi.SMC_RAW, 7, 0)
    c2.contracts = 16
    assert _cabi.LIB.smc_cf_fused_launch_count(ctypes.byref(c2)) == 1  # the RAW step is ONE kernel
    c2.normalization = _cabi.SMC_NORMALIZE
    assert _cabi.LIB.smc_cf_fused_launch_count(ctypes.byref(c2)) == 2


def test_stream_version_is_validated_without_a_device() -> None:
    from spectralmc_b200 import _cabi

    torch = __import__("torch")
    args = _cabi.make_fused_args(None, 1, 12, 16, 64, torch.float32, 0, _cabi.SMC_RAW, 42, 0, stream_version=2)
    args.contracts = 16
    assert _cabi.LIB.smc_cf_fused(ctypes.byref(args), 16, 16, 1 << 20, None) == 1 and b"stream_version" in _cabi.LIB.smc_last_error()
    assert _cabi.LIB.smc_philox_normals_v(16, 4, 4, 0, 1, 0, 5, None) == 1 and b"stream_version" in _cabi.LIB.smc_last_error()
    assert _cabi.make_fused_args(None, 1, 12, 16, 64, torch.float32, 0, _cabi.SMC_RAW, 42, 0).stream_version == _cabi.SMC_STREAM_PHILOX10
