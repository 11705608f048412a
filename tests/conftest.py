"""Test configuration.

Markers: ``gpu`` — needs a CUDA device (the parity tests proper; they call through the C ABI).
Everything else runs on CPU: the oracle against the golden vectors, the host logic, and the
"library loads and exports every declared symbol" check.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: requires a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config: pytest.Config, items: list[pytest.Item]) -> None:
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_cases() -> list[str]:
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f != "sobol_contracts.npz" and not f.startswith("cvnn"))


def load_golden(name: str) -> dict[str, np.ndarray]:
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(params=golden_cases())
def golden(request: pytest.FixtureRequest) -> dict[str, np.ndarray]:
    g = load_golden(request.param)
    g["name"] = request.param
    return g
