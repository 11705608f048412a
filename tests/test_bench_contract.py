"""bench.py's reference arm runs without a GPU: check the JSON contract of its one output line."""

from __future__ import annotations

import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def _run(extra_env=None, *args):
    env = dict(os.environ, **(extra_env or {}))
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--cpu-sample-batches", "64", *args], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    return [line for line in proc.stdout.splitlines() if line.startswith("{")]


def test_reference_arm_prints_one_contract_line() -> None:
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gbm_path_steps_per_sec" and d["unit"] == "path-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("c2:") and d["config"]["timesteps"] == 252 and d["config"]["network_size"] == 128
    base = d["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] == d["value"] and "64 of 65536 batch rows" in base["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_only_rank_zero_works_under_torchrun() -> None:
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2") == []
    lines = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, "--gpus", "2")
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2


def _inner_loops(kernel_prefix: str):
    """Instruction mix of the innermost loops of one kernel of the built library (cuobjdump -sass)."""
    import collections
    import re
    import shutil

    import pytest

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "spectralmc_b200", "lib", "libspectralmc_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, timeout=600).stdout
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        if not part.startswith(kernel_prefix):
            continue
        ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4})\*/\s+(.*?);", part)]
        spans = []
        for a, t in ins:
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                spans.append((int(m.group(1), 16), a))
        loops = []
        for lo, hi in spans:
            if any(l2 > lo and h2 < hi for l2, h2 in spans):
                continue
            body = [re.sub(r"^@!?U?P\d\s+", "", t).split()[0] for a, t in ins if lo <= a <= hi]
            loops.append(collections.Counter(body))
        return loops
    raise AssertionError(f"kernel {kernel_prefix} not found in {lib}")


def test_roofline_constants_match_the_built_kernels() -> None:
    """bench.py's rooflines multiply the measured rate by per-path-step instruction counts; re-count them from
    the SASS of the library that is actually built, so a kernel edit cannot silently invalidate them."""
    sys.path.insert(0, ROOT)
    import bench

    f32 = _inner_loops(bench.F32_LOOP["kernel"])[0]  # the hot loop comes first: two 6-normal blocks per iteration
    assert sum(f32.values()) == bench.F32_LOOP["instructions"], dict(f32)
    assert sum(v for k, v in f32.items() if k.startswith("MUFU")) == bench.F32_LOOP["mufu"]
    assert f32["IMAD.WIDE.U32"] == 32 and not any(k.startswith(("DFMA", "DMUL", "DADD")) for k in f32)
    f64 = _inner_loops(bench.F64_LOOP["kernel"])[0]  # one block = two pairs = four float64 normals per iteration
    assert sum(f64.values()) == bench.F64_LOOP["instructions"], dict(f64)
    assert sum(v for k, v in f64.items() if k.startswith(("DFMA", "DMUL", "DADD"))) == bench.F64_LOOP["fp64"]


def test_both_arms_describe_the_same_config() -> None:
    """The driver compares `config` of the two arms key by key: it must not depend on anything only one arm knows."""
    sys.path.insert(0, ROOT)
    import bench

    for gpus in (1, 2, 8):
        assert "collective" not in bench.workload_config("c2", gpus)
    d = json.loads(_run()[0])
    assert d["config"] == bench.workload_config("c2", 1)
