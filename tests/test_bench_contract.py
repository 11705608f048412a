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
