"""bench.py's reference arm on the host cores: the JSON line the driver parses (the contract in the task statement), and that the
GPU arm refuses to run without a CUDA device instead of falling back to anything on the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

from tests.conftest import ROOT


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("env-steps/sec ship-ice-v0") and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine WITHOUT a CUDA device")
def test_gpu_arm_fails_loudly_without_a_device():
    r = _run(["--steps", "1", "--warmup", "0", "--envs", "8"], timeout=300)
    assert r.returncode != 0, "the product path must not fall back to the CPU"
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{") and '"value"' in ln]
