"""bench.py's contract where it can be checked without a GPU: the reference arm prints one JSON line with the keys the driver
reads, and the product arm refuses to run without a CUDA device (there is no CPU fallback to fall into)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          env=dict(os.environ, **(env or {})), timeout=600)


def test_reference_arm_prints_the_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_engine.so")) and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref not built here")
    p = _run("--impl", "reference", "--points", "20000", "--cpu-sample", "4000", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "nn_queries_per_s" and line["unit"] == "queries/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_is_rank_zero_only():
    p = _run("--impl", "reference", "--points", "20000", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run("--points", "20000", "--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-e2e")
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
