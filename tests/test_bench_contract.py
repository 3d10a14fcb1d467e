"""bench.py contract checks that run without a GPU: the reference arm (CPU by definition) prints one JSON line with
the agreed keys, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(*args, timeout=240):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True,
                          text=True, timeout=timeout)


@pytest.mark.parametrize("workload,codes", [("tokenize", "400"), ("kmeans", "300")])
def test_reference_arm_json_line(workload, codes):
    r = _run("--impl", "reference", "--workload", workload, "--codes", codes, "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gesture chunks quantized/sec" and line["unit"] == "chunks/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["steps"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["gpu_launches"] == 0
    cb, e2e = line["cpu_baseline"], line["e2e"]
    assert cb["value"] == line["value"] and cb["cores"] >= 1 and cb["kind"] in ("port", "reference") and cb["sample"]
    assert e2e == {"value": line["value"], "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "1", timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
