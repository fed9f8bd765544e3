"""bench.py contract, CPU side: the reference arm prints exactly one JSON line on stdout with the
keys the driver reads, and the B200 arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, env=env, timeout=600)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "lbm_ref_fast")), reason="oracle/_ref not built")
def test_reference_arm_prints_one_json_line():
    r = run_bench("--impl", "reference", "--workload", "c1", "--steps", "4", "--warmup", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "MLUPS (fp64 D2Q9)" and j["unit"] == "MLUPS"
    assert j["value"] > 1.0 and j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 4 and j["warmup"] == 3
    assert j["dtype"] == "f64" and j["vs_baseline"] is None and "2048x512" in j["config"]["workload"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == j["value"] and "2048x512" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0


def test_b200_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    r = run_bench("--workload", "c1", "--steps", "4")
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]  # no number without a GPU
