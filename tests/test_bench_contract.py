"""bench.py's reference arm on CPU: one JSON line with the contract's keys (the b200 arm needs a GPU and is
exercised by the driver; its line is built by the same function)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--edges", "200000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "edge_checks_per_s" and d["unit"] == "edges/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic" and d["n_gpus"] == 1
    assert "workload" in d["config"] and "config4" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["value"] > 0


def test_reference_arm_other_ranks_are_silent():
    """Under torchrun (N > 1) rank 0 alone runs and prints the reference arm; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, cwd=ROOT,
                       env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_build_the_same_config_object():
    """The driver compares the two arms' `config` key by key: both come from bench.full_config."""
    sys.path.insert(0, ROOT)
    import bench
    for world, strong in ((1, False), (8, False), (8, True)):
        a = bench.full_config(10 ** 7, world, strong)
        b = bench.full_config(10 ** 7, world, strong)
        assert a == b and set(a) == {"workload", "edges_per_gpu", "dim", "boxes", "l2", "sharding", "scaling_mode", "merge"}
        assert ("none" in a["merge"]) == (world == 1)


def test_gpu_arm_fails_loudly_without_cuda():
    """No CPU fallback: without a CUDA device the b200 arm exits non-zero instead of measuring something else."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--edges", "1000"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "CUDA" in r.stderr or "cuda" in r.stderr
