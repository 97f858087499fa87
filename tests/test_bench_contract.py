"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm
prints one JSON line with the agreed keys, on the same metric / unit / config as the product arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tinyE",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, check=True)
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert "workload" in d["config"] and d["config"]["n_steps"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert cb["path_parallel_variant"]["value"] > 0 and "NOT the reference" in cb["path_parallel_variant"]["note"]


def test_reference_arm_other_ranks_exit_quietly():
    env = {**os.environ, "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--config", "tinyE",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
