"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm (CPU oracle port) prints one JSON
line with the agreed keys, and the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--small", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "iters/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["config"]["workload"].startswith("tracking_replica_")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_product_arm_fails_loudly_without_cuda():
    p = _run("--small", "--steps", "1", "--warmup", "0", "--no-cpu", timeout=300)
    assert p.returncode != 0
    assert "cuda" in (p.stderr + p.stdout).lower()
