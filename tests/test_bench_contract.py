"""bench.py's reference arm runs without a GPU and prints exactly one JSON line
with the keys the driver reads (the CUDA arm is exercised by the GPU box)."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "config2_small",
         "--steps", "1", "--warmup", "0"],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "halos_per_s" and d["unit"] == "halos/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "halos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_our_arm_refuses_to_run_without_a_gpu():
    """no CPU fallback: on a box without CUDA the product arm exits with an error"""
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "config2_small", "--steps", "1",
                        "--warmup", "0", "--no-cpu-baseline", "--no-e2e"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert r.stdout.strip() == ""
