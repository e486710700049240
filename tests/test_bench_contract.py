"""bench.py's reference arm runs on the host cores (no GPU): its JSON line must carry the keys the
driver reads, on the same metric / unit / config as the CUDA arm; ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_line(built):
    p = _run({})
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "chunks/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("64^3 DC chunks/sec") and line["config"]["workload"].startswith("configs[1]")
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1 and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent(built):
    p = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cuda_arm_refuses_without_a_device(built):
    """the CUDA arm has no CPU fallback: without a device it says so and exits non-zero"""
    import torch
    if torch.cuda.is_available():
        return
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout) and not p.stdout.strip().startswith("{")
