"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys the driver reads,
and the product arm refuses to produce a number without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    # 4 views keep the CPU suite short; the default (16 views, BASELINE config 2) differs only in the view count
    p = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--views", "4")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["metric"] == "4-view 256^2 DDIM denoise-steps/sec" and d["value"] > 0
    sys.path.insert(0, ROOT)
    import bench
    assert bench.metric_name(16, 32) == "16-view 256^2 DDIM denoise-steps/sec"   # BASELINE.json's metric by default
    assert bench.main.__module__ == "bench"
    assert d["config"]["workload"].startswith("FLAME face")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"]
    assert "4 of 4 views" in cb["sample"]   # full steps of the workload, never a view sub-sample scaled up
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["ms_per_step"] - 1000.0 / d["value"]) < 1e-6 * d["ms_per_step"]


def test_reference_arm_is_rank0_only_under_torchrun():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", PYTHONDONTWRITEBYTECODE="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return  # on a GPU box the product arm runs; covered by the driver's bench
    p = run_bench("--steps", "1", "--warmup", "0", "--no-cpu", timeout=300)
    assert p.returncode != 0
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]
