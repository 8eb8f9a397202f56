"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, and the
engine arm refuses to run without a B200 (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["unit"] == "observations/s" and line["value"] > 0 and line["vs_baseline"] is None
    # the unmodified reference module where /root/reference exists (build container), its restatement elsewhere
    assert line["cpu_baseline"]["kind"] == ("reference" if os.path.exists("/root/reference/bundleAdjuster.py") else "port")
    assert line["cpu_baseline"]["cores"] == 1 and line["scaling"] == "strong"
    full = line["cpu_baseline"]["full_size"]       # full-size wall time of the unmodified reference (golden file)
    assert full["config"] == "C4" and full["wall_s"] > 100 and full["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_runs_on_rank_zero_only():
    """Under torchrun (N > 1) rank 0 alone runs and prints the reference arm; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout, out.stderr[-2000:])


def test_engine_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
