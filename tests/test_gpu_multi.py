"""Multi-GPU parity (needs >= 2 B200s; skipped on a single-GPU box): the point-sharded solve with the
NCCL / peer-memory exchange must reproduce the single-GPU solve (tools/dist_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("peer_exchange", ["1", "0"])
def test_sharded_solve_equals_single_gpu(peer_exchange):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, MMBA_PEER_XCHG=peer_exchange)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29541" if peer_exchange == "1" else "29542", os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MISMATCH" not in out.stdout and out.stdout.count("-> OK") >= 8
