"""Sharded optimize() with one process per GPU (needs >= 2 GPUs; skipped on a single-GPU box): the exchanges over
peer-mapped mailboxes and over NCCL must both reproduce the in-process reference (scripts/check_nccl_sharded.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_two_ranks_match_the_in_process_reference(exchange):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, CHECK_EXCHANGE=exchange, CHECK_BATCH="16384", CHECK_CYCLES="4", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "check_nccl_sharded.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout
