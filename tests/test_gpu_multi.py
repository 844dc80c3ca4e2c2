"""Multi-GPU checks (need >= 2 visible GPUs; skipped on the 1-GPU box): the kernel-epilogue gather into the
root's matrix over NVLink peer memory (sharding.extract_sharded_to_root) and the NCCL gather must both
reproduce the single-GPU result bit for bit."""
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.gpu
def test_peer_store_gather_matches_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", str(ROOT / "tools" / "p2p_gather_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "peer-store gather bitwise == single GPU: True" in res.stdout
