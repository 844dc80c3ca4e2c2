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


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 5])
def test_extract_cli_under_torchrun_writes_the_same_files(tmp_path, world):
    """`torchrun --nproc-per-node W -m amcpy_b200.main extract`: every rank extracts a contiguous shard of the flattened
    (modulation, frame, snr) index space (5 ranks: shards cut through modulations), rank 0 assembles and writes; the six
    feature files must be bitwise what the single-process stage writes.  Runs on any number of GPUs (ranks share
    devices when there are fewer GPUs than ranks)."""
    import numpy as np
    import scipy.io

    from amcpy_b200 import synth
    from amcpy_b200.config import Config, Paths, SignalConfig
    from amcpy_b200.feature_extraction import run_extraction

    outs = {}
    for name in ("single", "torchrun"):
        cfg = Config(paths=Paths(root=tmp_path / name), signals=SignalConfig(num_frames=5))
        cfg.paths.ensure_dirs()
        snrs = [float(v) for v in cfg.signals.snr_values.values()]
        synth.write_all_modulations_mat(cfg.paths.mat_data / cfg.paths.mat_filename, synth.dataset(snrs, 5, 2048, 23),
                                        cfg.signals.mat_info)
        if name == "single":
            run_extraction(cfg)
        else:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                   "--master-addr", "127.0.0.1", "--master-port", str(29519 + world), "-m", "amcpy_b200.main",
                   "--root", str(tmp_path / name), "--num-frames", "5", "extract"]
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
            assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        outs[name] = {m: scipy.io.loadmat(str(cfg.paths.calculated_features / f"{m}_features.mat"))[cfg.signals.mat_info[m]]
                      for m in cfg.signals.modulations_with_noise}
    for m, a in outs["single"].items():
        assert a.shape == (16, 5, 18) and a.dtype == np.float32 and np.array_equal(a, outs["torchrun"][m]), m
