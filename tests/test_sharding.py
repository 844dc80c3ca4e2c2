"""CPU: multi-GPU host logic with world_size-2 gloo (the N>1 path of bench.py / run_extraction
uses the same shard math; on GPUs the backend is NCCL)."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def test_shard_ranges_partition_exactly():
    from amcpy_b200.sharding import shard_range, unflatten

    for total in (0, 1, 7, 48000, 2520000):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    assert unflatten(0, 16, 500) == (0, 0, 0)
    assert unflatten(16 * 500 + 501, 16, 500) == (1, 1, 1)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from amcpy_b200.sharding import gather_features, shard_range

        lo, hi = shard_range(total, rank, world)
        # every "frame" i gets the row [i, i+0.5, ...]: the gather must restore global order
        local = (torch.arange(lo, hi, dtype=torch.float64)[:, None] + torch.arange(18, dtype=torch.float64) / 36.0)
        full = gather_features(local, total)
        want = torch.arange(total, dtype=torch.float64)[:, None] + torch.arange(18, dtype=torch.float64) / 36.0
        q.put((rank, bool(torch.equal(full, want))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [11, 48])
def test_gather_features_world2_gloo(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_consumer_semantics_without_mat_round_trip():
    from amcpy_b200.config import Config, SignalConfig
    from amcpy_b200.consumer import load_feature_set, stack_features

    cfg = Config(signals=SignalConfig(num_frames=10))
    rng = np.random.default_rng(0)
    mats = {m: rng.standard_normal((16, 10, 18)).astype(np.float32) + i for i, m in
            enumerate(cfg.signals.modulations_with_noise)}
    x, y = stack_features(cfg, "training", mats)
    assert x.shape == (6 * 6 * 10, 6) and x.dtype == np.float32 and y.shape == (360,)
    # 0-based columns (2,4,6,8,12,14) of SNR index 10, modulation 0, frame 3 - graphics.py:46 convention
    assert np.array_equal(x[3], mats["BPSK"][10, 3, [2, 4, 6, 8, 12, 14]])
    xtr, xte, ytr, yte, sc = load_feature_set(cfg, "training", mats)
    assert xtr.shape == (288, 6) and xte.shape == (72, 6)
    assert sorted(np.bincount(yte).tolist()) == [12] * 6          # stratified
    allx = np.concatenate([xtr, xte])
    assert np.allclose(allx.mean(0), 0, atol=1e-5) and np.allclose(allx.std(0), 1, atol=1e-4)
