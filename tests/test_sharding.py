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


# ------------------------------------------------------------------ the stage's N > 1 path (gloo, CPU)
def _fake_rows(key: str, q0: int, q1: int):
    """Stand-in for the GPU features of frames q0..q1 of one variable: any deterministic function of (variable, q)."""
    base = float(sum(key.encode()))
    q = np.arange(q0, q1, dtype=np.float64)
    return base + q[:, None] * 0.25 + np.arange(18, dtype=np.float64)[None, :] / 64.0


def _stage_worker(rank, world, port, root, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    try:
        import amcpy_b200._native as nat
        import amcpy_b200.feature_extraction as fe
        from amcpy_b200.config import Config, Paths, SignalConfig

        nat.require_cuda = lambda: 1                       # host logic under test: no device in this container

        def fake_planar(arr, n_snr, n_frames, frame_size, device=0, q_range=None):
            S = arr.shape[0]
            key = [k for k, v in fe_src.planar.items() if v is arr][0]
            if q_range is None:
                return fe.assemble_matrix(_fake_rows(key, 0, S * n_frames), 0, S, n_snr, n_frames)
            return _fake_rows(key, *q_range)

        cfg = Config(paths=Paths(root=Path(root)), signals=SignalConfig(num_frames=3, frame_size=64))
        orig_init = fe._MatSource.__init__

        def capture(self, *a, **k):
            orig_init(self, *a, **k)
            nonlocal fe_src
            fe_src = self

        fe_src = None
        fe._MatSource.__init__ = capture
        fe.extract_modulation_planar = fake_planar
        res = fe.extract_all(cfg)
        q.put((rank, sorted(res)))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_stage_shards_assemble_to_the_single_process_files_gloo(tmp_path, world):
    """run_extraction under WORLD_SIZE ranks (gloo): contiguous shards of the flattened (modulation, frame, snr) space,
    gathered and written by rank 0, give the files a single process writes (the GPU features replaced by a
    deterministic stand-in: this is the host-side logic of the N > 1 path)."""
    import scipy.io

    from amcpy_b200 import synth
    from amcpy_b200.config import Config, Paths, SignalConfig

    outs = {}
    for name, w in (("single", 1), ("sharded", world)):
        root = tmp_path / name
        cfg = Config(paths=Paths(root=root), signals=SignalConfig(num_frames=3, frame_size=64))
        cfg.paths.ensure_dirs()
        snrs = [float(v) for v in cfg.signals.snr_values.values()]
        synth.write_all_modulations_mat(cfg.paths.mat_data / cfg.paths.mat_filename, synth.dataset(snrs, 4, 80, 3),
                                        cfg.signals.mat_info)
        ctx = mp.get_context("spawn")
        qq = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_stage_worker, args=(r, w, port, str(root), qq)) for r in range(w)]
        for p in procs:
            p.start()
        got = sorted(qq.get(timeout=180) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert got[0] == (0, sorted(cfg.signals.modulations_with_noise))          # rank 0 holds every matrix
        assert all(names == [] for _, names in got[1:])                            # the other ranks hold none
        outs[name] = {m: scipy.io.loadmat(str(cfg.paths.calculated_features / f"{m}_features.mat"))[cfg.signals.mat_info[m]]
                      for m in cfg.signals.modulations_with_noise}
    for m, a in outs["single"].items():
        assert a.shape == (16, 3, 18) and a.dtype == np.float32
        assert np.array_equal(a, outs["sharded"][m]), m
    a = outs["single"]["QPSK"]
    assert a[5, 2, 0] == np.float32(_fake_rows("signal_qpsk", 5 + 16 * 2, 5 + 16 * 2 + 1)[0, 0])   # q = snr + S*frame
