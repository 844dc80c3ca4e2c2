"""The two rows after the hot path against their oracles: the feature-matrix consumer against
sklearn (oracle/consumer_oracle.py restates preprocessing.py with its transpose fixed) and the classifier checkpoint
against the reference's `load_model` key layout; plus the drop-in packaging (import name `amcpy`, console script)."""

import sys
import tomllib

import numpy as np
import pytest

from conftest import ROOT


def _matrices(cfg, seed=0):
    rng = np.random.default_rng(seed)
    return {m: (rng.standard_normal((16, cfg.signals.num_frames, 18)) * (1 + i) + 3 * i).astype(np.float32)
            for i, m in enumerate(cfg.signals.modulations_with_noise)}


@pytest.mark.parametrize("mode", ["training", "test"])
def test_host_consumer_is_preprocess_data_with_the_transpose_fixed(mode):
    from amcpy_b200.config import Config, SignalConfig
    from amcpy_b200.consumer import load_feature_set, split_indices, stack_features
    from oracle.consumer_oracle import preprocess_data_fixed

    cfg = Config(signals=SignalConfig(num_frames=25))
    mats = _matrices(cfg)
    want = preprocess_data_fixed(cfg, mats, mode)
    got = load_feature_set(cfg, mode, mats)
    # identical partition (train_test_split(random_state=42, stratify=y) is part of the reference's result) ...
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
    # ... and StandardScaler's numbers: float32 rows, mean / population std per column
    assert got[0].dtype == want[0].dtype == np.float32
    assert np.allclose(got[0], want[0], rtol=2e-6, atol=2e-6) and np.allclose(got[1], want[1], rtol=2e-6, atol=2e-6)
    assert np.allclose(got[4].mean_, want[4].mean_, rtol=1e-12) and np.allclose(got[4].scale_, want[4].scale_, rtol=1e-12)
    # split_indices is that same partition as indices
    x, y = stack_features(cfg, mode, mats)
    tr, te = split_indices(y, cfg.training.test_size, cfg.training.random_state)
    assert np.array_equal(y[tr], want[2]) and np.array_equal(y[te], want[3])
    assert np.allclose(got[4].transform(x)[te], want[1], rtol=2e-6, atol=2e-6)


def test_checkpoint_has_the_reference_layout_and_round_trips(tmp_path):
    """nn_model.py:175-185 writes {model_state_dict, model_id, config}; :201-219 loads the state dict into
    `AMCClassifier`, whose Sequential is the attribute `layers` -> keys layers.N.*"""
    import torch
    import torch.nn as nn

    from amcpy_b200.classifier import AMCClassifier, load_model, model_for, save_model
    from amcpy_b200.config import Config, Paths

    cfg = Config(paths=Paths(root=tmp_path))
    torch.manual_seed(3)
    model = model_for(cfg)
    with torch.no_grad():                                # non-trivial BatchNorm statistics
        model.train()
        model(torch.randn(64, 6))
    path = save_model(model, "abcd1234", cfg)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"model_state_dict", "model_id", "config"} and ck["model_id"] == "abcd1234"
    assert ck["config"] == cfg                           # the Config object itself, as the reference stores its own

    class ReferenceLayout(nn.Module):                    # nn_model.py:28-75 restated: what the reference's load_model builds
        def __init__(self):
            super().__init__()
            self.layers = nn.Sequential(
                nn.Linear(6, 26), nn.BatchNorm1d(26), nn.ReLU(), nn.Dropout(0.4),
                nn.Linear(26, 29), nn.BatchNorm1d(29), nn.ReLU(), nn.Dropout(0.4),
                nn.Linear(29, 30), nn.BatchNorm1d(30), nn.ReLU(), nn.Dropout(0.4),
                nn.Linear(30, 6), nn.Softmax(dim=1))

        def forward(self, x):
            return self.layers(x)

    ref = ReferenceLayout()
    assert list(ck["model_state_dict"]) == list(ref.state_dict())          # layers.0.weight ... layers.12.bias
    ref.load_state_dict(ck["model_state_dict"])                            # strict: nothing missing, nothing unexpected
    ref.eval()
    ours = load_model("abcd1234", cfg)
    assert isinstance(ours, AMCClassifier) and not ours.training
    x = torch.randn(10, 6)
    model.eval()
    with torch.no_grad():
        assert torch.equal(ref(x), ours(x)) and torch.equal(ours(x), model(x))
    # a checkpoint written by the reference layout loads here as well
    torch.save({"model_state_dict": ref.state_dict(), "model_id": "ffff0000", "config": cfg}, cfg.paths.trained_ann / "model-ffff0000.pt")
    with torch.no_grad():
        assert torch.equal(load_model("ffff0000", cfg)(x), ref(x))


def test_packaging_gives_the_amcpy_names(monkeypatch):
    """pyproject.toml: console script `amcpy` and the import name `amcpy` (shim/amcpy) resolve to the B200 path
    (/root/reference/pyproject.toml:56-57 `amcpy = "amcpy.main:main"`)."""
    meta = tomllib.loads((ROOT / "pyproject.toml").read_text())
    assert meta["project"]["scripts"]["amcpy"] == "amcpy_b200.main:main"
    assert set(meta["tool"]["setuptools"]["packages"]) == {"amcpy_b200", "amcpy"}
    assert meta["tool"]["setuptools"]["package-dir"]["amcpy"] == "shim/amcpy"
    monkeypatch.syspath_prepend(str(ROOT / "shim"))
    for name in [k for k in sys.modules if k == "amcpy" or k.startswith("amcpy.")]:
        monkeypatch.delitem(sys.modules, name)
    import amcpy.config
    import amcpy.feature_extraction
    import amcpy.features
    import amcpy.main
    import amcpy.nn_model
    import amcpy.preprocessing

    import amcpy_b200.features as impl
    from amcpy_b200.main import main

    assert amcpy.features.calculate_features is impl.calculate_features
    assert amcpy.features._FEATURE_FUNCTIONS is impl._FEATURE_FUNCTIONS
    for fid, name in enumerate(impl.FEATURE_NAMES, start=1):              # features.py:192-211: names and ids
        assert getattr(amcpy.features, name) is impl._FEATURE_FUNCTIONS[fid]
    assert amcpy.main.main is main and callable(amcpy.feature_extraction.run_extraction)
    assert amcpy.config.Config().signals.frame_size == 2048
    assert amcpy.nn_model.AMCClassifier is not None and callable(amcpy.preprocessing.preprocess_data)


def test_stage_shard_plan_covers_every_frame_once():
    from amcpy_b200.config import Config, SignalConfig
    from amcpy_b200.feature_extraction import assemble_matrix, plan_shards

    cfg = Config(signals=SignalConfig(num_frames=7))
    mods = cfg.signals.modulations_with_noise
    S = 16
    rows_of = {m: np.random.default_rng(i).standard_normal((S * 7, 18)) for i, m in enumerate(mods)}
    want = {m: assemble_matrix(rows_of[m], 0, S, 16, 7) for m in mods}
    for world in (1, 2, 3, 5, 8, 13):
        plan = plan_shards(cfg, {m: S for m in mods}, world)
        sizes = [sum(q1 - q0 for r, _, q0, q1 in plan if r == rk) for rk in range(world)]
        assert sum(sizes) == 6 * S * 7 and max(sizes) - min(sizes) <= 1       # balanced: every GPU busy
        got = {m: np.zeros((16, 7, 18), dtype=np.float32) for m in mods}
        seen = {m: np.zeros(S * 7, dtype=int) for m in mods}
        for _, m, q0, q1 in plan:
            assemble_matrix(rows_of[m][q0:q1].astype(np.float32), q0, S, 16, 7, got[m])
            seen[m][q0:q1] += 1
        for m in mods:
            assert (seen[m] == 1).all() and np.array_equal(got[m], want[m])
    # a file with more SNR rows than the config asks for: the surplus rows are computed by nobody's matrix
    fm = assemble_matrix(np.ones((20 * 3, 18)), 0, 20, 16, 3)
    assert fm.shape == (16, 3, 18) and (fm == 1).all()


@pytest.mark.gpu
def test_device_consumer_matches_sklearn_oracle():
    import torch

    from amcpy_b200 import _native as nat
    from amcpy_b200.config import Config, SignalConfig
    from amcpy_b200.consumer import load_feature_set_device
    from oracle.consumer_oracle import preprocess_data_fixed

    nat.require_cuda()
    cfg = Config(signals=SignalConfig(num_frames=25))
    mats = _matrices(cfg, seed=5)
    feats = {m: torch.from_numpy(v).cuda() for m, v in mats.items()}
    for mode in ("training", "test"):
        want = preprocess_data_fixed(cfg, mats, mode)
        got = load_feature_set_device(cfg, feats, mode)
        assert all(t.is_cuda for t in got[:4]) and got[0].dtype == torch.float32 and got[2].dtype == torch.int64
        assert np.array_equal(got[2].cpu().numpy(), want[2]) and np.array_equal(got[3].cpu().numpy(), want[3])
        assert np.allclose(got[0].cpu().numpy(), want[0], rtol=2e-6, atol=2e-6)
        assert np.allclose(got[1].cpu().numpy(), want[1], rtol=2e-6, atol=2e-6)
        assert np.allclose(got[4].mean_, want[4].mean_, rtol=1e-12) and np.allclose(got[4].scale_, want[4].scale_, rtol=1e-12)
