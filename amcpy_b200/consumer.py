"""Feature-matrix consumer: what reads calculated-features/*.mat next.

Restates the intended semantics of the reference's `preprocess_data`
(/root/reference/src/amcpy/preprocessing.py:13-75): per modulation and per selected SNR take the
`FeatureConfig.used` columns (0-based, exactly as graphics.py:46 / preprocessing.py:55 index them),
stack to (n_samples, n_used) float32, standardise (zero mean / unit population variance, i.e.
sklearn's StandardScaler), stratified 80/20 split with random_state=42.  The reference's version
raises ValueError at preprocessing.py:55 (a (6, N) block assigned into an (N, 6) slot - SURVEY.md
App. B.3); the transpose is fixed here, nothing else is changed.
"""

from __future__ import annotations

import numpy as np
import scipy.io

from .config import Config


class Standardizer:
    """mean_/scale_ with StandardScaler's conventions (population std, zero variance -> scale 1)."""

    def fit(self, x: np.ndarray) -> "Standardizer":
        x64 = x.astype(np.float64)
        self.mean_ = x64.mean(axis=0)
        var = x64.var(axis=0)
        self.scale_ = np.where(var > 0, np.sqrt(var), 1.0)
        return self

    def transform(self, x: np.ndarray) -> np.ndarray:
        return ((x - self.mean_) / self.scale_).astype(x.dtype if x.dtype.kind == "f" else np.float64)


def stack_features(cfg: Config, mode: str = "training", matrices: dict | None = None):
    """(x float32 (n_samples, n_used), y int64).  `matrices` may hold {MOD: (n_snr, n_frames, 18)}
    arrays straight from the extraction (skipping the .mat round trip)."""
    s, t, f = cfg.signals, cfg.training, cfg.features
    snr_axis = t.training_snr if mode == "training" else t.all_snr
    cols = list(f.used)
    xs, ys = [], []
    for mod_idx, mod in enumerate(s.modulations_with_noise):
        if matrices is not None:
            m = np.asarray(matrices[mod])
        else:
            m = scipy.io.loadmat(str(cfg.paths.calculated_features / f"{mod}_features.mat"))[s.mat_info[mod]]
        for snr in snr_axis:
            xs.append(np.asarray(m[snr, : s.num_frames][:, cols], dtype=np.float32))   # (n_frames, n_used)
            ys.append(np.full(s.num_frames, s.labels[mod_idx], dtype=np.int64))
    return np.concatenate(xs), np.concatenate(ys)


def load_feature_set(cfg: Config, mode: str = "training", matrices: dict | None = None):
    """x_train, x_test, y_train, y_test, scaler - the return shape of preprocessing.py:13-75."""
    from sklearn.model_selection import train_test_split

    x, y = stack_features(cfg, mode, matrices)
    scaler = Standardizer().fit(x)
    xs = scaler.transform(x)
    x_train, x_test, y_train, y_test = train_test_split(
        xs, y, test_size=cfg.training.test_size, random_state=cfg.training.random_state, stratify=y)
    return x_train, x_test, y_train, y_test, scaler


# ------------------------------------------------------------------------------------------
# device-resident form: fed straight from ops.extract_features, no .mat round trip
# ------------------------------------------------------------------------------------------
def stack_features_device(cfg: Config, feats: dict, mode: str = "training"):
    """feats: {MOD: CUDA float64/float32 tensor (n_snr, n_frames, 18)} -> (x float32 (n, n_used), y int64),
    same row order as `stack_features` (modulation, then SNR of the chosen axis, then frame)."""
    import torch

    s, t, f = cfg.signals, cfg.training, cfg.features
    snr_axis = list(t.training_snr if mode == "training" else t.all_snr)
    cols = list(f.used)
    xs, ys = [], []
    for mod_idx, mod in enumerate(s.modulations_with_noise):
        m = feats[mod][snr_axis][:, : s.num_frames][:, :, cols]          # (n_sel_snr, n_frames, n_used)
        xs.append(m.reshape(-1, len(cols)).to(torch.float32))            # float32 like the reference's .mat
        ys.append(torch.full((m.shape[0] * m.shape[1],), s.labels[mod_idx], dtype=torch.int64, device=m.device))
    return torch.cat(xs), torch.cat(ys)


def standardize_device(x):
    """(x - mean) / std with StandardScaler's conventions, on the device; returns (x_std, mean, scale)."""
    import torch

    x64 = x.to(torch.float64)
    mean = x64.mean(dim=0)
    var = x64.var(dim=0, unbiased=False)
    scale = torch.where(var > 0, var.sqrt(), torch.ones_like(var))
    return ((x64 - mean) / scale).to(torch.float32), mean, scale


def split_indices(y, test_size: float, random_state: int):
    """(train_idx, test_idx): EXACTLY the partition `train_test_split(..., test_size, random_state, stratify=y)` applies
    (preprocessing.py:65-71) - sklearn's own StratifiedShuffleSplit run on the integer labels on the host (its
    Mersenne-Twister stream is part of the reference's result; the feature rows never leave the device)."""
    from sklearn.model_selection import train_test_split

    y = np.asarray(y)
    tr, te = train_test_split(np.arange(y.shape[0]), test_size=test_size, random_state=random_state, stratify=y)
    return tr, te


def load_feature_set_device(cfg: Config, feats: dict, mode: str = "training"):
    """Device-resident `preprocess_data` (preprocessing.py:13-75 with the (6, N) -> (N, 6) transpose of :55 fixed):
    feats {MOD: CUDA tensor (n_snr, n_frames, 18)} -> x_train, x_test, y_train, y_test (CUDA tensors, float32 / int64)
    and a `Standardizer` holding the fitted mean_ / scale_ (numpy) for later evaluation.  Same values and the same
    partition as the host consumer `load_feature_set`."""
    import torch

    x, y = stack_features_device(cfg, feats, mode)
    xs, mean, scale = standardize_device(x)
    tr, te = split_indices(y.cpu().numpy(), cfg.training.test_size, cfg.training.random_state)
    tr_d, te_d = torch.as_tensor(tr, device=x.device), torch.as_tensor(te, device=x.device)
    scaler = Standardizer()
    scaler.mean_, scaler.scale_ = mean.cpu().numpy(), scale.cpu().numpy()
    return xs[tr_d], xs[te_d], y[tr_d], y[te_d], scaler


def stratified_split_device(x, y, test_size: float, seed: int):
    """Per-class random split (same class proportions in both parts), all on the device, with torch's generator.
    NOT the reference's partition - `load_feature_set_device` uses `split_indices` for that; this one is for callers
    that only need a stratified split and want no host round trip at all."""
    import torch

    g = torch.Generator(device=x.device).manual_seed(seed)
    tr, te = [], []
    for c in torch.unique(y).tolist():
        idx = torch.nonzero(y == c, as_tuple=False).flatten()
        idx = idx[torch.randperm(idx.numel(), device=x.device, generator=g)]
        k = int(round(idx.numel() * test_size))
        te.append(idx[:k])
        tr.append(idx[k:])
    tr, te = torch.cat(tr), torch.cat(te)
    return x[tr], x[te], y[tr], y[te]
