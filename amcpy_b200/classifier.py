"""Downstream classifier on the extracted features - stock PyTorch, NOT part of
the hot path.  Same recipe as the reference so that `amcpy full` completes on the GPU features:
MLP n_used -> 26 -> 29 -> 30 -> n_classes, BatchNorm + activation + Dropout after every hidden
Linear, Softmax output fed to CrossEntropyLoss (sic), RMSprop lr 1.418e-3, batch 128, 21 epochs
(/root/reference/src/amcpy/nn_model.py:28-75, :88-198; config.py:151-171), per-SNR accuracy
(nn_model.py:227-267), checkpoint `ann/model-<id>.pt` in the reference's format (nn_model.py:175-185, :201-219:
state-dict keys `layers.N.*`, loadable by the reference's `load_model` and vice versa)."""

from __future__ import annotations

import uuid

import numpy as np
import torch.nn as nn


class AMCClassifier(nn.Module):
    """The reference's classifier (nn_model.py:28-75).  The Sequential lives under the attribute `layers`, so the
    state-dict keys are `layers.0.weight` ... `layers.12.bias` - the layout `load_model` (nn_model.py:201-219) expects."""

    def __init__(self, n_features: int, n_classes: int, hl1: int = 26, hl2: int = 29, hl3: int = 30,
                 dropout: float = 0.4, activation: str = "relu") -> None:
        super().__init__()
        act = {"relu": nn.ReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}.get(activation, nn.ReLU)
        mods, width = [], n_features
        for h in (hl1, hl2, hl3):
            mods += [nn.Linear(width, h), nn.BatchNorm1d(h), act(), nn.Dropout(dropout)]
            width = h
        mods += [nn.Linear(width, n_classes), nn.Softmax(dim=1)]
        self.layers = nn.Sequential(*mods)

    def forward(self, x):
        return self.layers(x)


def build_model(n_features: int, n_classes: int, hidden=(26, 29, 30), dropout: float = 0.4, activation: str = "relu"):
    return AMCClassifier(n_features, n_classes, *hidden, dropout=dropout, activation=activation)


def model_for(cfg):
    t = cfg.training
    return build_model(cfg.features.num_used, len(cfg.signals.modulations_with_noise),
                       (t.layer_size_hl1, t.layer_size_hl2, t.layer_size_hl3), t.dropout, t.activation)


def save_model(model, model_id: str, cfg):
    """`ann/model-<id>.pt` with the reference's keys (nn_model.py:175-185): model_state_dict / model_id / config (the
    Config object itself, as the reference stores its own)."""
    import torch

    cfg.paths.trained_ann.mkdir(parents=True, exist_ok=True)
    path = cfg.paths.trained_ann / f"model-{model_id}.pt"
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    torch.save({"model_state_dict": state, "model_id": model_id, "config": cfg}, path)
    return path


def load_model(model_id: str, cfg):
    """nn_model.py:201-219: rebuild the architecture from cfg, load `model_state_dict`, eval mode.  Reads checkpoints
    written here and by the reference alike (same keys)."""
    import torch

    path = cfg.paths.trained_ann / f"model-{model_id}.pt"
    checkpoint = torch.load(path, map_location="cpu", weights_only=False)
    model = model_for(cfg)
    model.load_state_dict(checkpoint["model_state_dict"])
    model.eval()
    return model


def train_classifier(cfg, x_train, y_train, x_test, y_test, device=None, epochs: int | None = None, seed: int = 0,
                     verbose: bool = True):
    """Returns (model, model_id, history).  All tensors stay on `device` for the whole run."""
    import torch
    import torch.nn as nn

    t = cfg.training
    device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
    torch.manual_seed(seed)
    n_classes = len(cfg.signals.modulations_with_noise)
    model = build_model(x_train.shape[1], n_classes, (t.layer_size_hl1, t.layer_size_hl2, t.layer_size_hl3),
                        t.dropout, t.activation).to(device)
    opt_cls = {"rmsprop": torch.optim.RMSprop, "adam": torch.optim.Adam}.get(t.optimizer, torch.optim.NAdam)
    opt = opt_cls(model.parameters(), lr=t.learning_rate)
    loss_fn = nn.CrossEntropyLoss()
    def dev(a, dt):   # device tensors from the device consumer pass through; host arrays are uploaded once
        if isinstance(a, torch.Tensor):
            return a.to(device=device, dtype=dt)
        return torch.as_tensor(np.asarray(a), dtype=dt, device=device)

    xt, yt = dev(x_train, torch.float32), dev(y_train, torch.long)
    xv, yv = dev(x_test, torch.float32), dev(y_test, torch.long)
    hist = {"loss": [], "accuracy": [], "val_loss": [], "val_accuracy": []}
    n = xt.shape[0]
    for ep in range(epochs if epochs is not None else t.epochs):
        model.train()
        order = torch.randperm(n, device=device)
        tot_loss = torch.zeros((), device=device)
        hits = torch.zeros((), device=device)
        for i in range(0, n, t.batch_size):
            idx = order[i:i + t.batch_size]
            if idx.numel() < 2:      # BatchNorm needs more than one sample
                continue
            out = model(xt[idx])
            loss = loss_fn(out, yt[idx])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            tot_loss += loss.detach() * idx.numel()
            hits += (out.argmax(1) == yt[idx]).sum()
        model.eval()
        with torch.no_grad():
            pv = model(xv)
            hist["val_loss"].append(float(loss_fn(pv, yv)))
            hist["val_accuracy"].append(float((pv.argmax(1) == yv).float().mean()))
        hist["loss"].append(float(tot_loss) / n)
        hist["accuracy"].append(float(hits) / n)
        if verbose:
            print(f"Epoch {ep + 1:3d} | loss {hist['loss'][-1]:.4f} | acc {hist['accuracy'][-1]:.4f} | "
                  f"val_loss {hist['val_loss'][-1]:.4f} | val_acc {hist['val_accuracy'][-1]:.4f}")
    model_id = uuid.uuid4().hex[:8]
    save_model(model, model_id, cfg)
    return model, model_id, hist


def accuracy_by_snr(model, scaler, cfg, matrices: dict | None = None, device=None) -> dict:
    """{snr index -> accuracy over all modulations} with the selected columns standardised by the
    training scaler (nn_model.py:227-267, columns as in graphics.py:46)."""
    import scipy.io
    import torch

    device = device or next(model.parameters()).device
    s, f = cfg.signals, cfg.features
    cols = list(f.used)
    model.eval()
    res = {}
    mats = {}
    for mod in s.modulations_with_noise:
        mats[mod] = (np.asarray(matrices[mod]) if matrices is not None else
                     scipy.io.loadmat(str(cfg.paths.calculated_features / f"{mod}_features.mat"))[s.mat_info[mod]])
    with torch.no_grad():
        for snr in cfg.training.all_snr:
            ok = tot = 0
            for label, mod in zip(s.labels, s.modulations_with_noise):
                x = scaler.transform(np.asarray(mats[mod][snr, : s.num_frames][:, cols], dtype=np.float32))
                pred = model(torch.as_tensor(x, dtype=torch.float32, device=device)).argmax(1)
                ok += int((pred == label).sum())
                tot += x.shape[0]
            res[snr] = ok / max(tot, 1)
    return res
