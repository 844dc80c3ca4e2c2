"""Command line for the extraction path: `python -m amcpy_b200.main {extract,full,synth}`.

Mirrors the reference's console script for the sub-commands that reach the hot path
(/root/reference/src/amcpy/main.py:31-32 `extract`, :62-63 `full`, dispatch :160-175).  Unlike the
reference - whose dispatcher passes (cfg, args) to the one-argument `cmd_extract` and raises
TypeError (main.py:85 vs :175) - `extract` works.  `full` runs the extraction, the feature
consumer (column select + standardise + stratified split) and the stock-PyTorch classifier
(train + per-SNR accuracy); plotting is outside the path.
`synth` writes a synthetic mat-data/all_modulations.mat the reference can consume as well.
"""

from __future__ import annotations

import argparse
from pathlib import Path

from .config import Config, Paths, SignalConfig
from .feature_extraction import extract_all, run_extraction


def _build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="amcpy", description="AMCPy feature extraction on B200")
    p.add_argument("--root", type=Path, default=None, help="project root (default: current directory)")
    p.add_argument("--num-frames", type=int, default=None, help="frames per (modulation, SNR) (default 1000)")
    p.add_argument("--frame-size", type=int, default=None, help="samples per frame (default 2048)")
    sub = p.add_subparsers(dest="command", required=True)
    sub.add_parser("extract", help="Extract features from raw .mat data")
    fp = sub.add_parser("full", help="extract -> standardise/split -> train/evaluate the classifier")
    fp.add_argument("--epochs", type=int, default=None)
    s = sub.add_parser("synth", help="write a synthetic mat-data/all_modulations.mat")
    s.add_argument("--seed", type=int, default=2024)
    return p


def _config(args) -> Config:
    sig = {}
    if args.num_frames is not None:
        sig["num_frames"] = args.num_frames
    if args.frame_size is not None:
        sig["frame_size"] = args.frame_size
    return Config(paths=Paths(root=args.root) if args.root else Paths(), signals=SignalConfig(**sig))


def _init_distributed_if_launched() -> int:
    """Under torchrun (WORLD_SIZE > 1) join the process group so that run_extraction's closing barrier works
    (every rank extracts a contiguous shard of the flattened (modulation, frame, snr) space; rank 0 assembles and writes
    the files).  Host-side exchange of the small feature blocks only: gloo."""
    import os

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist

        if dist.is_available() and not dist.is_initialized():
            dist.init_process_group(backend="gloo")
    return int(os.environ.get("RANK", "0"))


def cmd_extract(cfg: Config, args=None) -> None:
    _init_distributed_if_launched()
    run_extraction(cfg)


def cmd_synth(cfg: Config, args) -> None:
    from . import synth

    cfg.paths.ensure_dirs()
    snrs = [float(v) for v in cfg.signals.snr_values.values()]
    data = synth.dataset(snrs, cfg.signals.num_frames, cfg.signals.frame_size, args.seed)
    out = cfg.paths.mat_data / cfg.paths.mat_filename
    synth.write_all_modulations_mat(out, data, cfg.signals.mat_info)
    print(f"wrote {out}: 6 x {data.shape[1]} x {data.shape[2]} x {data.shape[3]} complex128")


def cmd_full(cfg: Config, args=None) -> None:
    """extract -> feature consumer -> classifier, the matrices handed over in memory (main.py:153-157 chains the
    stages through calculated-features/*.mat; the files are still written - they are the stage's contract)."""
    import torch

    from .classifier import accuracy_by_snr, train_classifier
    from .consumer import load_feature_set_device

    rank = _init_distributed_if_launched()
    mats = extract_all(cfg)
    if rank != 0:
        return                                   # the classifier is a single-GPU job: rank 0 trains
    dev = torch.device("cuda", int(__import__("os").environ.get("LOCAL_RANK", "0")) % max(1, torch.cuda.device_count()))
    feats = {m: torch.from_numpy(v).to(dev) for m, v in mats.items()}    # float32 (n_snr, n_frames, 18), as in the .mat
    x_train, x_test, y_train, y_test, scaler = load_feature_set_device(cfg, feats, mode="training")
    print(f"feature set ready: train {tuple(x_train.shape)}, test {tuple(x_test.shape)}, "
          f"{int(torch.unique(y_train).numel())} classes")
    model, model_id, hist = train_classifier(cfg, x_train, y_train, x_test, y_test, device=dev,
                                             epochs=getattr(args, "epochs", None))
    acc = accuracy_by_snr(model, scaler, cfg, matrices=mats)
    print(f"model {model_id}: val_acc {hist['val_accuracy'][-1]:.4f}; accuracy by SNR index: "
          + ", ".join(f"{k}:{v:.2f}" for k, v in acc.items()))


def main(argv=None) -> None:
    args = _build_parser().parse_args(argv)
    cfg = _config(args)
    cfg.paths.ensure_dirs()
    try:
        {"extract": cmd_extract, "full": cmd_full, "synth": cmd_synth}[args.command](cfg, args)
    finally:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
