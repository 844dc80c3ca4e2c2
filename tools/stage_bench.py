"""End-to-end timing of the drop-in STAGE (BASELINE config 1): `run_extraction(cfg)` on a synthetic
mat-data/all_modulations.mat of 6 modulations x 16 SNRs x 500 frames x 2048 complex128 samples
(1.57 GB): loadmat + GPU extraction + savemat, wall clock.  usage: python tools/stage_bench.py [--frames 500]"""
import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import scipy.io  # noqa: E402
import torch  # noqa: E402

from amcpy_b200 import synth  # noqa: E402
from amcpy_b200.config import Config, Paths, SignalConfig  # noqa: E402
from amcpy_b200.feature_extraction import run_extraction  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=500)
args = ap.parse_args()
with tempfile.TemporaryDirectory() as td:
    cfg = Config(paths=Paths(root=Path(td)), signals=SignalConfig(num_frames=args.frames))
    cfg.paths.ensure_dirs()
    snrs = [float(v) for v in cfg.signals.snr_values.values()]
    t0 = time.perf_counter()
    x = synth.dataset_device(6, snrs, args.frames, 2048, torch.device("cuda"), seed=1).view(6, 16, args.frames, 2048)
    data = x.cpu().numpy()
    synth.write_all_modulations_mat(cfg.paths.mat_data / cfg.paths.mat_filename, data, cfg.signals.mat_info)
    t_write = time.perf_counter() - t0
    del x, data
    t0 = time.perf_counter()
    m = scipy.io.loadmat(str(cfg.paths.mat_data / cfg.paths.mat_filename))
    t_load = time.perf_counter() - t0
    del m
    run_extraction(cfg)          # warm-up (library load, staging buffers)
    t0 = time.perf_counter()
    run_extraction(cfg)
    t_stage = time.perf_counter() - t0
    out = scipy.io.loadmat(str(cfg.paths.calculated_features / "QPSK_features.mat"))["signal_qpsk"]
    frames = 6 * 16 * args.frames
    print(json.dumps({"frames": frames, "write_mat_s": round(t_write, 2), "loadmat_alone_s": round(t_load, 2),
                      "run_extraction_s": round(t_stage, 3), "frames_per_s": round(frames / t_stage),
                      "out_shape": list(out.shape), "out_dtype": str(out.dtype)}))
