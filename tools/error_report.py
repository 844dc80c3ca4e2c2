"""Per-feature worst relative error of the CUDA kernels against the golden fixtures of the reference (a report for
humans; the pass/fail version of it is tests/test_gpu_parity.py).  usage: python tools/error_report.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from conftest import golden_frames, golden_generic  # noqa: E402

from amcpy_b200 import ops  # noqa: E402

np.set_printoptions(linewidth=200, precision=3)


def report(tag, got, want):
    got = got.reshape(-1, 18)
    want = want.reshape(-1, 18)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
    print(tag, "worst rel err per feature:")
    print("  ", " ".join(f"{e:.1e}" for e in np.nanmax(err, axis=0)))
    if not np.isfinite(got).all():
        print("   non-finite outputs:", np.argwhere(~np.isfinite(got))[:10].tolist())


for n in (2048, 256, 1024, 4096):
    x, want = golden_frames(n)
    xd = torch.from_numpy(x).cuda()
    try:
        report(f"fused n={n}", ops.extract_features(xd).cpu().numpy(), want)
    except Exception as e:  # noqa: BLE001
        print(f"fused n={n} FAILED: {e}")
    try:
        report(f"general n={n}", ops.extract_features(xd, force_general=True).cpu().numpy(), want)
    except Exception as e:  # noqa: BLE001
        print(f"general n={n} FAILED: {e}")
for n in (10, 31, 100, 1000, 3000, 512, 8192, 16384):
    x, want = golden_generic(n)
    try:
        report(f"auto n={n}", ops.extract_features(torch.from_numpy(x).cuda()).cpu().numpy(), want)
    except Exception as e:  # noqa: BLE001
        print(f"auto n={n} FAILED: {e}")
