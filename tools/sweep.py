"""Frame-size sweep (BASELINE config 4): device-resident throughput of the extraction kernels for
N in {256, 512, 1024, 2048, 4096, 8192, 16384}, same total sample count (~98 M complex128 samples).
usage: python tools/sweep.py [--steps 20] [--features 10 11 ...]   -> one JSON line per N
--features: ids the caller wants (feature_mask of the C ABI): the library may run a reduced feature profile."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from amcpy_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--sizes", type=int, nargs="*", default=[256, 512, 1024, 2048, 4096, 8192, 16384])
ap.add_argument("--dtype", default="c128")
ap.add_argument("--features", type=int, nargs="*", default=list(range(1, 19)))
args = ap.parse_args()
peak = 6454.0
try:
    peak = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass
total = 6 * 16 * 500 * 2048
mask = ops.feature_mask_of(args.features)
dt = torch.complex128 if args.dtype == "c128" else torch.complex64
g = torch.Generator(device="cuda").manual_seed(1)
for n in args.sizes:
    frames = total // n
    x = torch.randn((frames, n), dtype=dt, device="cuda", generator=g)
    out = torch.empty((frames, 18), dtype=torch.float64, device="cuda")
    for _ in range(3):
        ops.extract_features(x, out=out, feature_mask=mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ops.extract_features(x, out=out, feature_mask=mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    bytes_ = frames * (n * x.element_size() + 144)
    gbs = bytes_ / (ms * 1e-3) / 1e9
    print(json.dumps({"frame_size": n, "dtype": args.dtype, "features": args.features, "frames": frames, "ms": round(ms, 4),
                      "frames_per_s": round(frames / (ms * 1e-3)), "samples_per_s": round(frames * n / (ms * 1e-3)),
                      "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4)}), flush=True)
    del x, out
