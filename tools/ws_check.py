"""A/B of the warp-specialised N=2048 kernel (ops.extract_features(..., ws=True)) against the default kernel:
bitwise comparison on BASELINE config 2 + timing.  usage: python tools/ws_check.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from amcpy_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda")
snrs = [-10.0 + 2.0 * i for i in range(16)]
for nf in (1, 3, 700, 48000):
    x = synth.dataset_device(6, snrs, 500, 2048, dev, seed=5)[:nf].contiguous()
    a = ops.extract_features(x)
    b = ops.extract_features(x, extra_flags=4)
    torch.cuda.synchronize()
    print(nf, "bitwise equal:", bool(torch.equal(a, b)), "max abs diff", float((a - b).abs().nan_to_num().max()))
x = synth.dataset_device(6, snrs, 500, 2048, dev, seed=5)
out = torch.empty((x.shape[0], 18), dtype=torch.float64, device=dev)
for ws in (False, True, False, True):
    for _ in range(5):
        ops.extract_features(x, out=out, extra_flags=4 if ws else 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        ops.extract_features(x, out=out, extra_flags=4 if ws else 0)
    e1.record()
    torch.cuda.synchronize()
    print("ws" if ws else "default", round(e0.elapsed_time(e1) / 100, 4), "ms")
