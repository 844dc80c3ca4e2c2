"""Small invocation of every fused kernel for compute-sanitizer (one tool per gpurun call)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from amcpy_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
for n, frames in ((2048, 700), (256, 3000), (512, 900), (1024, 500), (4096, 300), (8192, 200), (16384, 160)):
    x = torch.randn((frames, n), dtype=torch.complex128, device="cuda", generator=g)
    a = ops.extract_features(x)
    torch.cuda.synchronize()
    print(n, bool(torch.isfinite(a).all()))
x = torch.randn((300, 2048), dtype=torch.complex64, device="cuda", generator=g)
print("c64", bool(torch.isfinite(ops.extract_features(x)).all()))
# reduced feature profiles of the 16-samples-per-thread kernel (different barrier / mbarrier placement without the FFT)
for ids in ([10, 18], [6, 12], [2, 4, 6, 8, 12, 14]):
    for n, frames in ((2048, 700), (512, 900), (4096, 300)):
        x = torch.randn((frames, n), dtype=torch.complex128, device="cuda", generator=g)
        a = ops.extract_features(x, feature_mask=ops.feature_mask_of(ids))
        torch.cuda.synchronize()
        print(ids, n, bool(torch.isfinite(a[:, [i - 1 for i in ids]]).all()))
# round 2: the careful path (tagged rows recomputed by the general kernel in redo mode), long frames through the global FFT
# workspace, Bluestein beyond the shared-memory limit, the host pipeline (pipe pool) and the re-layout kernels
import numpy as np  # noqa: E402

for n in (256, 2048, 8192):
    x = torch.randn((64, n), dtype=torch.complex128, device="cuda", generator=g)
    x[::4] = 1.0 + 1e-3 * x[::4]                 # narrow phase clusters -> careful path
    x[1::8] *= 1e-30                             # out of the float32 range -> careful path
    a = ops.extract_features(x)
    torch.cuda.synchronize()
    print("careful", n, bool(torch.isfinite(a).all()))
for n, frames in ((32768, 6), (12000, 6)):
    x = torch.randn((frames, n), dtype=torch.complex128, device="cuda", generator=g)
    print("long", n, bool(torch.isfinite(ops.extract_features(x)).all()))
xh = (np.random.default_rng(0).standard_normal((300, 2048)) + 1j * np.random.default_rng(1).standard_normal((300, 2048)))
print("host", bool(np.isfinite(ops.extract_features_host(xh)).all()),
      bool(np.isfinite(ops.extract_features_host(np.asfortranarray(xh))).all()))
