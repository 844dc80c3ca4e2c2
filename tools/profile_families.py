"""One launch of every shipped kernel family inside a cudaProfilerStart/Stop window, for
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_families python tools/profile_families.py
(plain run first: prints one line per family with its CUDA-event time).  Families: warp-per-frame (N = 256), fused16 at
512 / 1024 / 2048 / 4096, long-frame at 8192 / 16384, general kernel (power of two in shared memory, Bluestein, long
frame through the global workspace), the two re-layout kernels and the generator."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from amcpy_b200 import _native as nat  # noqa: E402
from amcpy_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda")
SNRS = [-10.0 + 2.0 * i for i in range(16)]
total = 6 * 16 * 500 * 2048


def timed(name, fn, alg_bytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(json.dumps({"family": name, "ms": round(ms, 4), "algorithmic_bytes": alg_bytes,
                      "GBps": round(alg_bytes / (ms * 1e-3) / 1e9, 1)}), flush=True)


for n in (256, 512, 1024, 2048, 4096, 8192, 16384):
    per = total // n // 96
    x = synth.dataset_device(6, SNRS, per, n, dev, seed=5)
    out = torch.empty((x.shape[0], 18), dtype=torch.float64, device=dev)
    timed(f"fused N={n} c128", lambda: ops.extract_features(x, out=out), x.shape[0] * (n * 16 + 144))
    del x, out
# general kernel: power of two (shared-memory FFT), Bluestein, long frame (global FFT workspace)
for n, frames in ((2048, 4000), (3000, 3000), (32768, 300)):
    x = synth.dataset_device(6, SNRS[:1], frames // 6, n, dev, seed=6)
    out = torch.empty((x.shape[0], 18), dtype=torch.float64, device=dev)
    timed(f"general N={n} c128", lambda: ops.extract_features(x, out=out, force_general=True), x.shape[0] * (n * 16 + 144))
    del x, out
# re-layout kernels (sample-major / planar -> one row per frame) and the generator
nf, n = 24000, 2048
src = synth.dataset_device(6, SNRS, nf // 96, n, dev, seed=7)
sm = src.t().contiguous()                                   # element (f, k) at f + k*nf
timed("frames_from_sample_major N=2048 c128", lambda: ops.frames_from_sample_major(sm, nf, n, nf), 2 * nf * n * 16)
del sm
xh = src[:6144].cpu().numpy()
re_p = np.ascontiguousarray(xh.real.T).reshape(-1)
im_p = np.ascontiguousarray(xh.imag.T).reshape(-1)
timed("extract_host_planar (frames_from_planar + fused) 6144 frames", lambda: ops.extract_features_host_planar(
    re_p, im_p, 6144, n, 6144), 6144 * (n * 16 + 144))
cell_mod = torch.tensor([m for m in range(6) for _ in SNRS], dtype=torch.int32, device=dev)
cell_snr = torch.tensor([s for _ in range(6) for s in range(16)], dtype=torch.int32, device=dev)
cell_sig = torch.tensor([float(np.sqrt(10.0 ** (-s / 10.0) / 2.0)) for _ in range(6) for s in SNRS], dtype=torch.float64, device=dev)
gen_out = torch.empty((96 * 250, n), dtype=torch.complex128, device=dev)
timed("generate_frames 24000 x 2048 c128", lambda: nat.check(nat.lib().amc_generate_frames(
    gen_out.data_ptr(), nat.AMC_C128, 96, 250, 0, n, cell_mod.data_ptr(), cell_snr.data_ptr(), cell_sig.data_ptr(), 11,
    torch.cuda.current_stream().cuda_stream)), 24000 * n * 16)
