"""Randomised parity sweep: frames whose scale, DC offset, carrier offset, SNR and modulation are drawn at
random (far outside the benchmark's tidy unit-power grid) through every fused kernel, against the oracle at the
north_star tolerances.  Seeded, ~1 s of GPU + a few seconds of oracle per size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_frames(rng, n_frames, n):
    from amcpy_b200 import synth

    out = np.empty((n_frames, n), dtype=np.complex128)
    t = np.arange(n)
    for i in range(n_frames):
        mod = int(rng.integers(0, 6))
        snr = float(rng.uniform(-12.0, 32.0))
        x = synth.frame(mod, snr, int(rng.integers(0, 16)), int(rng.integers(0, 1 << 20)), n, seed=int(rng.integers(1, 1 << 30)))
        scale = 10.0 ** rng.uniform(-3.0, 3.0)                       # 120 dB of dynamic range between frames
        cfo = rng.uniform(-0.2, 0.2) if rng.random() < 0.5 else 0.0  # carrier offset in cycles/sample
        dc = (rng.standard_normal() + 1j * rng.standard_normal()) * (0.3 if rng.random() < 0.3 else 0.0)
        ph0 = np.exp(2j * np.pi * rng.random())
        out[i] = scale * ((x + dc) * ph0 * np.exp(2j * np.pi * cfo * t))
    return out


@pytest.mark.parametrize("n,frames", [(256, 96), (512, 64), (1024, 64), (2048, 96), (4096, 32), (8192, 16), (16384, 8)])
def test_random_scale_offset_cfo_frames_against_oracle(n, frames):
    import torch

    from amcpy_b200 import ops
    from oracle import amc_oracle as orc

    rng = np.random.default_rng(1000 + n)
    x = _random_frames(rng, frames, n)
    want = orc.features_batch(x)
    got = ops.extract_features(torch.from_numpy(x).cuda()).cpu().numpy()
    worst = {}
    for fid in range(1, 19):
        w, g = want[:, fid - 1], got[:, fid - 1]
        rel = np.abs(g - w) / np.abs(w)
        worst[fid] = float(np.nanmax(rel))
        assert np.array_equal(np.isnan(w), np.isnan(g)), f"feature {fid}: NaN pattern differs"
        assert worst[fid] <= orc.RTOL[fid], f"N={n} feature {fid}: rel err {worst[fid]:.3e} > {orc.RTOL[fid]:g}"


@pytest.mark.parametrize("n,frames", [(256, 48), (2048, 48), (8192, 8)])
def test_random_frames_complex64_input_equals_widened_complex128(n, frames):
    """complex64 input is widened exactly on the device and then runs the same arithmetic: bitwise the result of
    the complex128 path on the widened data, hence within the classes of the oracle on that data."""
    import torch

    from amcpy_b200 import ops
    from oracle import amc_oracle as orc

    rng = np.random.default_rng(2000 + n)
    x32 = _random_frames(rng, frames, n).astype(np.complex64)
    got32 = ops.extract_features(torch.from_numpy(x32).cuda())
    got64 = ops.extract_features(torch.from_numpy(x32.astype(np.complex128)).cuda())
    assert torch.equal(got32, got64)
    want = orc.features_batch(x32.astype(np.complex128))
    g = got32.cpu().numpy()
    for fid in range(1, 19):
        rel = np.nanmax(np.abs(g[:, fid - 1] - want[:, fid - 1]) / np.abs(want[:, fid - 1]))
        assert rel <= orc.RTOL[fid], f"N={n} feature {fid}: rel err {rel:.3e}"


def test_randomised_soak_prefix():
    """The first 150 cases of tests/soak.py, seed 1 (a deterministic sequence): random frame sizes incl. non-powers of
    two, batch sizes, dtypes, scales 1e-4..1e4, carrier / DC offsets, layouts (contiguous, padded, sample-major, host
    pipeline) and feature masks against the oracle."""
    import json
    import subprocess
    import sys

    from conftest import ROOT

    res = subprocess.run([sys.executable, str(ROOT / "tests" / "soak.py"), "--cases", "150", "--seed", "1"],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    summary = json.loads(res.stdout.strip().splitlines()[-1])
    assert summary["soak"] == "ok" and summary["cases"] == 150
