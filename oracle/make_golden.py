"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Run:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py
Needs /root/reference/src (read-only, never present on the GPU box); the resulting small
fixtures are committed so tests can pin both the oracle and the CUDA path without it.
Inputs are NOT stored (they are regenerated from amcpy_b200.synth with the recorded seed);
a sha256 of each input block is stored so generator drift is detected.
"""

from __future__ import annotations

import hashlib
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True

from amcpy.config import Config, Paths, SignalConfig  # noqa: E402  (the reference)
from amcpy.feature_extraction import run_extraction  # noqa: E402
from amcpy.features import (  # noqa: E402
    InstantaneousValues,
    MomentValues,
    _test_signal,
    calculate_features,
)

from amcpy_b200 import synth  # noqa: E402
from oracle.hard_cases import HARD_SIZES, NARROW_SIZES, hard_case_inputs, narrow_case_inputs  # noqa: E402

GOLD = ROOT / "tests" / "golden"
SEED = 2024
IDS = list(range(1, 19))
SNRS = [-10.0 + 2.0 * i for i in range(16)]  # reference config.py:75-94


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_features(frames: np.ndarray) -> np.ndarray:
    with np.errstate(all="ignore"):
        return np.array([calculate_features(IDS, f) for f in frames], dtype=np.float64)


def hard_cases() -> None:
    """Outputs of the unmodified reference on oracle/hard_cases.py: non-power-of-two sizes (pocketfft takes any
    length), carrier offset / DC / scale, complex64 input (numpy computes it in float32), frames holding a NaN."""
    for n in HARD_SIZES:
        x = hard_case_inputs(n)
        x64 = x[:4].astype(np.complex64)
        np.savez(GOLD / f"hard_n{n}.npz", n=n, input_sha256=sha(x), features=ref_features(x),
                 features_c64=ref_features(x64))


def narrow_cases() -> None:
    """Outputs of the unmodified reference on oracle/hard_cases.py:narrow_case_inputs - the frames the fused kernels
    hand to the careful (float64) path: narrow phase clusters, extreme scales, degenerate amplitude distributions."""
    for n in NARROW_SIZES:
        x = narrow_case_inputs(n)
        np.savez(GOLD / f"narrow_n{n}.npz", n=n, input_sha256=sha(x), features=ref_features(x))


def main() -> None:
    GOLD.mkdir(parents=True, exist_ok=True)
    if "--narrow-only" in sys.argv:
        narrow_cases()
        return
    if "--long-only" in sys.argv:      # long frames: beyond the shared-memory FFT of the general kernel
        for n in (12000, 32768, 65536):
            x = np.stack([synth.frame(m, SNRS[8], 8, 0, n, SEED) for m in range(6)])
            np.savez(GOLD / f"generic_n{n}.npz", seed=SEED, snr_idx=8, n=n, features=ref_features(x), input_sha256=sha(x))
        return
    if "--hard-only" in sys.argv:      # leaves the other fixtures (and their zip timestamps) untouched
        hard_cases()
        return

    # 1. the reference's own 10-sample fixture: features + helper types
    sig = _test_signal()
    iv, mv = InstantaneousValues(sig), MomentValues(sig)
    np.savez(
        GOLD / "kat10.npz",
        signal=sig,
        features=np.array(calculate_features(IDS, sig)),
        **{f"iv_{k}": getattr(iv, k) for k in ("abs", "phase", "unwrapped_phase", "frequency", "cn_amplitude")},
        **{f"mv_{k}": np.asarray(getattr(mv, k)) for k in
           ("m20", "m21", "m22", "m40", "m41", "m42", "m43", "m60", "m61", "m62", "m63")},
    )

    # 2. realistic frames: 6 modulations x SNR idx {0,5,10,15} x 3 frames, several frame sizes
    snr_pick = [0, 5, 10, 15]
    for n in (256, 1024, 2048, 4096):
        blocks, feats = [], []
        for m in range(6):
            for si in snr_pick:
                fr = synth.cell(m, SNRS[si], si, range(3), n, SEED)
                blocks.append(fr)
                feats.append(ref_features(fr))
        x = np.stack(blocks).reshape(6, len(snr_pick), 3, n)
        np.savez(GOLD / f"frames_n{n}.npz", seed=SEED, snr_idx=np.array(snr_pick), n=n,
                 features=np.stack(feats).reshape(6, len(snr_pick), 3, 18), input_sha256=sha(x))

    # 3. ragged / generic sizes (the reference accepts any length): one frame per class
    for n in (10, 31, 100, 1000, 3000, 512, 8192, 16384, 12000, 32768, 65536):
        x = np.stack([synth.frame(m, SNRS[8], 8, 0, n, SEED) for m in range(6)])
        np.savez(GOLD / f"generic_n{n}.npz", seed=SEED, snr_idx=8, n=n,
                 features=ref_features(x), input_sha256=sha(x))

    # 4. helper types on one realistic frame each (N=256 keeps the fixture small)
    x = synth.frame(1, SNRS[10], 10, 0, 256, SEED)
    iv, mv = InstantaneousValues(x), MomentValues(x)
    np.savez(
        GOLD / "helpers_n256.npz", seed=SEED, input_sha256=sha(x),
        **{f"iv_{k}": getattr(iv, k) for k in ("abs", "phase", "unwrapped_phase", "frequency", "cn_amplitude")},
        **{f"mv_{k}": np.asarray(getattr(mv, k)) for k in
           ("m20", "m21", "m22", "m40", "m41", "m42", "m43", "m60", "m61", "m62", "m63")},
    )

    # 5. the stage: run_extraction on a tiny all_modulations.mat (16 SNR x 2 frames x 2048, with
    #    8 surplus samples per frame so the 0:frame_size truncation of feature_extraction.py:68 is
    #    exercised); keep the six float32 matrices + what loadmat returns for "Modulation".
    import scipy.io

    with tempfile.TemporaryDirectory() as td:
        cfg = Config(paths=Paths(root=Path(td)), signals=SignalConfig(num_frames=2, num_threads=1))
        cfg.paths.ensure_dirs()
        data = synth.dataset(SNRS, 2, 2048 + 8, SEED)
        synth.write_all_modulations_mat(cfg.paths.mat_data / cfg.paths.mat_filename, data, cfg.signals.mat_info)
        run_extraction(cfg)
        out = {"input_sha256": sha(data), "seed": SEED}
        for mod in cfg.signals.modulations_with_noise:
            m = scipy.io.loadmat(str(cfg.paths.calculated_features / f"{mod}_features.mat"))
            key = cfg.signals.mat_info[mod]
            assert m[key].dtype == np.float32 and m[key].shape == (16, 2, 18)
            out[f"{mod}_matrix"] = m[key]
            out[f"{mod}_modulation"] = m["Modulation"]
        np.savez(GOLD / "stage_16x2x2048.npz", **out)

    # 6. hard cases
    hard_cases()
    narrow_cases()

    for p in sorted(GOLD.glob("*.npz")):
        print(p.name, os.path.getsize(p))


if __name__ == "__main__":
    main()
