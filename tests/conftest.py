import hashlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
SNRS = [-10.0 + 2.0 * i for i in range(16)]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible and -m gpu was not asked for."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _library_present():
    """A checkout without build products (the .so is git-ignored) builds the CUDA library once; an existing library
    is used as it is.  Nothing here makes a CPU path available: without the library every op raises."""
    import shutil

    from amcpy_b200 import _native as nat

    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        nat.build_if_missing()


def run_c_example(tmp_path, n_frames=64):
    """Compile examples/extract_host.c with gcc against the built library and run it; returns CompletedProcess."""
    import shutil
    import subprocess

    from amcpy_b200 import _native as nat

    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        pytest.skip("no C compiler")
    exe = tmp_path / "extract_host"
    lib_dir = nat.LIB_PATH.parent
    cmd = [gcc, "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "extract_host.c"), "-o", str(exe),
           f"-L{lib_dir}", "-lamcpy_b200", "-lm", f"-Wl,-rpath,{lib_dir}"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    return subprocess.run([str(exe), str(n_frames)], capture_output=True, text=True, timeout=300)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_golden(name):
    return np.load(GOLDEN / name, allow_pickle=False)


def golden_frames(n):
    """Regenerate the inputs of tests/golden/frames_n{n}.npz; returns (x[6,4,3,n], features)."""
    from amcpy_b200 import synth

    g = load_golden(f"frames_n{n}.npz")
    seed = int(g["seed"])
    x = np.stack([synth.cell(m, SNRS[si], si, range(3), n, seed) for m in range(6) for si in g["snr_idx"]])
    x = x.reshape(6, len(g["snr_idx"]), 3, n)
    assert sha(x) == str(g["input_sha256"]), "synthetic generator drifted from the golden inputs"
    return x, g["features"]


def golden_generic(n):
    from amcpy_b200 import synth

    g = load_golden(f"generic_n{n}.npz")
    x = np.stack([synth.frame(m, SNRS[8], 8, 0, n, int(g["seed"])) for m in range(6)])
    assert sha(x) == str(g["input_sha256"])
    return x, g["features"]


def golden_hard(n):
    """Inputs (regenerated, checked by hash) and reference outputs of tests/golden/hard_n{n}.npz:
    (x[10, n], features[10, 18], features_c64[4, 18]); rows 8 and 9 hold a NaN."""
    from oracle.hard_cases import hard_case_inputs

    g = load_golden(f"hard_n{n}.npz")
    x = hard_case_inputs(n)
    assert sha(x) == str(g["input_sha256"]), "hard-case inputs drifted from the golden fixture"
    return x, g["features"], g["features_c64"]


def golden_narrow(n):
    """Inputs (regenerated, checked by hash) and reference outputs of tests/golden/narrow_n{n}.npz: (x[10, n], features[10, 18])."""
    from oracle.hard_cases import narrow_case_inputs

    g = load_golden(f"narrow_n{n}.npz")
    x = narrow_case_inputs(n)
    assert sha(x) == str(g["input_sha256"]), "narrow-case inputs drifted from the golden fixture"
    return x, g["features"]


def assert_features_close(got, want, *, scale=1.0, ids=range(1, 19)):
    """Per-feature relative tolerance classes of BASELINE.json north_star (written here):
    1e-6 on FFT/atan2-derived features (1,2,3,5,9), 1e-9 on all others."""
    got = np.asarray(got, dtype=np.float64).reshape(-1, 18)
    want = np.asarray(want, dtype=np.float64).reshape(-1, 18)
    assert got.shape == want.shape
    for fid in ids:
        rtol = (1e-6 if fid in (1, 2, 3, 5, 9) else 1e-9) * scale
        g, w = got[:, fid - 1], want[:, fid - 1]
        both_nan = np.isnan(g) & np.isnan(w)
        err = np.abs(g - w) / np.maximum(np.abs(w), 1e-300)
        err[both_nan] = 0.0
        bad = ~(err <= rtol)
        assert not bad.any(), (
            f"feature {fid}: {bad.sum()} of {g.size} frames outside rtol={rtol:g}; "
            f"worst rel err {np.nanmax(err):.3e} at frame {int(np.nanargmax(err))}"
        )
