"""CPU: host-side logic - the library loads and exports every symbol of include/amcpy_b200.h,
config parity with the reference's field names/defaults, layout planning of the stage, the
synthetic generator, and that no product module routes through the oracle."""

import ast
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="session")
def built_lib():
    from amcpy_b200 import _native as nat

    nat.build()
    return nat.lib()


def test_library_exports_every_declared_symbol(built_lib):
    from amcpy_b200 import _native as nat

    header = (ROOT / "include" / "amcpy_b200.h").read_text()
    declared = set(re.findall(r"\b(amc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in the header"
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    for name in declared:
        assert getattr(built_lib, name) is not None
    assert built_lib.amc_version() >= 1000
    assert built_lib.amc_last_error_string() is not None
    assert built_lib.amc_launch_count() == 0


def test_python_constants_mirror_the_header():
    """Every `#define AMC_*` of include/amcpy_b200.h has the same value in the ctypes binding, and the feature-mask
    helper maps ids to the header's bit convention (bit k = feature id k+1)."""
    from amcpy_b200 import _native as nat
    from amcpy_b200 import ops

    header = (ROOT / "include" / "amcpy_b200.h").read_text()
    defines = dict(re.findall(r"#define\s+(AMC_[A-Z0-9_]+)\s+\(?(-?(?:0x[0-9A-Fa-f]+|\d+))u?\)?", header))
    assert {"AMC_OK", "AMC_ERR_INVALID_ARG", "AMC_C128", "AMC_ALL_FEATURES", "AMC_FLAG_DIRECT_DFT"} <= set(defines)
    for name, text in defines.items():
        if hasattr(nat, name):
            assert getattr(nat, name) == int(text, 0), name
    for name in ("AMC_C64", "AMC_C128", "AMC_ALL_FEATURES", "AMC_FLAG_FORCE_GENERAL", "AMC_FLAG_FUSED_SPT8",
                 "AMC_FLAG_FUSED_WS", "AMC_FLAG_DIRECT_DFT"):
        assert hasattr(nat, name), name
    assert ops.feature_mask_of(range(1, 19)) == nat.AMC_ALL_FEATURES
    assert ops.feature_mask_of([1]) == 1 and ops.feature_mask_of([18]) == 1 << 17
    assert ops.feature_mask_of([2, 4, 6, 8, 12, 14]) == 0b10100010101010
    with pytest.raises(KeyError):
        ops.feature_mask_of([19])


def test_plain_c_host_links_against_the_abi(built_lib, tmp_path):
    """examples/extract_host.c (no Python, no torch) compiles with -Wall -Werror against include/amcpy_b200.h and the
    library; without a device it stops after the ABI checks that need no GPU and says so (no CPU compute path)."""
    import torch

    from conftest import run_c_example

    res = run_c_example(tmp_path, 8)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "ABI version" in res.stdout and "iq_dtype 99 unknown" in res.stdout
    if not torch.cuda.is_available():
        assert "no CUDA device" in res.stdout and "frame 0:" not in res.stdout


def test_no_cuda_device_is_an_error_not_a_fallback(built_lib):
    import torch

    from amcpy_b200 import _native as nat

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(nat.AmcError):
        nat.require_cuda()
    from amcpy_b200.config import Config, Paths
    from amcpy_b200.feature_extraction import run_extraction

    with pytest.raises(nat.AmcError):
        run_extraction(Config(paths=Paths(root=Path("/tmp/amcpy_b200_nonexistent"))))


def test_argument_checks_run_without_a_gpu(built_lib):
    from amcpy_b200 import _native as nat

    # validation happens before any CUDA call, so these are safe on a CPU-only box
    assert built_lib.amc_extract_batch(None, 5, 1, 16, 16, 1, None, 18, nat.AMC_ALL_FEATURES, 0, None) == -1
    assert b"iq_dtype" in built_lib.amc_last_error_string()
    assert built_lib.amc_extract_batch(None, nat.AMC_C128, 0, 16, 16, 1, None, 18, nat.AMC_ALL_FEATURES, 0, None) == 0
    assert built_lib.amc_extract_batch(None, nat.AMC_C128, 1, 16, 16, 1, None, 18, nat.AMC_ALL_FEATURES, 0, None) == -1
    assert built_lib.amc_extract_host(None, nat.AMC_C128, -1, 16, 16, 1, None, 18, nat.AMC_ALL_FEATURES, 0, 0) == -1
    # planar entry: sample_stride must cover the frames, mask must select something, n_frames == 0 is a no-op
    buf = (ctypes.c_double * 64)()
    out = (ctypes.c_double * 36)()
    assert built_lib.amc_extract_host_planar(buf, buf, nat.AMC_C128, 2, 16, 1, out, 18, nat.AMC_ALL_FEATURES, 0, 0) == -1
    assert b"sample_stride" in built_lib.amc_last_error_string()
    assert built_lib.amc_extract_host_planar(buf, buf, nat.AMC_C128, 2, 16, 2, out, 18, 0, 0, 0) == -1
    assert built_lib.amc_extract_host_planar(buf, None, nat.AMC_C128, 0, 16, 2, out, 18, nat.AMC_ALL_FEATURES, 0, 0) == 0
    assert built_lib.amc_extract_host_planar(None, None, nat.AMC_C128, 2, 16, 2, out, 18, nat.AMC_ALL_FEATURES, 0, 0) == -1


def test_product_never_imports_the_oracle():
    for py in (ROOT / "amcpy_b200").rglob("*.py"):
        tree = ast.parse(py.read_text())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n.split(".")[0] == "oracle" for n in names), f"{py} imports the oracle"
    for cu in (ROOT / "amcpy_b200" / "csrc").rglob("*.cu*"):
        assert "oracle" not in cu.read_text()


def test_config_defaults_match_reference_contract():
    from amcpy_b200.config import Config

    c = Config()
    assert c.signals.frame_size == 2048 and c.signals.num_frames == 1000 and c.signals.num_threads == 8
    assert c.signals.modulations_with_noise == ("BPSK", "QPSK", "8PSK", "16QAM", "64QAM", "WGN")
    assert list(c.signals.snr_values.values()) == [str(v) for v in range(-10, 21, 2)]
    assert c.signals.mat_info == {"BPSK": "signal_bpsk", "QPSK": "signal_qpsk", "8PSK": "signal_8psk",
                                  "16QAM": "signal_qam16", "64QAM": "signal_qam64", "WGN": "signal_noise"}
    assert c.features.all_features == tuple(range(1, 19)) and c.features.used == (2, 4, 6, 8, 12, 14)
    assert c.features.names[1] == r"$\gamma_{max}$" and c.features.names[18] == "$C_{63}$"
    assert c.paths.mat_filename == "all_modulations.mat"
    assert c.paths.calculated_features.name == "calculated-features" and c.paths.mat_data.name == "mat-data"
    with pytest.raises(Exception):
        c.signals.frame_size = 1  # frozen


def test_stage_layout_planning_fortran_and_c_order():
    from amcpy_b200.feature_extraction import _frames_view

    S, F, L = 4, 5, 40
    base = (np.arange(S * F * L) + 1j * np.arange(S * F * L)).reshape(S, F, L)
    for arr in (np.asfortranarray(base), np.ascontiguousarray(base)):
        seen = np.zeros((S, 3), dtype=bool)
        for view, si, fi in _frames_view(arr, S, 3, 32):
            assert view.shape[1] == 32
            for row, s, f in zip(view, si, fi):
                if f < 3:
                    assert np.array_equal(row, arr[s, f, :32])
                    seen[s, f] = True
        assert seen.all()
    # loadmat's order is consumed as ONE sample-major block (no host transpose)
    views = _frames_view(np.asfortranarray(base), S, 3, 32)
    assert len(views) == 1 and views[0][0].strides == (16, 16 * S * F)


def test_synth_recipe_and_shard_independence():
    from amcpy_b200 import synth

    a = synth.cell(3, 10.0, 10, range(4), 512, seed=9)
    b = synth.cell(3, 10.0, 10, [2, 3], 512, seed=9)
    assert np.array_equal(a[2:], b)                       # frames are keyed individually
    for m, name in enumerate(synth.MODULATIONS[:5]):
        pts = synth.constellation(name)
        assert np.isclose(np.mean(np.abs(pts) ** 2), 1.0)
        x = synth.frame(m, 100.0, 0, 0, 256, seed=1)      # ~noise-free: every sample on the constellation
        assert np.min(np.abs(x[:, None] - pts[None, :]), axis=1).max() < 1e-3
    w = synth.cell(5, 0.0, 5, range(8), 4096, seed=2)
    assert abs(np.mean(np.abs(w) ** 2) - 1.0) < 0.05      # WGN at 0 dB: unit noise power


def test_features_module_surface():
    from amcpy_b200 import features as F

    assert sorted(F._FEATURE_FUNCTIONS) == list(range(1, 19))
    assert [F._FEATURE_FUNCTIONS[i].__name__ for i in (1, 2, 9, 10, 18)] == [
        "_gmax", "_std_abs_phase", "_kurtosis_cnf", "_cumulant_20", "_cumulant_63"]
    assert F._cumulant_42 is F._FEATURE_FUNCTIONS[14]
    with pytest.raises(KeyError):
        F.calculate_features([19], np.zeros(4, dtype=complex))   # unknown id: KeyError before any GPU work
    assert F.calculate_features([], np.zeros(4, dtype=complex)) == []


def test_matio_planar_reader_matches_scipy_and_knows_its_limits(tmp_path):
    """matio.read_planar returns zero-copy planes that reassemble to exactly what scipy.io.loadmat returns;
    compressed variables are inflated (in parallel) and parsed the same way; non-float variables are skipped."""
    import scipy.io

    from amcpy_b200 import matio

    rng = np.random.default_rng(3)
    a = rng.standard_normal((16, 5, 64)) + 1j * rng.standard_normal((16, 5, 64))
    b = (rng.standard_normal((4, 3, 8)) + 1j * rng.standard_normal((4, 3, 8))).astype(np.complex64)
    c = rng.standard_normal((2, 7))
    p = tmp_path / "all_modulations.mat"
    scipy.io.savemat(str(p), {"signal_qpsk": a, "signal_bpsk": b, "real": c, "Modulation": "QPSK", "ints": np.arange(4)})
    got = matio.read_planar(p)
    ref = scipy.io.loadmat(str(p))
    assert set(got) == {"signal_qpsk", "signal_bpsk", "real"}          # char / integer variables are skipped
    for k in got:
        q = got[k]
        full = q.re.reshape(q.shape, order="F")
        if q.im is not None:
            full = full + 1j * q.im.reshape(q.shape, order="F")
        assert q.shape == ref[k].shape and np.array_equal(full, ref[k])
        assert isinstance(q.re, np.memmap)                              # views of the file, not copies
    assert got["signal_bpsk"].dtype == np.float32 and got["real"].im is None
    # element (s, f, n) of the column-major variable = plane[s + S*f + S*F*n]
    S, F, _ = a.shape
    assert got["signal_qpsk"].re[3 + S * 2 + S * F * 10] == a[3, 2, 10].real
    # compressed elements (`save -v7` / do_compression=True): same planes, views of the inflated buffers
    scipy.io.savemat(str(p), {"signal_qpsk": a, "signal_bpsk": b, "real": c, "Modulation": "QPSK"}, do_compression=True)
    got = matio.read_planar(p, max_workers=2)
    ref = scipy.io.loadmat(str(p))
    assert set(got) == {"signal_qpsk", "signal_bpsk", "real"}
    for k in got:
        q = got[k]
        full = q.re.reshape(q.shape, order="F")
        if q.im is not None:
            full = full + 1j * q.im.reshape(q.shape, order="F")
        assert q.shape == ref[k].shape and np.array_equal(full, ref[k]), k
    assert set(matio.read_planar(p, only=["signal_bpsk"])) == {"signal_bpsk"}
    # a damaged zlib stream skips that variable (the caller then asks scipy, which raises)
    raw = bytearray(p.read_bytes())
    raw[128 + 8 + 20:128 + 8 + 40] = bytes(20)
    (tmp_path / "bad.mat").write_bytes(bytes(raw))
    assert len(matio.read_planar(tmp_path / "bad.mat")) < 3
    (tmp_path / "junk.mat").write_bytes(b"not a mat file")
    assert matio.read_planar(tmp_path / "junk.mat") is None


def test_magic_number_unwrap_step_is_the_numpy_step_away_from_ties():
    """The float32 step the fused kernels use for np.unwrap (amc_device.cuh: wrap_step_f32 = dd - 2 pi rint(dd / 2 pi), the
    rounding done with the 1.5 * 2^23 trick) restated in numpy: outside the tie band the kernels re-decide in float64
    (|wrapped| within kTieEps = 4e-6 of pi) it must BE the conditional form of the reference's unwrap
    (features.py:28 -> np.unwrap: |dd| > pi -> dd -+ 2 pi), bit for bit."""
    rng = np.random.default_rng(5)
    two_pi, pi, inv = np.float32(6.28318530717958647692), np.float32(3.14159265358979323846), np.float32(0.15915494309189533577)
    magic = np.float64(12582912.0)
    dd = np.concatenate([rng.uniform(-2 * np.pi, 2 * np.pi, 400000),
                         np.pi + rng.uniform(-1e-4, 1e-4, 50000), -np.pi + rng.uniform(-1e-4, 1e-4, 50000),
                         [0.0, -0.0, 6.2831, -6.2831, 3.0, -3.0]]).astype(np.float32)
    # fmaf(dd, inv, magic): the product of two float32 is exact in float64; one rounding to float32 after the add
    k = (dd.astype(np.float64) * np.float64(inv) + magic).astype(np.float32) - np.float32(magic)
    assert set(np.unique(k)) <= {-1.0, 0.0, 1.0}
    wrapped = (dd.astype(np.float64) - k.astype(np.float64) * np.float64(two_pi)).astype(np.float32)   # fmaf(k, -2 pi, dd)
    conditional = np.where(np.abs(dd) > pi, dd - np.copysign(two_pi, dd), dd).astype(np.float32)
    away = np.abs(pi - np.abs(wrapped)) >= np.float32(4.0e-6)
    assert away.sum() > 400000
    assert np.array_equal(wrapped[away].view(np.uint32), conditional[away].view(np.uint32))
    # and inside the band both forms end within the band: the kernels' tie detection on the wrapped value sees them
    assert np.all(np.abs(pi - np.abs(conditional[~away])) < np.float32(8.0e-6))


def test_atan2_fast_restated_in_numpy_stays_in_the_1e6_class():
    """amc_device.cuh: atan2_fast (degree-7 minimax in q^2 on [0, 1] + the five-instruction octant / quadrant fix-up
    copysign(pi/2 - copysign(|r0 - [|y| <= |x|] pi/2|, x), y)) restated in float32 numpy against np.arctan2 in float64:
    absolute error below 5e-7 over the plane, exact quadrant behaviour for signed zeros (np.angle, features.py:19)."""
    f32 = np.float32
    coef = [0.0026222064831683623, -0.015132382451695708, 0.0411216062546977, -0.07366684174003307,
            0.10573921500635099, -0.14185972498001842, 0.19990396259243107, -0.33332987041851964]

    def atan2_fast(y, x):
        ax, ay = np.abs(x), np.abs(y)
        mx = np.maximum(np.maximum(ax, ay), f32(1.0e-37))
        q = (np.minimum(ax, ay) * (f32(1.0) / mx)).astype(f32)
        s = (q * q).astype(f32)
        p = np.full_like(q, f32(coef[0]))
        for c in coef[1:]:
            p = (p * s + f32(c)).astype(f32)
        r0 = ((q * s).astype(f32) * p + q).astype(f32)
        keep = np.where(ay > ax, f32(0.0), f32(1.0))
        u = (keep * f32(-np.pi / 2) + r0).astype(f32)
        v = np.copysign(np.abs(u), x).astype(f32)
        r = (f32(np.pi / 2) - v).astype(f32)
        return np.copysign(r, y).astype(f32)

    rng = np.random.default_rng(9)
    x = np.concatenate([rng.standard_normal(300000), [0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 1.0, -1.0, 0.0, 0.0, 3.0, -3.0]]).astype(f32)
    y = np.concatenate([rng.standard_normal(300000), [0.0, 0.0, -0.0, -0.0, 1.0, 1.0, -1.0, -1.0, 2.0, -2.0, 0.0, 0.0]]).astype(f32)
    got = atan2_fast(y, x)
    want = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.max(np.abs(got.astype(np.float64) - want)) < 5e-7
    # signed zeros: atan2(+0, +0) = +0, (+0, -0) = pi, (-0, +0) = -0, (-0, -0) = -pi
    z = atan2_fast(np.array([0.0, 0.0, -0.0, -0.0], f32), np.array([0.0, -0.0, 0.0, -0.0], f32))
    assert z[0] == 0 and not np.signbit(z[0]) and z[2] == 0 and np.signbit(z[2])
    assert abs(z[1] - np.pi) < 1e-6 and abs(z[3] + np.pi) < 1e-6


def test_integration_md_binding_blocks_are_valid_python_and_name_real_symbols(built_lib):
    """The two code blocks of INTEGRATION.md section 2 (executed on the GPU by tests/test_gpu_boundary.py) parse, and every
    `_lib.<symbol>` they touch is exported by the built library."""
    text = (Path(__file__).resolve().parent.parent / "INTEGRATION.md").read_text()
    sec = text[text.index("## 2. Patch the reference's stage"):text.index("## 3. Device-resident use")]
    blocks = re.findall(r"```python\n(.*?)```", sec, flags=re.S)
    assert len(blocks) == 2 and "_modulation_process" in blocks[0] and "calculate_features" in blocks[1]
    for b in blocks:
        ast.parse(b)
        for sym in set(re.findall(r"_lib\.(amc_\w+)", b)):
            assert hasattr(built_lib, sym), sym
