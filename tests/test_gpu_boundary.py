"""GPU: the drop-in boundary - helper value types, layouts/strides/dtypes, the host pipeline,
the `run_extraction` stage with its .mat output contract, and error behaviour of the C ABI."""

import ctypes

import numpy as np
import pytest

from conftest import SNRS, assert_features_close, golden_frames, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    from amcpy_b200 import _native as nat

    nat.require_cuda()
    return torch


# ------------------------------------------------------------------ helper value types
def test_instantaneous_values_kat(torch_cuda):
    # features.py:258-271 through the drop-in class, then bit-level vs the reference's arrays
    from amcpy_b200.features import InstantaneousValues
    from oracle import amc_oracle as orc

    iv = InstantaneousValues(orc.kat_signal())
    assert len(iv.abs) == 10 and len(iv.phase) == 10 and len(iv.unwrapped_phase) == 10
    assert len(iv.frequency) == 9 and len(iv.cn_amplitude) == 10
    assert np.isclose(iv.abs[1], np.sqrt(2), atol=1e-10)
    assert np.isclose(iv.cn_amplitude[0], -1.0, atol=1e-10)
    assert np.isclose(iv.cn_amplitude[-1], 1.0, atol=1e-10)
    g = load_golden("kat10.npz")
    for k in ("abs", "phase", "unwrapped_phase", "frequency", "cn_amplitude"):
        assert np.allclose(getattr(iv, k), g[f"iv_{k}"], rtol=1e-13, atol=1e-14), k
    # the fixture sits exactly on np.unwrap's +-pi tie: the sign pattern must be the reference's
    assert np.array_equal(np.sign(iv.frequency), np.sign(g["iv_frequency"]))


def test_moment_values_kat(torch_cuda):
    # features.py:274-280
    from amcpy_b200.features import MomentValues
    from oracle import amc_oracle as orc

    m = MomentValues(orc.kat_signal())
    assert np.isclose(m.m21, 57.0, atol=1e-10)
    assert np.isclose(m.m42, 6133.2, atol=1e-6)
    assert np.isclose(m.m63, 782724.0, atol=1e-6)
    assert isinstance(m.m21, float) and isinstance(m.m20, complex)
    g = load_golden("kat10.npz")
    for k in ("m20", "m21", "m22", "m40", "m41", "m42", "m43", "m60", "m61", "m62", "m63"):
        assert np.isclose(getattr(m, k), complex(g[f"mv_{k}"]), rtol=1e-12, atol=1e-9), k


def test_helpers_realistic_frame_and_batch(torch_cuda):
    from amcpy_b200 import ops, synth
    from oracle import amc_oracle as orc

    g = load_golden("helpers_n256.npz")
    x = synth.frame(1, 10.0, 10, 0, 256, int(g["seed"]))
    xd = torch_cuda.from_numpy(np.stack([x, x[::-1].copy(), x * 1j])).cuda()   # 3 frames
    iv = ops.instantaneous_batch(xd)
    for k in ("abs", "phase", "unwrapped_phase", "frequency", "cn_amplitude"):
        got = iv[k][0].cpu().numpy()
        assert np.allclose(got, g[f"iv_{k}"], rtol=1e-12, atol=1e-12), k
    want2 = orc.instantaneous(x * 1j)
    for k in want2:
        assert np.allclose(iv[k][2].cpu().numpy(), want2[k], rtol=1e-12, atol=1e-12), k
    mv = ops.moments_batch(xd).cpu().numpy()
    for i, k in enumerate(ops.MOMENT_NAMES):
        assert np.isclose(mv[0, i], complex(g[f"mv_{k}"]), rtol=1e-11, atol=1e-13), k


def test_unwrapped_phase_long_frame_scan(torch_cuda):
    # several 256-sample chunks with carries: unwrapped phase of a noisy rotating tone
    from amcpy_b200 import ops
    from oracle import amc_oracle as orc

    rng = np.random.default_rng(5)
    n = 3000
    x = np.exp(1j * 2.9 * np.arange(n)) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iv = ops.instantaneous_batch(torch_cuda.from_numpy(x[None, :]).cuda())
    want = orc.instantaneous(x)
    assert np.allclose(iv["unwrapped_phase"][0].cpu().numpy(), want["unwrapped_phase"], rtol=1e-12, atol=1e-9)
    assert np.allclose(iv["frequency"][0].cpu().numpy(), want["frequency"], rtol=0, atol=1e-12)


# ------------------------------------------------------------------ layouts, dtypes, strides
def test_batched_entry_takes_snr_frames_samples_tensor(torch_cuda):
    from amcpy_b200 import ops

    x, want = golden_frames(2048)
    xd = torch_cuda.from_numpy(x[2]).cuda()            # (n_snr=4, n_frames=3, 2048) - north_star's entry shape
    got = ops.extract_features(xd)
    assert tuple(got.shape) == (4, 3, 18) and got.dtype == torch_cuda.float64
    assert_features_close(got.cpu().numpy(), want[2])


def test_padded_rows_and_noncontiguous_views(torch_cuda):
    from amcpy_b200 import ops

    x, want = golden_frames(2048)
    flat = x.reshape(-1, 2048)
    pad = np.zeros((flat.shape[0], 2048 + 8), dtype=np.complex128)
    pad[:, :2048] = flat
    pd = torch_cuda.from_numpy(pad).cuda()
    assert_features_close(ops.extract_features(pd[:, :2048]).cpu().numpy(), want)      # row stride 2056: fused
    # sample-major (what loadmat hands out): general kernel in place, or re-laid-out on the device
    sm = torch_cuda.from_numpy(np.asfortranarray(flat)).cuda()    # torch keeps the column-major strides
    assert sm.stride() == (1, flat.shape[0])
    assert_features_close(ops.extract_features(sm, relayout=False).cpu().numpy(), want)    # general kernel in place
    assert torch_cuda.equal(ops.extract_features(sm), ops.extract_features(torch_cuda.from_numpy(flat).cuda()))
    rel = ops.frames_from_sample_major(sm.t().contiguous().view(-1), flat.shape[0], 2048, flat.shape[0])
    assert torch_cuda.equal(rel, torch_cuda.from_numpy(flat).cuda())
    # every second frame
    assert_features_close(ops.extract_features(pd[::2, :2048]).cpu().numpy(), want.reshape(-1, 18)[::2])


def test_complex64_input(torch_cuda):
    from amcpy_b200 import ops
    from oracle import amc_oracle as orc

    x, _ = golden_frames(1024)
    x64 = x.reshape(-1, 1024).astype(np.complex64)
    want = orc.features_batch(x64.astype(np.complex128))   # exact widening: same numbers, float64 arithmetic
    for force in (False, True):
        got = ops.extract_features(torch_cuda.from_numpy(x64).cuda(), force_general=force).cpu().numpy()
        assert_features_close(got, want)
    # and against numpy computing in float32 like the reference does for c64 input (loose: 1e-4)
    ref32 = orc.features_batch(x64)
    got = ops.extract_features(torch_cuda.from_numpy(x64).cuda()).cpu().numpy()
    assert np.allclose(got, ref32, rtol=2e-4, atol=0)


def test_host_pipeline_row_major_and_sample_major(torch_cuda):
    from amcpy_b200 import ops

    x, want = golden_frames(2048)
    flat = x.reshape(-1, 2048)
    big = np.concatenate([flat] * 40)                      # 2880 frames: more than one 64 MiB chunk
    want_big = np.concatenate([want.reshape(-1, 18)] * 40)
    got = ops.extract_features_host(big)
    assert_features_close(got, want_big)
    got_f = ops.extract_features_host(np.asfortranarray(big))
    assert np.array_equal(got_f, got)                      # re-layout on device, then the same kernel
    assert_features_close(ops.extract_features_host(flat[:, :1000]), ops.extract_features_host(
        np.ascontiguousarray(flat[:, :1000])), scale=1e-3)  # padded rows through the general kernel


def test_host_pipeline_pageable_sources_are_staged_through_pinned_memory(torch_cuda):
    """Pageable host sources >= 8 MB (numpy arrays, memory-mapped files) are gathered by host threads into the
    library's pinned staging buffers; pinned sources are copied directly.  Same frames, same kernel: bitwise equal,
    for contiguous rows, padded rows and the sample-major layout."""
    from amcpy_b200 import ops

    x, _ = golden_frames(2048)
    big = np.concatenate([x.reshape(-1, 2048)] * 40)        # 94 MB pageable: two chunks
    pinned = torch_cuda.from_numpy(big).pin_memory()
    direct = ops.extract_features_host(pinned.numpy())      # pinned: no staging
    assert np.array_equal(ops.extract_features_host(big), direct)
    assert np.array_equal(ops.extract_features_host(np.asfortranarray(big)), direct)
    view = big[:, :1024]                                    # padded rows (stride 2048), 47 MB of payload
    assert np.array_equal(ops.extract_features_host(view), ops.extract_features(torch_cuda.from_numpy(
        np.ascontiguousarray(view)).cuda()).cpu().numpy())
    one = big[:1]                                           # a single frame stays on the direct path
    assert np.array_equal(ops.extract_features_host(one), direct[:1])


def test_host_pipeline_planar_planes_match_interleaved(torch_cuda):
    """amc_extract_host_planar (split real / imaginary sample-major planes, as memory-mapped from a .mat file)
    feeds the same kernel the same frames as the interleaved host path: bitwise equal results."""
    from amcpy_b200 import ops

    x, _ = golden_frames(2048)
    big = np.concatenate([x.reshape(-1, 2048)] * 40)        # 2880 frames: more than one 64 MiB chunk
    want = ops.extract_features_host(big)
    nf, n = big.shape
    pad = 5                                                 # plane rows longer than the frames used
    re = np.zeros((n, nf + pad))
    im = np.zeros((n, nf + pad))
    re[:, :nf], im[:, :nf] = big.real.T, big.imag.T
    got = ops.extract_features_host_planar(re.reshape(-1), im.reshape(-1), nf, n, nf + pad)
    assert np.array_equal(got, want)
    got32 = ops.extract_features_host_planar(re.astype(np.float32).reshape(-1), im.astype(np.float32).reshape(-1), nf, n,
                                             nf + pad)
    assert np.array_equal(got32, ops.extract_features_host(big.astype(np.complex64)))
    real_only = ops.extract_features_host_planar(re.reshape(-1), None, nf, n, nf + pad)
    assert np.array_equal(real_only, ops.extract_features_host(big.real.astype(np.complex128)), equal_nan=True)
    with pytest.raises(ValueError):
        ops.extract_features_host_planar(re.reshape(-1)[:100], im.reshape(-1)[:100], nf, n, nf + pad)


def test_determinism_and_grid_independence(torch_cuda):
    from amcpy_b200 import ops

    x, _ = golden_frames(2048)
    xd = torch_cuda.from_numpy(np.concatenate([x.reshape(-1, 2048)] * 10)).cuda()
    a = ops.extract_features(xd)
    b = ops.extract_features(xd)
    assert torch_cuda.equal(a, b)
    # a frame's result must not depend on how many frames share the launch
    c = ops.extract_features(xd[:5])
    assert torch_cuda.equal(a[:5], c)
    assert torch_cuda.equal(a[:72], a[72:144])


# ------------------------------------------------------------------ the stage (.mat in, .mat out)
def test_run_extraction_matches_reference_mat_files(torch_cuda, tmp_path):
    import scipy.io

    from amcpy_b200 import synth
    from amcpy_b200.config import Config, Paths, SignalConfig
    from amcpy_b200.feature_extraction import run_extraction

    g = load_golden("stage_16x2x2048.npz")
    cfg = Config(paths=Paths(root=tmp_path), signals=SignalConfig(num_frames=2))
    cfg.paths.ensure_dirs()
    data = synth.dataset(SNRS, 2, 2048 + 8, int(g["seed"]))
    synth.write_all_modulations_mat(cfg.paths.mat_data / cfg.paths.mat_filename, data, cfg.signals.mat_info)
    run_extraction(cfg)
    for mod in cfg.signals.modulations_with_noise:
        m = scipy.io.loadmat(str(cfg.paths.calculated_features / f"{mod}_features.mat"))
        key = cfg.signals.mat_info[mod]
        assert m[key].dtype == np.float32 and m[key].shape == (16, 2, 18)
        assert np.array_equal(m["Modulation"], g[f"{mod}_modulation"])
        want = g[f"{mod}_matrix"].astype(np.float64)
        got = m[key].astype(np.float64)
        # float32 store (feature_extraction.py:56): 1e-6 class features may differ by one more ulp
        assert np.allclose(got, want, rtol=1.3e-6, atol=0), mod
        cols_1e9 = [c for c in range(18) if c + 1 not in (1, 2, 3, 5, 9)]
        assert np.allclose(got[..., cols_1e9], want[..., cols_1e9], rtol=1.2e-7, atol=0), mod


def test_run_extraction_same_files_from_compressed_and_uncompressed_input(torch_cuda, tmp_path, monkeypatch):
    """Uncompressed Level-5 input goes through the memory-mapped planar path, compressed input is inflated
    by matio and takes the same planar path: the written feature files must be identical (and equal to what
    the scipy.io.loadmat route produces)."""
    import scipy.io

    from amcpy_b200 import matio, synth
    from amcpy_b200.config import Config, Paths, SignalConfig
    from amcpy_b200.feature_extraction import run_extraction

    data = synth.dataset(SNRS, 3, 2048, 17)
    outs = []
    for name, compress in (("plain", False), ("zip", True), ("scipy", True)):
        if name == "scipy":                                  # third leg: the scipy.io.loadmat route
            monkeypatch.setattr(matio, "read_planar", lambda path, **kw: None)
        cfg = Config(paths=Paths(root=tmp_path / name), signals=SignalConfig(num_frames=3))
        cfg.paths.ensure_dirs()
        path = cfg.paths.mat_data / cfg.paths.mat_filename
        scipy.io.savemat(str(path), {cfg.signals.mat_info[m]: data[i] for i, m in enumerate(synth.MODULATIONS)},
                         do_compression=compress)
        if name != "scipy":
            assert set(matio.read_planar(path)) == set(cfg.signals.mat_info.values())
        run_extraction(cfg)
        outs.append({m: scipy.io.loadmat(str(cfg.paths.calculated_features / f"{m}_features.mat"))[cfg.signals.mat_info[m]]
                     for m in cfg.signals.modulations_with_noise})
    for m in outs[0]:
        assert np.array_equal(outs[0][m], outs[1][m], equal_nan=True), m
        assert np.array_equal(outs[0][m], outs[2][m], equal_nan=True), m


def test_run_extraction_fails_loudly_on_short_data(torch_cuda, tmp_path):
    from amcpy_b200 import synth
    from amcpy_b200.config import Config, Paths, SignalConfig
    from amcpy_b200.feature_extraction import run_extraction

    cfg = Config(paths=Paths(root=tmp_path), signals=SignalConfig(num_frames=5))
    cfg.paths.ensure_dirs()
    data = synth.dataset(SNRS, 2, 2048, 3)                  # only 2 frames in the file
    synth.write_all_modulations_mat(cfg.paths.mat_data / cfg.paths.mat_filename, data, cfg.signals.mat_info)
    with pytest.raises(ValueError):
        run_extraction(cfg)


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("dtype", ["c128", "c64"])
def test_feature_mask_profiles_skip_work_not_accuracy(torch_cuda, n, dtype):
    """feature_mask of the C ABI: requested columns are BITWISE what the all-features call returns, columns of
    feature groups the library skipped hold NaN (frame sizes 256..4096 have reduced profiles: moments only,
    amplitude + moments, everything but the FFT)."""
    from amcpy_b200 import ops

    rng = np.random.default_rng(n)
    x = (rng.standard_normal((70, n)) + 1j * rng.standard_normal((70, n))) * 0.8
    xd = torch_cuda.from_numpy(x.astype(np.complex64) if dtype == "c64" else x).cuda()
    full = ops.extract_features(xd).cpu().numpy()
    assert np.isfinite(full).all()
    groups = {"fft": [1], "phase": [2, 3, 5, 9], "amp": [4, 6, 7, 8], "mom": list(range(10, 19))}
    cases = [
        ([10, 12, 18], {"mom"}),                          # moments only
        ([14], {"mom"}),
        ([6, 13], {"amp", "mom"}),                        # amplitude + moments
        ([4, 7, 8], {"amp", "mom"}),                      # (cheapest compiled superset)
        ([2, 4, 6, 8, 12, 14], {"phase", "amp", "mom"}),  # the reference's default feature list: no FFT
        ([3, 5, 7, 9, 13, 15], {"phase", "amp", "mom"}),
        ([9], {"phase", "amp", "mom"}),
        ([1], {"fft", "phase", "amp", "mom"}),            # anything with feature 1: the full kernel
        (list(range(1, 19)), {"fft", "phase", "amp", "mom"}),
    ]
    for ids, computed in cases:
        got = ops.extract_features(xd, feature_mask=ops.feature_mask_of(ids)).cpu().numpy()
        for name, fids in groups.items():
            cols = [f - 1 for f in fids]
            if name in computed:
                assert np.array_equal(got[:, cols], full[:, cols]), (ids, name)
            else:
                assert np.isnan(got[:, cols]).all(), (ids, name)
    # the batched operator API passes the mask of the ids it returns
    from amcpy_b200 import features as F
    sel = F.calculate_features_batch([12, 14, 10], xd).cpu().numpy()
    assert np.array_equal(sel, full[:, [11, 13, 9]])
    sel_h = F.calculate_features_batch([2, 6], np.ascontiguousarray(xd.cpu().numpy()))
    assert np.array_equal(sel_h, full[:, [1, 5]])


def test_feature_mask_is_ignored_where_no_reduced_profile_exists(torch_cuda):
    from amcpy_b200 import ops

    rng = np.random.default_rng(1)
    for n in (100, 8192):
        x = torch_cuda.from_numpy(rng.standard_normal((5, n)) + 1j * rng.standard_normal((5, n))).cuda()
        full = ops.extract_features(x).cpu().numpy()
        got = ops.extract_features(x, feature_mask=ops.feature_mask_of([10, 11])).cpu().numpy()
        assert np.array_equal(got, full)
    with pytest.raises(KeyError):
        ops.feature_mask_of([0])


def test_plain_c_host_runs_the_hot_path(torch_cuda, tmp_path):
    """examples/extract_host.c: a C program with host buffers through amc_extract_host - QPSK cumulants come out
    where theory puts them (|C20| ~ 0, |C40| ~ 1, |C42| ~ 1) and the kernels were launched by the library."""
    from conftest import run_c_example

    res = run_c_example(tmp_path, 256)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "frame 0:" in res.stdout and "mean |C20|" in res.stdout
    launches = int(res.stdout.rsplit(";", 1)[1].split()[0])
    assert launches >= 1


# ------------------------------------------------------------------ C ABI error behaviour
def test_c_abi_error_codes(torch_cuda):
    from amcpy_b200 import _native as nat

    lib = nat.lib()
    x = torch_cuda.zeros((4, 256), dtype=torch_cuda.complex128, device="cuda")
    out = torch_cuda.zeros((4, 18), dtype=torch_cuda.float64, device="cuda")
    args = dict(dt=nat.AMC_C128, nf=4, n=256, fs=256, ss=1, os=18, mask=nat.AMC_ALL_FEATURES)

    def call(**kw):
        a = {**args, **kw}
        return lib.amc_extract_batch(x.data_ptr(), a["dt"], a["nf"], a["n"], a["fs"], a["ss"], out.data_ptr(), a["os"],
                                     a["mask"], 0, None)

    assert call() == 0
    assert call(nf=0) == 0
    assert call(dt=7) == -1 and b"iq_dtype" in lib.amc_last_error_string()
    assert call(n=0) == -1
    assert call(os=17) == -1
    assert call(mask=0) == -1
    assert call(ss=0) == -1
    assert call(n=1 << 21, fs=1 << 21) == -2          # power of two beyond the general kernel's FFT workspace limit
    assert call(n=(1 << 20) + 1, fs=(1 << 20) + 1) == -2 and b"Bluestein" in lib.amc_last_error_string()
    assert lib.amc_extract_batch(None, nat.AMC_C128, 4, 256, 256, 1, out.data_ptr(), 18, nat.AMC_ALL_FEATURES, 0, None) == -1
    with pytest.raises(nat.AmcError):
        nat.check(call(dt=7))
    torch_cuda.cuda.synchronize()


# ------------------------------------------------------------------ BASELINE config 5: `amcpy full`
def test_full_pipeline_extract_consume_train_eval(torch_cuda, tmp_path, capsys):
    """synth -> extract (GPU) -> column select / standardise / split -> classifier train + per-SNR eval,
    through the CLI entry (the reference's `amcpy full` cannot run: SURVEY.md App. B.1-B.3)."""
    from amcpy_b200 import main as cli

    root = str(tmp_path)
    cli.main(["--root", root, "--num-frames", "60", "--frame-size", "1024", "synth", "--seed", "5"])
    cli.main(["--root", root, "--num-frames", "60", "--frame-size", "1024", "full", "--epochs", "8"])
    out = capsys.readouterr().out
    assert "All feature calculations complete!" in out and "accuracy by SNR index" in out
    assert len(list((tmp_path / "calculated-features").glob("*_features.mat"))) == 6
    assert len(list((tmp_path / "ann").glob("model-*.pt"))) == 1
    val_acc = float(out.split("val_acc ")[-1].split(";")[0])
    assert val_acc > 0.6, out[-400:]            # 6 classes, chance = 0.17; high-SNR features separate them


# ------------------------------------------------------------------ on-device generator
def test_device_generator_recipe_determinism_and_shards(torch_cuda):
    from amcpy_b200 import ops, synth
    from oracle import amc_oracle as orc

    snrs = [-10.0, 0.0, 20.0]
    full = synth.dataset_device(6, snrs, 8, 2048, torch_cuda.device("cuda"), seed=99).view(6, 3, 8, 2048)
    again = synth.dataset_device(6, snrs, 8, 2048, torch_cuda.device("cuda"), seed=99).view(6, 3, 8, 2048)
    assert torch_cuda.equal(full, again)
    shard = synth.dataset_device(6, snrs, 3, 2048, torch_cuda.device("cuda"), seed=99, first_frame=5).view(6, 3, 3, 2048)
    assert torch_cuda.equal(shard, full[:, :, 5:8])                      # any frame shard == the same frames of the full set
    other = synth.dataset_device(6, snrs, 8, 2048, torch_cuda.device("cuda"), seed=100).view(6, 3, 8, 2048)
    assert not torch_cuda.equal(other, full)
    p = (full.abs() ** 2).mean(dim=(2, 3)).cpu().numpy()                 # mean power per (mod, snr)
    for m in range(5):
        assert np.allclose(p[m], 1.0 + 10.0 ** (-np.array(snrs) / 10.0), rtol=0.05), (m, p[m])
    assert np.allclose(p[5], 10.0 ** (-np.array(snrs) / 10.0), rtol=0.05)
    x = full.cpu().numpy()
    for m, name in enumerate(synth.MODULATIONS[:5]):                     # at 20 dB every sample sits near a constellation point
        pts = synth.constellation(name)
        d = np.abs(x[m, 2].reshape(-1)[:, None] - pts[None, :]).min(axis=1)
        assert np.percentile(d, 99) < 0.35, name
        counts = np.bincount(np.abs(x[m, 2].reshape(-1)[:, None] - pts[None, :]).argmin(axis=1), minlength=len(pts))
        assert counts.min() > 0.6 * counts.mean(), (name, counts)        # symbols roughly uniform
    # features of generated frames agree with the oracle on the same bits
    sub = full[:, 1, 0]                                                  # one 0 dB frame per class
    assert_features_close(ops.extract_features(sub).cpu().numpy(), orc.features_batch(sub.cpu().numpy()))


def test_device_consumer_matches_host_consumer(torch_cuda):
    from amcpy_b200 import ops, synth
    from amcpy_b200.config import Config, SignalConfig
    from amcpy_b200.consumer import (Standardizer, stack_features, stack_features_device, standardize_device,
                                     stratified_split_device)

    cfg = Config(signals=SignalConfig(num_frames=12, frame_size=512))
    snrs = [float(v) for v in cfg.signals.snr_values.values()]
    x = synth.dataset_device(6, snrs, 12, 512, torch_cuda.device("cuda"), seed=4).view(6, 16, 12, 512)
    feats = {m: ops.extract_features(x[i]) for i, m in enumerate(cfg.signals.modulations_with_noise)}
    xd, yd = stack_features_device(cfg, feats, "training")
    xh, yh = stack_features(cfg, "training", {m: v.cpu().numpy().astype(np.float32) for m, v in feats.items()})
    assert np.array_equal(xd.cpu().numpy(), xh) and np.array_equal(yd.cpu().numpy(), yh)
    xs, mean, scale = standardize_device(xd)
    sc = Standardizer().fit(xh)
    assert np.allclose(xs.cpu().numpy(), sc.transform(xh), rtol=1e-5, atol=1e-6)
    xtr, xte, ytr, yte = stratified_split_device(xs, yd, cfg.training.test_size, cfg.training.random_state)
    assert xtr.shape[0] + xte.shape[0] == xs.shape[0] and xte.shape[0] == 6 * round(0.2 * 72)
    assert torch_cuda.bincount(yte).tolist() == [round(0.2 * 72)] * 6


def test_cuda_graph_capture_and_replay(torch_cuda):
    """amc_init builds the tables eagerly, after which amc_extract_batch only enqueues kernels: the call can be captured
    into a CUDA graph (launch-bound small batches) and replayed on new data in the same buffers."""
    from amcpy_b200 import _native as nat
    from amcpy_b200 import ops, synth

    nat.check(nat.lib().amc_init(0))
    snrs = [0.0, 10.0]
    x = synth.dataset_device(6, snrs, 4, 1024, torch_cuda.device("cuda"), seed=1)
    out = torch_cuda.empty((x.shape[0], 18), dtype=torch_cuda.float64, device="cuda")
    want1 = ops.extract_features(x).clone()
    s = torch_cuda.cuda.Stream()
    s.wait_stream(torch_cuda.cuda.current_stream())
    g = torch_cuda.cuda.CUDAGraph()
    with torch_cuda.cuda.stream(s):
        ops.extract_features(x, out=out, stream=s)          # warm-up on the capture stream
        s.synchronize()
        with torch_cuda.cuda.graph(g, stream=s):
            ops.extract_features(x, out=out, stream=s)
    out.zero_()
    g.replay()
    torch_cuda.cuda.synchronize()
    assert torch_cuda.equal(out, want1)
    x.copy_(synth.dataset_device(6, snrs, 4, 1024, torch_cuda.device("cuda"), seed=2))
    want2 = ops.extract_features(x).clone()
    g.replay()
    torch_cuda.cuda.synchronize()
    assert torch_cuda.equal(out, want2) and not torch_cuda.equal(want1, want2)


def test_concurrent_host_calls_share_a_device(torch_cuda):
    """Four host threads call amc_extract_host on the same device at once: each takes its own copy/compute pipe of the
    library (amc_api.cu: HostPipe pool); results are bitwise what a single call returns."""
    from concurrent.futures import ThreadPoolExecutor

    from amcpy_b200 import ops, synth

    rng_sets = [np.concatenate([synth.cell(m, 6.0, 8, range(40), 2048, seed=50 + k) for m in range(6)]) for k in range(6)]
    want = [ops.extract_features_host(a) for a in rng_sets]
    with ThreadPoolExecutor(max_workers=6) as pool:          # more threads than pipes: two of them queue
        got = list(pool.map(ops.extract_features_host, rng_sets))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_concurrent_streams_with_careful_path_frames(torch_cuda):
    """Eight streams enqueue extractions at once, every batch holding frames the careful path has to redo: the per-launch
    ticket ring must neither lose a tagged row nor touch another launch's output (slots only grow: atomicMax)."""
    from amcpy_b200 import ops, synth
    from conftest import golden_narrow

    xn, _ = golden_narrow(2048)
    batches = []
    for k in range(8):
        base = np.concatenate([synth.cell(m, 8.0, 9, range(20), 2048, seed=70 + k) for m in range(6)])
        base[k::11] = xn[k % 10]                               # sprinkle narrow / extreme-scale frames
        batches.append(torch_cuda.from_numpy(base).cuda())
    want = [ops.extract_features(b).clone() for b in batches]
    torch_cuda.cuda.synchronize()
    streams = [torch_cuda.cuda.Stream() for _ in batches]
    outs = [torch_cuda.empty_like(w) for w in want]
    for rep in range(5):
        for o in outs:
            o.zero_()
        torch_cuda.cuda.synchronize()
        for b, o, s in zip(batches, outs, streams):
            ops.extract_features(b, out=o, stream=s)
        torch_cuda.cuda.synchronize()
        for o, w in zip(outs, want):
            assert torch_cuda.equal(o, w)
            assert not (o[:, 0].view(torch_cuda.int64) == 0x7ff8b200a3c10001).any()   # no tag left behind


def test_integration_md_reference_side_binding_runs_as_printed(torch_cuda, tmp_path):
    """INTEGRATION.md section 2 prints the ctypes stub a maintainer of the reference would paste into
    feature_extraction.py (`_modulation_process`, :42-82) and features.py (`calculate_features`, :214-232).  The test
    takes those two code blocks from the document as they stand, points the CDLL at the built library, runs them on a
    small all_modulations.mat and compares with the reference's own files (golden) - the document cannot rot."""
    import ctypes
    import re
    from pathlib import Path

    import scipy.io

    from amcpy_b200 import _native, synth
    from amcpy_b200.config import Config, Paths, SignalConfig

    text = (Path(__file__).resolve().parent.parent / "INTEGRATION.md").read_text()
    sec = text[text.index("## 2. Patch the reference's stage"):text.index("## 3. Device-resident use")]
    blocks = re.findall(r"```python\n(.*?)```", sec, flags=re.S)
    assert len(blocks) == 2 and "_modulation_process" in blocks[0] and "calculate_features" in blocks[1]
    lib_path = str(_native.build())
    ns = {"scipy": scipy, "ctypes": ctypes, "np": np}
    exec(blocks[0].replace('ctypes.CDLL("libamcpy_b200.so")', f"ctypes.CDLL({lib_path!r})"), ns)   # noqa: S102
    ns["_FEATURE_FUNCTIONS"] = {i: None for i in range(1, 19)}
    exec(blocks[1], ns)                                                                              # noqa: S102

    g = load_golden("stage_16x2x2048.npz")
    cfg = Config(paths=Paths(root=tmp_path), signals=SignalConfig(num_frames=2))
    cfg.paths.ensure_dirs()
    data = synth.dataset(SNRS, 2, 2048 + 8, int(g["seed"]))
    synth.write_all_modulations_mat(cfg.paths.mat_data / cfg.paths.mat_filename, data, cfg.signals.mat_info)
    for mod in cfg.signals.modulations_with_noise:
        ns["_modulation_process"](mod, cfg)
        m = scipy.io.loadmat(str(cfg.paths.calculated_features / f"{mod}_features.mat"))
        key = cfg.signals.mat_info[mod]
        assert m[key].dtype == np.float32 and m[key].shape == (16, 2, 18)
        assert np.allclose(m[key].astype(np.float64), g[f"{mod}_matrix"].astype(np.float64), rtol=1.3e-6, atol=0), mod
    # the per-frame operator of the same section: values in the order of feature_ids, KeyError on an unknown id
    x = data[1][3, 0, :2048]
    vals = ns["calculate_features"]([18, 1, 7], x)
    want = g["QPSK_matrix"][3, 0].astype(np.float64)
    assert np.allclose(vals, [want[17], want[0], want[6]], rtol=1.3e-6)
    with pytest.raises(KeyError):
        ns["calculate_features"]([19], x)
