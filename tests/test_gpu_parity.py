"""GPU parity tests proper: the CUDA path (through the C ABI) against the committed outputs of
the unmodified reference (tests/golden/) and against the oracle on the same seeded inputs.
Tolerances are the north_star's: relative 1e-9 on moment/cumulant (and amplitude) features,
1e-6 on atan2- / FFT-derived features (1, 2, 3, 5, 9)."""

import numpy as np
import pytest

from conftest import assert_features_close, golden_frames, golden_generic, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    from amcpy_b200 import _native as nat

    nat.require_cuda()
    return torch


def test_kat_10_samples_through_calculate_features(torch_cuda):
    # the reference's own test_all_features (features.py:283-311), rtol=1e-5, via the drop-in API
    from amcpy_b200.features import _FEATURE_FUNCTIONS, calculate_features
    from oracle import amc_oracle as orc

    sig = orc.kat_signal()
    got = calculate_features(list(range(1, 19)), sig)
    for fid, exp in orc.KAT_EXPECTED.items():
        assert np.isclose(got[fid - 1], exp, rtol=1e-5), (fid, got[fid - 1], exp)
        assert np.isclose(_FEATURE_FUNCTIONS[fid](sig), exp, rtol=1e-5)
    # and at our own classes against the reference's float64 outputs
    assert_features_close(got, load_golden("kat10.npz")["features"])
    assert calculate_features([18, 1], sig) == [got[17], got[0]]
    with pytest.raises(KeyError):
        calculate_features([0], sig)


@pytest.mark.parametrize("n", [256, 1024, 2048, 4096])
def test_fused_kernel_vs_reference_golden(torch_cuda, n):
    from amcpy_b200 import ops

    x, want = golden_frames(n)
    got = ops.extract_features(torch_cuda.from_numpy(x).cuda()).cpu().numpy()
    assert got.shape == want.shape == (6, 4, 3, 18)
    assert_features_close(got, want)


@pytest.mark.parametrize("n", [256, 2048])
def test_general_kernel_vs_reference_golden(torch_cuda, n):
    from amcpy_b200 import ops

    x, want = golden_frames(n)
    got = ops.extract_features(torch_cuda.from_numpy(x).cuda(), force_general=True).cpu().numpy()
    assert_features_close(got, want)


@pytest.mark.parametrize("n", [10, 31, 100, 1000, 3000, 512, 8192, 16384, 12000, 32768, 65536])
def test_ragged_and_large_frame_sizes(torch_cuda, n):
    from amcpy_b200 import ops

    x, want = golden_generic(n)
    got = ops.extract_features(torch_cuda.from_numpy(x).cuda()).cpu().numpy()
    assert_features_close(got, want)


@pytest.mark.parametrize("n", [127, 128, 129, 255, 257, 1000, 1023, 1025, 1536, 2047, 3000, 4097, 6000, 8191, 8193])
def test_non_power_of_two_sizes_bluestein_vs_oracle_and_direct_dft(torch_cuda, n):
    """np.fft.fft takes any length (features.py:68): sizes 128..8192 that are not powers of two run a float32
    Bluestein FFT in the general kernel (127 and 8193 the float64 direct DFT).  Both must match the oracle;
    everything but feature 1 must be bitwise the same in the two modes."""
    from amcpy_b200 import ops, synth
    from oracle import amc_oracle as orc

    x = np.concatenate([synth.cell(m, snr, 3, range(2), n, seed=n) for m, snr in ((0, 18.0), (3, 4.0), (5, -6.0))])
    x[1] *= 37.5                                           # scale must not matter
    x[2] += 0.3 - 0.2j                                     # a DC line: one dominant bin
    x[3] *= np.exp(2j * np.pi * 0.123 * np.arange(n))      # carrier offset: the peak moves between bins
    want = orc.features_batch(x)
    xd = torch_cuda.from_numpy(x).cuda()
    got = ops.extract_features(xd).cpu().numpy()
    ref = ops.extract_features(xd, direct_dft=True).cpu().numpy()
    assert_features_close(got, want)
    assert_features_close(ref, want)
    assert np.array_equal(got[:, 1:], ref[:, 1:])
    rel = np.max(np.abs(got[:, 0] - ref[:, 0]) / ref[:, 0])
    assert rel < 1e-6, f"N={n}: Bluestein vs float64 DFT {rel:.2e}"
    # complex64 input takes the same route
    g32 = ops.extract_features(torch_cuda.from_numpy(x.astype(np.complex64)).cuda()).cpu().numpy()
    w32 = orc.features_batch(x.astype(np.complex64).astype(np.complex128))
    assert_features_close(g32, w32)


@pytest.mark.parametrize("n", [127, 129, 1536, 2047, 2048, 6000, 8191])
def test_hard_cases_against_outputs_of_the_reference(torch_cuda, n):
    """tests/golden/hard_n*.npz hold what the UNMODIFIED reference returned for frames with a carrier offset, a DC
    line, extreme scales and NaNs, at power-of-two and other sizes, for complex128 and complex64 input."""
    from amcpy_b200 import ops
    from conftest import golden_hard

    x, want, want64 = golden_hard(n)
    for force in (False, True):
        got = ops.extract_features(torch_cuda.from_numpy(x).cuda(), force_general=force).cpu().numpy()
        assert np.isnan(got[8:]).all()
        assert_features_close(got[:8], want[:8])
    # complex64: the reference's numpy computes in float32 (4e-7 away from its own complex128 result and worse on
    # the cumulants); the GPU widens exactly and computes like complex128 - compared at 2e-4
    got64 = ops.extract_features(torch_cuda.from_numpy(x[:4].astype(np.complex64)).cuda()).cpu().numpy()
    assert np.allclose(got64, want64, rtol=2e-4, atol=0)


@pytest.mark.parametrize("n", [256, 2048, 16384])
def test_exact_zeros_octant_points_and_signed_zero(torch_cuda, n):
    """Samples that are exactly 0, on the axes / diagonals, or carry a negative zero: np.angle's
    conventions (atan2 of signed zeros, exact octant values) and |x| = 0 must survive the fast paths."""
    from amcpy_b200 import ops
    from oracle import amc_oracle as orc

    rng = np.random.default_rng(n)
    x = (rng.standard_normal((4, n)) + 1j * rng.standard_normal((4, n))) * 0.7
    special = np.array([0 + 0j, 1 + 1j, -2 + 2j, -3 - 3j, 0.5 - 0.5j, 0 + 3j, 0 - 1j, -1 + 0j, 2 + 0j,
                        complex(-0.0, 0.0), complex(-0.0, -0.0), complex(0.0, -0.0), complex(-1.0, -0.0)])
    for r in range(4):
        pos = rng.choice(n, size=n // 8, replace=False)
        x[r, pos] = special[rng.integers(0, len(special), size=pos.size)]
    x[3, : n // 2] = 0.0                                  # half of a frame exactly zero
    want = orc.features_batch(x)
    for force in (False, True):
        got = ops.extract_features(torch_cuda.from_numpy(x).cuda(), force_general=force).cpu().numpy()
        assert_features_close(got, want)


@pytest.mark.parametrize("n", [256, 2048])
@pytest.mark.parametrize("snr_db", [30.0, 40.0, 50.0])
def test_high_snr_amplitude_features_keep_the_1e9_class(torch_cuda, n, snr_db):
    """The amplitude features (4 std of |cn|, 8 kurtosis of cn) must meet 1e-9 against the two-pass oracle on
    nearly constant-modulus frames, far above the reference's own SNR grid (<= 20 dB).  (Deriving sum (r-mu)^2
    from sum|x|^2 and sum|x| - one FP64 operation per sample cheaper - fails this test: cancellation
    ~ (mu/sigma)^2 ulps; the kernels accumulate the centred sums explicitly.)"""
    from amcpy_b200 import ops, synth
    from oracle import amc_oracle as orc

    x = np.concatenate([synth.cell(m, snr_db, 15, range(3), n, seed=99) for m in (0, 1, 2)])   # BPSK, QPSK, 8PSK
    want = orc.features_batch(x)
    got = ops.extract_features(torch_cuda.from_numpy(x).cuda()).cpu().numpy()
    for fid in (4, 6, 7, 8):
        rel = np.max(np.abs(got[:, fid - 1] - want[:, fid - 1]) / np.abs(want[:, fid - 1]))
        assert rel <= 1e-9, f"feature {fid} at {snr_db} dB, N={n}: rel err {rel:.3e}"


@pytest.mark.parametrize("n", [100, 256, 1024, 2048, 4096, 16384])
def test_nan_samples_poison_their_frame_and_only_their_frame(torch_cuda, n):
    """A NaN anywhere in a frame makes all 18 reference features NaN (FFT, means, sums); GPU min/max
    instructions drop NaNs, so the kernels apply the rule explicitly.  Neighbouring frames of the same
    launch (same CTA, same finalisation batch) must be untouched."""
    from amcpy_b200 import ops
    from oracle import amc_oracle as orc

    rng = np.random.default_rng(n + 5)
    x = (rng.standard_normal((40, n)) + 1j * rng.standard_normal((40, n))) * 0.9
    bad = {3: (0, complex(np.nan, 0.25)), 4: (n - 1, complex(0.5, np.nan)), 17: (n // 2 + 1, complex(np.nan, np.nan)),
           38: (n // 3, complex(-np.nan, 1.0))}
    for r, (pos, val) in bad.items():
        x[r, pos] = val
    with np.errstate(all="ignore"):
        want = orc.features_batch(x)
    assert np.isnan(want[list(bad)]).all() and np.isfinite(np.delete(want, list(bad), axis=0)).all()
    for force in (False, True):
        got = ops.extract_features(torch_cuda.from_numpy(x).cuda(), force_general=force).cpu().numpy()
        assert np.isnan(got[list(bad)]).all(), (n, force)
        good = np.delete(np.arange(40), list(bad))
        assert_features_close(got[good], want[good])
    # the reduced feature profiles apply the same rule
    got = ops.extract_features(torch_cuda.from_numpy(x).cuda(), feature_mask=ops.feature_mask_of([2, 6, 12])).cpu().numpy()
    assert np.isnan(got[list(bad)]).all() and np.isfinite(got[good][:, [1, 5, 11]]).all()


@pytest.mark.parametrize("n", [256, 2048, 8192])
def test_careful_path_frames_vs_reference_golden(torch_cuda, n):
    """Frames the float32 parts of the fused kernels cannot handle alone (narrow phase clusters up to 60 dB, scales
    1e-30 .. 1e+20, degenerate amplitude distributions): detected from the frame's own sums and recomputed by the
    general kernel in the launch that follows every fused launch - WITHOUT force_general, at the ordinary tolerances,
    against outputs of the unmodified reference (tests/golden/narrow_n*.npz)."""
    from amcpy_b200 import ops
    from conftest import golden_narrow

    x, want = golden_narrow(n)
    for force in (False, True):
        got = ops.extract_features(torch_cuda.from_numpy(x).cuda(), force_general=force).cpu().numpy()
        assert np.isfinite(got).all()
        # feature 4 of the noise-free two-level frame is rounding noise around an exact 0 (reference: 1e-16)
        assert abs(got[6, 3]) < 1e-12 and abs(want[6, 3]) < 1e-12
        keep = [i for i in range(10) if i != 6]
        assert_features_close(got[keep], want[keep])
        assert_features_close(got[6:7], want[6:7], ids=[f for f in range(1, 19) if f != 4])
    # the hand-over is per frame: ordinary frames in the same launch are bitwise what they are alone
    xo, _ = golden_frames(n) if n != 8192 else (golden_generic(n)[0], None)
    xo = xo.reshape(-1, n)[:6]
    mixed = np.concatenate([xo[:3], x[:5], xo[3:], x[5:]])
    got = ops.extract_features(torch_cuda.from_numpy(mixed).cuda())
    alone = ops.extract_features(torch_cuda.from_numpy(xo).cuda())
    assert torch_cuda.equal(got[[0, 1, 2, 8, 9, 10]], alone)
    # reduced feature profiles hand over too (requested columns still in tolerance)
    ids = [2, 4, 6, 12]
    got = ops.extract_features(torch_cuda.from_numpy(x).cuda(), feature_mask=ops.feature_mask_of(ids)).cpu().numpy()
    keep = [i for i in range(10) if i != 6]
    assert_features_close(got[keep], want[keep], ids=ids)


def test_unknown_flag_bits_are_rejected(torch_cuda):
    """The A/B experiment kernels are not part of the product ABI: their flag bits are invalid arguments."""
    from amcpy_b200 import _native as nat
    from amcpy_b200 import ops

    x, _ = golden_frames(2048)
    xd = torch_cuda.from_numpy(x.reshape(-1, 2048)[:2]).cuda()
    for bits in (nat.AMC_FLAG_FUSED_SPT8, nat.AMC_FLAG_FUSED_WS, 1 << 9):
        with pytest.raises(nat.AmcError) as ei:
            ops.extract_features(xd, extra_flags=bits)
        assert ei.value.code == -1


def test_first_call_does_not_block_and_init_is_idempotent(torch_cuda):
    from amcpy_b200 import _native as nat

    assert nat.lib().amc_init(0) == 0
    assert nat.lib().amc_init(0) == 0
    assert nat.lib().amc_init(4096) == -1
    assert nat.lib().amc_workspace_bytes(nat.AMC_C128, 48000, 2048, 0) == 0
    assert nat.lib().amc_workspace_bytes(nat.AMC_C128, 10, 1000, 0) == (1000 + 2048) * 8
    assert nat.lib().amc_workspace_bytes(nat.AMC_C128, 48000, 2048, 1) == 2 * (2 * 2048 * 32768 + 2048 * 144)
