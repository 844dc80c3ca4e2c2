"""CPU: pin oracle/amc_oracle.py against the reference's own known answers
(/root/reference/src/amcpy/features.py:258-311) and against the committed outputs of the
unmodified reference (tests/golden/, made by oracle/make_golden.py)."""

import numpy as np
import pytest

from conftest import golden_frames, golden_generic, load_golden
from oracle import amc_oracle as orc


def test_kat_all_features_rtol_1e5():
    # features.py:283-311 - 18 known answers at rtol=1e-5
    got = orc.features_frame(orc.kat_signal())
    for fid, exp in orc.KAT_EXPECTED.items():
        assert np.isclose(got[fid - 1], exp, rtol=1e-5), (fid, got[fid - 1], exp)


def test_kat_faithful_form_matches_and_raises_keyerror():
    sig = orc.kat_signal()
    got = orc.calculate_features_faithful(list(range(1, 19)), sig)
    assert np.allclose(got, [orc.KAT_EXPECTED[i] for i in range(1, 19)], rtol=1e-5)
    assert orc.calculate_features_faithful([18, 1], sig) == [got[17], got[0]]  # order follows ids
    with pytest.raises(KeyError):
        orc.calculate_features_faithful([19], sig)


def test_kat_instantaneous_values():
    # features.py:258-271
    iv = orc.instantaneous(orc.kat_signal())
    assert len(iv["abs"]) == 10 and len(iv["phase"]) == 10 and len(iv["unwrapped_phase"]) == 10
    assert len(iv["frequency"]) == 9 and len(iv["cn_amplitude"]) == 10
    assert np.isclose(iv["abs"][1], np.sqrt(2), atol=1e-10)
    assert np.isclose(iv["cn_amplitude"][0], -1.0, atol=1e-10)
    assert np.isclose(iv["cn_amplitude"][-1], 1.0, atol=1e-10)


def test_kat_moment_values():
    # features.py:274-280
    m = orc.moments(orc.kat_signal())
    assert np.isclose(m["m21"], 57.0, atol=1e-10)
    assert np.isclose(m["m42"], 6133.2, atol=1e-6)
    assert np.isclose(m["m63"], 782724.0, atol=1e-6)


def test_kat_bitwise_vs_reference_outputs():
    g = load_golden("kat10.npz")
    assert np.array_equal(g["signal"], orc.kat_signal())
    assert np.array_equal(orc.features_frame(g["signal"]), g["features"])
    iv = orc.instantaneous(g["signal"])
    for k, v in iv.items():
        assert np.array_equal(v, g[f"iv_{k}"]), k
    for k, v in orc.moments(g["signal"]).items():
        assert np.array_equal(np.asarray(v), g[f"mv_{k}"]), k


@pytest.mark.parametrize("n", [256, 1024, 2048, 4096])
def test_oracle_bitwise_on_realistic_frames(n):
    x, want = golden_frames(n)
    got = orc.features_batch(x)
    # same numpy calls in the same order -> identical bits, except scipy's kurtosis
    # (restated by formula): allow 4 ulp there.
    for fid in range(1, 19):
        if fid in (8, 9):
            assert np.allclose(got[..., fid - 1], want[..., fid - 1], rtol=1e-15, atol=0)
        else:
            assert np.array_equal(got[..., fid - 1], want[..., fid - 1]), fid


@pytest.mark.parametrize("n", [10, 31, 100, 1000, 3000, 512, 8192, 16384, 12000, 32768, 65536])
def test_oracle_on_ragged_sizes(n):
    x, want = golden_generic(n)
    assert np.allclose(orc.features_batch(x), want, rtol=1e-14, atol=0)


@pytest.mark.parametrize("n", [127, 129, 1536, 2047, 2048, 6000, 8191])
def test_oracle_on_hard_cases_from_the_reference(n):
    """Non-power-of-two sizes, carrier offset / DC / scale, NaN frames, complex64 input: the oracle restates what the
    unmodified reference returned (tests/golden/hard_n*.npz)."""
    from conftest import golden_hard

    x, want, want64 = golden_hard(n)
    with np.errstate(all="ignore"):
        got = orc.features_batch(x)
        got64 = orc.features_batch(x[:4].astype(np.complex64))
    assert np.isnan(want[8:]).all() and np.isnan(got[8:]).all()           # a NaN poisons all 18 features
    assert np.allclose(got[:8], want[:8], rtol=1e-13, atol=0)
    assert np.allclose(got64, want64, rtol=1e-5, atol=0)                    # float32 arithmetic in both


@pytest.mark.parametrize("n", [256, 2048, 8192])
def test_oracle_on_narrow_and_extreme_frames_from_the_reference(n):
    """Narrow phase clusters, scales 1e-30 .. 1e+20, degenerate amplitude distributions (tests/golden/narrow_n*.npz)."""
    from conftest import golden_narrow

    x, want = golden_narrow(n)
    with np.errstate(all="ignore"):
        got = orc.features_batch(x)
    keep = np.ones_like(want, dtype=bool)
    keep[6, 3] = False                                   # rounding noise around an exact 0 in both (1e-16)
    assert np.allclose(got[keep], want[keep], rtol=1e-12, atol=0)
    assert abs(got[6, 3]) < 1e-12 and abs(want[6, 3]) < 1e-12


def test_helpers_realistic_frame():
    from amcpy_b200 import synth

    g = load_golden("helpers_n256.npz")
    x = synth.frame(1, 10.0, 10, 0, 256, int(g["seed"]))
    for k, v in orc.instantaneous(x).items():
        assert np.array_equal(v, g[f"iv_{k}"]), k
    for k, v in orc.moments(x).items():
        assert np.array_equal(np.asarray(v), g[f"mv_{k}"]), k


def test_stage_faithful_matches_reference_run_extraction():
    """extract_modulation_faithful == the float32 matrices the reference's run_extraction wrote."""
    from amcpy_b200 import synth

    g = load_golden("stage_16x2x2048.npz")
    snrs = [-10.0 + 2.0 * i for i in range(16)]
    data = synth.dataset(snrs, 2, 2048 + 8, int(g["seed"]))
    for mi, mod in enumerate(synth.MODULATIONS):
        fm = orc.extract_modulation_faithful(data[mi], 16, 2, 2048)
        assert fm.dtype == np.float32
        assert np.allclose(fm, g[f"{mod}_matrix"], rtol=1e-6, atol=0), mod


def test_kurtosis_nan_rule():
    # scipy returns NaN when m2 <= (eps*mean)**2
    assert np.isnan(orc.pearson_kurtosis(np.full(16, 3.0)))
