"""CPU oracle for the amcpy feature-extraction hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's per-frame operator
(`/root/reference/src/amcpy/features.py`) and of its per-modulation driver
(`/root/reference/src/amcpy/feature_extraction.py`).  It is *not* part of the
product: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  The product path
(`amcpy_b200`) never touches it and fails loudly when its CUDA library is missing.

Parity status: PINNED.  The restatement is checked against
  * the reference's own 18 known answers + moment/instantaneous spot checks
    (features.py:258-311) in tests/test_oracle.py, and
  * outputs of the unmodified reference, imported from /root/reference/src in the build
    container by oracle/make_golden.py and committed under tests/golden/.

Numerical semantics restated here (numpy 2.3 / scipy 1.18 as probed in SURVEY.md App. A):
  abs      = hypot(re, im)                          features.py:27
  angle    = arctan2(im, re), angle(0+0j) = 0       features.py:28
  unwrap   = np.unwrap defaults (period 2*pi)       features.py:29
  freq     = diff(unwrapped) / (2*pi)               features.py:30
  cn_amp   = abs / mean(abs) - 1                    features.py:31
  std      = ddof=1                                 features.py:74,79,85,91
  kurtosis = biased Pearson m4/m2**2 about the mean, NaN when m2 <= (eps*mean)**2
             (scipy.stats.kurtosis(fisher=False))   features.py:107,113
  M_pq     = mean(x**(p-q) * conj(x)**q)            features.py:46-58
"""

from __future__ import annotations

import numpy as np

TWO_PI = 2.0 * np.pi
N_FEATURES = 18

# Tolerance classes (BASELINE.json north_star).  Index = feature id.
#   1e-6 relative: FFT / atan2 derived (1, 2, 3, 5, 9);  1e-9 relative: everything else.
RTOL = {fid: (1e-6 if fid in (1, 2, 3, 5, 9) else 1e-9) for fid in range(1, N_FEATURES + 1)}


# --------------------------------------------------------------------------------------
# helper value types (features.py:17-31 and :39-58)
# --------------------------------------------------------------------------------------
def instantaneous(x: np.ndarray) -> dict:
    """The five arrays of the reference's InstantaneousValues (features.py:27-31)."""
    amp = np.abs(x)
    ph = np.angle(x)
    up = np.unwrap(ph)
    return {
        "abs": amp,
        "phase": ph,
        "unwrapped_phase": up,
        "frequency": np.diff(up) / TWO_PI,
        "cn_amplitude": amp / np.mean(amp) - 1,
    }


def moments(x: np.ndarray) -> dict:
    """The eleven mixed moments of the reference's MomentValues (features.py:46-58).

    m21, m42 and m62 are real parts, exactly as the reference truncates them
    (m62 is the real part of a genuinely complex quantity - features.py:57)."""
    c = np.conj(x)
    return {
        "m20": np.mean(x**2),
        "m21": np.mean(x * c).real,
        "m22": np.mean(c**2),
        "m40": np.mean(x**4),
        "m41": np.mean(x**3 * c),
        "m42": np.mean(x**2 * c**2).real,
        "m43": np.mean(x * c**3),
        "m60": np.mean(x**6),
        "m61": np.mean(x**5 * c),
        "m62": np.mean(x**4 * c**2).real,
        "m63": np.mean(x**3 * c**3),
    }


def pearson_kurtosis(v: np.ndarray) -> float:
    """scipy.stats.kurtosis(v, fisher=False, bias=True) restated."""
    mu = np.mean(v)
    d = v - mu
    m2 = np.mean(d**2)
    m4 = np.mean(d**4)
    with np.errstate(all="ignore"):
        if m2 <= (np.finfo(np.float64).eps * mu) ** 2:
            return float("nan")
        return float(m4 / m2**2.0)


def cumulant_features(m: dict) -> list:
    """Features 10..18 from a moments dict (features.py:116-185), incl. the reference's
    `+3*m20**3` in C60 (features.py:147) and the real-only m62 in C62 (features.py:160-170)."""
    m20, m21, m22 = m["m20"], m["m21"], m["m22"]
    m40, m41, m42, m43 = m["m40"], m["m41"], m["m42"], m["m43"]
    m60, m61, m62, m63 = m["m60"], m["m61"], m["m62"], m["m63"]
    return [
        np.abs(m20),
        np.abs(m21),
        np.abs(m40 - 3 * m20**2),
        np.abs(m41 - 3 * m20 * m21),
        np.abs(m42 - np.abs(m20) ** 2 - 2 * m21**2),
        np.abs(m60 - 15 * m20 * m40 + 3 * m20**3),
        np.abs(m61 - 5 * m21 * m40 - 10 * m20 * m41 + 30 * m20**2 * m21),
        np.abs(m62 - 6 * m20 * m42 - 8 * m21 * m41 - m22 * m40 + 6 * m20**2 * m22 + 24 * m21**2 * m20),
        np.abs(m63 - 9 * m21 * m42 + 12 * m21**3 - 3 * m20 * m43 - 3 * m22 * m41 + 18 * m20 * m21 * m22),
    ]


# --------------------------------------------------------------------------------------
# all 18 features of one frame, intermediates shared (fast form used by the parity tests)
# --------------------------------------------------------------------------------------
def features_frame(x: np.ndarray) -> np.ndarray:
    """float64[18]; column k = feature id k+1 (features.py:192-211)."""
    x = np.asarray(x)
    n = x.shape[0]
    iv = instantaneous(x)
    out = np.empty(N_FEATURES, dtype=np.float64)
    spec = np.abs(np.fft.fft(x))
    out[0] = np.max(spec**2 / n)                               # features.py:68-69
    out[1] = np.std(np.abs(iv["phase"]), ddof=1)              # :74
    out[2] = np.std(iv["phase"], ddof=1)                      # :79
    out[3] = np.std(np.abs(iv["cn_amplitude"]), ddof=1)       # :85
    out[4] = np.std(iv["frequency"], ddof=1)                  # :91
    out[5] = np.mean(iv["abs"])                               # :96
    out[6] = np.sqrt(np.sum(iv["abs"])) / n                   # :101
    out[7] = pearson_kurtosis(iv["cn_amplitude"])             # :107
    out[8] = pearson_kurtosis(iv["frequency"])                # :113
    out[9:18] = cumulant_features(moments(x))                 # :116-185
    return out


def features_batch(frames: np.ndarray) -> np.ndarray:
    """frames: (..., N) complex -> (..., 18) float64."""
    frames = np.asarray(frames)
    flat = frames.reshape(-1, frames.shape[-1])
    out = np.empty((flat.shape[0], N_FEATURES), dtype=np.float64)
    with np.errstate(all="ignore"):
        for i in range(flat.shape[0]):
            out[i] = features_frame(flat[i])
    return out.reshape(frames.shape[:-1] + (N_FEATURES,))


# --------------------------------------------------------------------------------------
# "faithful" form: the reference's work pattern (used only to time a CPU baseline)
# --------------------------------------------------------------------------------------
def _kurt(v):
    try:  # the reference calls scipy (features.py:107,113); same image on the GPU box
        from scipy import stats

        return float(stats.kurtosis(v, fisher=False))
    except Exception:  # pragma: no cover
        return pearson_kurtosis(v)


# one entry per feature id, every entry rebuilding its inputs from the raw frame exactly
# as features.py:66-185 does (InstantaneousValues 4x, MomentValues 9x per frame).
_FAITHFUL = {
    1: lambda x: float(np.max(np.abs(np.fft.fft(x)) ** 2 / len(x))),
    2: lambda x: float(np.std(np.abs(np.angle(x)), ddof=1)),
    3: lambda x: float(np.std(np.angle(x), ddof=1)),
    4: lambda x: float(np.std(np.abs(instantaneous(x)["cn_amplitude"]), ddof=1)),
    5: lambda x: float(np.std(instantaneous(x)["frequency"], ddof=1)),
    6: lambda x: float(np.mean(np.abs(x))),
    7: lambda x: float(np.sqrt(np.sum(np.abs(x))) / len(x)),
    8: lambda x: _kurt(instantaneous(x)["cn_amplitude"]),
    9: lambda x: _kurt(instantaneous(x)["frequency"]),
}
for _i in range(9):
    _FAITHFUL[10 + _i] = (lambda k: (lambda x: float(cumulant_features(moments(x))[k])))(_i)


def calculate_features_faithful(feature_ids, x) -> list:
    """Same call shape and cost structure as features.py:214-232 (KeyError on unknown id)."""
    return [_FAITHFUL[fid](x) for fid in feature_ids]


def extract_modulation_faithful(parsed: np.ndarray, n_snr: int, n_frames: int, frame_size: int) -> np.ndarray:
    """feature_extraction.py:52-74 for one modulation, single-threaded: every (snr, frame)
    view truncated to frame_size -> float32 (n_snr, n_frames, 18)."""
    fm = np.zeros((n_snr, n_frames, N_FEATURES), dtype=np.float32)
    ids = list(range(1, N_FEATURES + 1))
    with np.errstate(all="ignore"):
        for s in range(n_snr):
            for f in range(n_frames):
                fm[s, f, :] = calculate_features_faithful(ids, parsed[s, f, 0:frame_size])
    return fm


# --------------------------------------------------------------------------------------
# the reference's own fixture (features.py:240-255) and known answers (:286-305)
# --------------------------------------------------------------------------------------
def kat_signal() -> np.ndarray:
    k = np.arange(10)
    return (k * (-1.0) ** k) * (1 - 1j)


KAT_EXPECTED = {
    1: 405.0, 2: 0.940293603578649, 3: 1.5903100728408748, 4: 0.3312693299999689,
    5: 0.5153882032022075, 6: 6.363961030678928, 7: 0.7977443845417482,
    8: 1.7757575757575754, 9: 1.0627162629757787, 10: 57.0, 11: 57.0, 12: 3613.8,
    13: 3613.8, 14: 3613.8, 15: 3905583.0, 16: 1094628.0, 17: 311904.0, 18: 1094628.0,
}
