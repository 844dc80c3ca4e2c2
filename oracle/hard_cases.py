"""Seeded inputs of the hard-case golden fixtures (tests/golden/hard_n*.npz, written by oracle/make_golden.py from the
unmodified reference).  Test infrastructure: imported by oracle/make_golden.py and tests/ only."""

from __future__ import annotations

import numpy as np

from amcpy_b200 import synth

HARD_SIZES = (127, 129, 1536, 2047, 2048, 6000, 8191)

def hard_case_inputs(n: int) -> np.ndarray:
    """10 frames of `n` samples: 0-3 plain (BPSK 18 dB, 16QAM 4 dB, 64QAM 10 dB, WGN), 4 scaled by 37.5, 5 with a DC
    offset, 6 with a carrier offset of 0.123 cycles/sample, 7 scaled by 1e-3 with carrier offset -0.31, 8 and 9 with a NaN
    (real part of sample 0; imaginary part of the last sample)."""
    k = np.arange(n)
    x = np.stack([
        synth.frame(0, 18.0, 14, 0, n, 77), synth.frame(3, 4.0, 7, 1, n, 77), synth.frame(4, 10.0, 10, 2, n, 77),
        synth.frame(5, -6.0, 2, 3, n, 77), synth.frame(1, 12.0, 11, 4, n, 77) * 37.5,
        synth.frame(2, 8.0, 9, 5, n, 77) + (0.3 - 0.2j), synth.frame(1, 14.0, 12, 6, n, 77) * np.exp(2j * np.pi * 0.123 * k),
        synth.frame(3, 16.0, 13, 7, n, 77) * 1e-3 * np.exp(-2j * np.pi * 0.31 * k),
        synth.frame(0, 6.0, 8, 8, n, 77), synth.frame(4, 6.0, 8, 9, n, 77),
    ])
    x[8, 0] = complex(np.nan, x[8, 0].imag)
    x[9, n - 1] = complex(x[9, n - 1].real, np.nan)
    return x


NARROW_SIZES = (256, 2048, 8192)

def narrow_case_inputs(n: int) -> np.ndarray:
    """10 frames of `n` samples that the float32 parts of the fused kernels cannot handle on their own (round-2 soak
    findings + ADVICE.md): 0 unmodulated carrier at 40 dB, 1 carrier at phase 2.0 rad / 50 dB, 2 WGN 26 dB below a DC
    line (phase spread 0.025 rad: the round-1 soak case), 3 carrier at 60 dB with a slow 1e-4 cycles/sample rotation,
    4 QPSK scaled by 1e-30, 5 16QAM scaled by 1e+20, 6 noise-free two-level amplitude (|x| in {1, 3}, every
    |cn_amplitude| equal: the reference's feature 4 is an exact 0), 7 two-level amplitude at 70 dB SNR, 8 WGN scaled by
    1e-25, 9 BPSK at 18 dB scaled by 3e+17."""
    rng = np.random.default_rng(4242 + n)
    k = np.arange(n)

    def noise(db):
        s = 10.0 ** (-db / 20.0) / np.sqrt(2.0)
        return s * (rng.standard_normal(n) + 1j * rng.standard_normal(n))

    lvl = np.where(np.arange(n) % 2 == 0, 1.0, 3.0)
    ph = np.exp(1j * rng.uniform(-np.pi, np.pi, n))
    x = np.stack([
        1.0 + noise(40.0),
        np.exp(2.0j) * (1.0 + noise(50.0)),
        (1.0 + 0.0j) + noise(26.0),
        np.exp(2j * np.pi * 1e-4 * k) * (1.0 + noise(60.0)),
        synth.frame(1, 12.0, 11, 4, n, 78) * 1e-30,
        synth.frame(3, 16.0, 13, 5, n, 78) * 1e20,
        lvl * ph,
        lvl * ph + noise(70.0),
        synth.frame(5, 0.0, 5, 8, n, 78) * 1e-25,
        synth.frame(0, 18.0, 14, 9, n, 78) * 3e17,
    ])
    return x
