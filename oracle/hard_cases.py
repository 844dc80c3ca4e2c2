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
