"""Synthetic IQ frames for the six signal classes of the extraction stage.

The reference ships no generator (its data came from GNU Radio captures,
/root/reference/old/read_binary_stream.py:19-59); SURVEY.md §8(d) fixes the recipe used here:
i.i.d. uniform symbols, one sample per symbol, unit mean power, plus complex AWGN of total
variance 10**(-SNR/10) (half per rail).  WGN is noise only.

Every frame is keyed by (seed, modulation index, snr index, frame index) through a
counter-based Philox stream, so any shard on any rank regenerates bit-identical frames
regardless of how the (mod, snr, frame) index space is partitioned.
"""

from __future__ import annotations

import numpy as np

MODULATIONS = ("BPSK", "QPSK", "8PSK", "16QAM", "64QAM", "WGN")


def _qam_grid(m_side: int) -> np.ndarray:
    lv = np.arange(-(m_side - 1), m_side, 2, dtype=np.float64)
    pts = (lv[:, None] + 1j * lv[None, :]).ravel()
    return pts / np.sqrt(np.mean(np.abs(pts) ** 2))


def constellation(mod: str) -> np.ndarray:
    """Unit-mean-power constellation points (complex128); empty for WGN."""
    if mod == "BPSK":
        return np.array([1.0 + 0j, -1.0 + 0j])
    if mod == "QPSK":
        return np.exp(1j * (np.pi / 4 + np.arange(4) * np.pi / 2))
    if mod == "8PSK":
        return np.exp(1j * np.arange(8) * np.pi / 4)
    if mod == "16QAM":
        return _qam_grid(4)
    if mod == "64QAM":
        return _qam_grid(8)
    if mod == "WGN":
        return np.zeros(0, dtype=np.complex128)
    raise KeyError(mod)


def frame(mod_idx: int, snr_db: float, snr_idx: int, frame_idx: int, n: int, seed: int = 2024) -> np.ndarray:
    """One complex128 frame of n samples."""
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[frame_idx, snr_idx, mod_idx, 0]))
    pts = constellation(MODULATIONS[mod_idx])
    sigma = np.sqrt(10.0 ** (-snr_db / 10.0) / 2.0)
    noise = sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    if pts.size == 0:
        return noise
    return pts[rng.integers(0, pts.size, size=n)] + noise


def cell(mod_idx: int, snr_db: float, snr_idx: int, frame_ids, n: int, seed: int = 2024) -> np.ndarray:
    """(len(frame_ids), n) complex128, C-contiguous."""
    frame_ids = list(frame_ids)
    out = np.empty((len(frame_ids), n), dtype=np.complex128)
    for i, f in enumerate(frame_ids):
        out[i] = frame(mod_idx, snr_db, snr_idx, f, n, seed)
    return out


def dataset(snr_dbs, n_frames: int, n: int, seed: int = 2024, mods=range(6)) -> np.ndarray:
    """(len(mods), n_snr, n_frames, n) complex128."""
    snr_dbs = list(snr_dbs)
    mods = list(mods)
    out = np.empty((len(mods), len(snr_dbs), n_frames, n), dtype=np.complex128)
    for mi, m in enumerate(mods):
        for si, s in enumerate(snr_dbs):
            out[mi, si] = cell(m, s, si, range(n_frames), n, seed)
    return out


def write_all_modulations_mat(path, data: np.ndarray, mat_info: dict, mod_names=MODULATIONS) -> None:
    """Write mat-data/all_modulations.mat with the six variables the reference reads
    (/root/reference/src/amcpy/config.py:101-110, feature_extraction.py:46-48)."""
    import scipy.io

    scipy.io.savemat(str(path), {mat_info[m]: data[i] for i, m in enumerate(mod_names)})


def dataset_torch(n_mods: int, n_snr: int, n_frames: int, n: int, device, seed: int = 2024,
                  snr_lo: float = -10.0, snr_step: float = 2.0, dtype=None):
    """Device-side equivalent for bench-scale sets that must never touch the host
    (82.6 GB at BASELINE config 3).  Same recipe, torch's Philox stream (so not bit-identical
    to `dataset`); returns (n_mods*n_snr*n_frames, n) complex128 on `device`.  Plumbing only."""
    import torch

    dtype = dtype or torch.complex128
    rdt = torch.float64 if dtype == torch.complex128 else torch.float32
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_mods, n_snr, n_frames, n), dtype=dtype, device=device)
    for mi in range(n_mods):
        pts = constellation(MODULATIONS[mi % 6])
        pts_t = torch.from_numpy(pts).to(device=device, dtype=dtype) if pts.size else None
        for si in range(n_snr):
            sigma = float(np.sqrt(10.0 ** (-(snr_lo + snr_step * si) / 10.0) / 2.0))
            re = torch.randn((n_frames, n), generator=g, device=device, dtype=rdt)
            im = torch.randn((n_frames, n), generator=g, device=device, dtype=rdt)
            z = torch.complex(re, im) * sigma
            if pts_t is not None:
                idx = torch.randint(0, pts_t.numel(), (n_frames, n), generator=g, device=device)
                z = z + pts_t[idx]
            out[mi, si] = z
            del re, im, z
    return out.reshape(n_mods * n_snr * n_frames, n)


def dataset_device(n_mods: int, snr_dbs, n_frames: int, n: int, device, seed: int = 2024, first_frame: int = 0,
                   dtype=None, mods=None):
    """Hand-written CUDA generator (C ABI `amc_generate_frames`): (n_mods*n_snr*n_frames, n) complex on
    `device`, cells ordered (modulation, snr).  Counter-based: `first_frame` selects a shard of the
    frame axis that is bit-identical to the same frames of the full set."""
    import torch

    from . import _native as nat

    dtype = dtype or torch.complex128
    snr_dbs = list(snr_dbs)
    mods = list(mods) if mods is not None else list(range(n_mods))
    cell_mod = torch.tensor([m % 6 for m in mods for _ in snr_dbs], dtype=torch.int32, device=device)
    cell_snr = torch.tensor([si for _ in mods for si in range(len(snr_dbs))], dtype=torch.int32, device=device)
    cell_sig = torch.tensor([float(np.sqrt(10.0 ** (-s / 10.0) / 2.0)) for _ in mods for s in snr_dbs],
                            dtype=torch.float64, device=device)
    n_cells = cell_mod.numel()
    out = torch.empty((n_cells * n_frames, n), dtype=dtype, device=device)
    with torch.cuda.device(device):
        rc = nat.lib().amc_generate_frames(
            out.data_ptr(), nat.AMC_C128 if dtype == torch.complex128 else nat.AMC_C64, n_cells, n_frames, first_frame,
            n, cell_mod.data_ptr(), cell_snr.data_ptr(), cell_sig.data_ptr(), seed, torch.cuda.current_stream().cuda_stream)
    nat.check(rc)
    return out
