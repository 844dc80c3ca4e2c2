"""Drop-in for `amcpy.feature_extraction` (/root/reference/src/amcpy/feature_extraction.py).

`run_extraction(cfg)` keeps the reference's contract (feature_extraction.py:85-99, :42-82):
reads `<root>/mat-data/all_modulations.mat`, takes the first `frame_size` samples of the first
`num_frames` frames at each of `len(snr_values)` SNRs of every modulation variable, and writes
`<root>/calculated-features/{MOD}_features.mat` holding "Modulation" and a float32
(n_snr, n_frames, 18) matrix under `mat_info[MOD]`.

What replaces the 6 processes x num_threads threads x Queue: the file is parsed ONCE, each
modulation's Fortran-ordered block goes to a GPU in place (sample-major, no host transpose) through
the library's chunked copy/compute pipeline, and the modulations are spread over the visible GPUs
(one host thread per GPU; under torchrun one rank per GPU takes every world_size-th modulation).
An uncompressed Level-5 file (scipy's `savemat` default) is not even parsed by scipy: `matio.read_planar`
memory-maps the separately stored real / imaginary planes and the GPU interleaves them
(`ops.extract_features_host_planar`); compressed (`save -v7`) variables are inflated in parallel by `matio`
and take the same route.  `scipy.io.loadmat` is only the I/O fallback for files `matio` cannot parse.
Deliberate deviations from the reference: shape mismatches and per-modulation
failures raise instead of being printed over, and elapsed (not CPU) time is reported.
"""

from __future__ import annotations

import os
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.io

from . import matio, ops
from .config import Config


def _frames_view(parsed: np.ndarray, n_snr: int, n_frames: int, frame_size: int):
    """2-D (frames, frame_size) views of parsed[:n_snr, :n_frames, :frame_size] that the host
    pipeline can consume in place, with the index map back to (snr, frame).

    Returns a list of (view2d, snr_index_array, frame_index_array)."""
    es = parsed.itemsize
    S, F, L = parsed.shape
    st = tuple(s // es for s in parsed.strides)
    if st == (1, S, S * F):  # Fortran order (scipy.io.loadmat): frame i = snr + S*frame, sample-major
        v = np.lib.stride_tricks.as_strided(parsed, shape=(S * n_frames, frame_size), strides=(es, S * F * es),
                                            writeable=False)
        idx = np.arange(S * n_frames)
        return [(v, idx % S, idx // S)]
    views = []
    for s in range(n_snr):  # anything else: one (row-per-frame where possible) block per SNR
        views.append((parsed[s, :n_frames, :frame_size], np.full(n_frames, s), np.arange(n_frames)))
    return views


def extract_modulation(parsed: np.ndarray, n_snr: int, n_frames: int, frame_size: int, device: int = 0) -> np.ndarray:
    """float32 (n_snr, n_frames, 18) for one modulation variable (feature_extraction.py:52-74)."""
    if parsed.ndim != 3:
        raise ValueError(f"expected a 3-D (snr, frame, sample) array, got shape {parsed.shape}")
    S, F, L = parsed.shape
    if S < n_snr or F < n_frames or L < frame_size:
        raise ValueError(
            f"data shape {parsed.shape} is smaller than the configured "
            f"(n_snr={n_snr}, num_frames={n_frames}, frame_size={frame_size})"
        )
    if not np.iscomplexobj(parsed):
        parsed = parsed.astype(np.complex128)
    if parsed.dtype not in (np.complex64, np.complex128):
        parsed = parsed.astype(np.complex128)
    fm = np.zeros((n_snr, n_frames, ops.N_FEATURES), dtype=np.float32)  # feature_extraction.py:56
    for view, si, fi in _frames_view(parsed, n_snr, n_frames, frame_size):
        feats = ops.extract_features_host(view, device=device)
        keep = si < n_snr
        fm[si[keep], fi[keep], :] = feats[keep]  # float64 -> float32 on store (feature_extraction.py:35)
    return fm


def extract_modulation_planar(arr: "matio.PlanarArray", n_snr: int, n_frames: int, frame_size: int,
                              device: int = 0) -> np.ndarray:
    """Same result as `extract_modulation` from the memory-mapped planes of a Level-5 variable: element
    (snr, frame, sample) of the column-major (S, F, L) variable is plane[snr + S*frame + S*F*sample], i.e.
    frame q = snr + S*frame of a sample-major block with sample stride S*F."""
    if len(arr.shape) != 3:
        raise ValueError(f"expected a 3-D (snr, frame, sample) array, got shape {arr.shape}")
    S, F, L = arr.shape
    if S < n_snr or F < n_frames or L < frame_size:
        raise ValueError(
            f"data shape {arr.shape} is smaller than the configured "
            f"(n_snr={n_snr}, num_frames={n_frames}, frame_size={frame_size})"
        )
    nq = S * n_frames
    feats = ops.extract_features_host_planar(arr.re, arr.im, nq, frame_size, S * F, device=device)
    fm = np.zeros((n_snr, n_frames, ops.N_FEATURES), dtype=np.float32)  # feature_extraction.py:56
    idx = np.arange(nq)
    si, fi = idx % S, idx // S
    keep = si < n_snr
    fm[si[keep], fi[keep], :] = feats[keep]  # float64 -> float32 on store (feature_extraction.py:35)
    return fm


class _MatSource:
    """The input file, parsed once: memory-mapped planes when the file allows it, scipy.io.loadmat otherwise
    (loaded lazily and only if some variable needs it)."""

    def __init__(self, path, only=None):
        self.path = str(path)
        self.planar = matio.read_planar(self.path, only=only) or {}
        self._mat = None
        self._lock = threading.Lock()   # run_extraction uses one host thread per GPU

    def loadmat(self):
        with self._lock:
            if self._mat is None:
                self._mat = scipy.io.loadmat(self.path)
            return self._mat

    def features(self, key: str, modulation: str, n_snr: int, n_frames: int, frame_size: int, device: int):
        arr = self.planar.get(key)
        if arr is not None and len(arr.shape) == 3:
            return extract_modulation_planar(arr, n_snr, n_frames, frame_size, device=device)
        mat = self.loadmat()
        if key not in mat:
            raise KeyError(f"variable {key!r} for {modulation} not found in {self.path}")
        return extract_modulation(mat[key], n_snr, n_frames, frame_size, device=device)


def _rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _modulation_process(modulation: str, cfg: Config, data_mat=None, device: int = 0) -> None:
    """One modulation: slice -> GPU features -> savemat (feature_extraction.py:42-82)."""
    t0 = time.perf_counter()
    print(f"[{modulation}] Starting feature extraction on cuda:{device} ...")
    if data_mat is None:
        data_mat = _MatSource(cfg.paths.mat_data / cfg.paths.mat_filename)
    key = cfg.signals.mat_info[modulation]
    fm = data_mat.features(key, modulation, len(cfg.signals.snr_values), cfg.signals.num_frames,
                           cfg.signals.frame_size, device)
    out_path = cfg.paths.calculated_features / f"{modulation}_features.mat"
    scipy.io.savemat(str(out_path), {"Modulation": modulation, key: fm})
    print(f"[{modulation}] Done in {time.perf_counter() - t0:.2f}s -> {out_path}")


def run_extraction(cfg: Config) -> None:
    """Feature extraction for all modulation types (feature_extraction.py:85-99)."""
    import torch

    from . import _native as nat

    nat.require_cuda()  # no CPU fallback
    cfg.paths.ensure_dirs()
    rank, world, local_rank = _rank_world()
    mods = list(cfg.signals.modulations_with_noise)
    mine = mods[rank::world] if world > 1 else mods  # one rank per GPU: every world_size-th modulation
    data_mat = _MatSource(cfg.paths.mat_data / cfg.paths.mat_filename, only=[cfg.signals.mat_info[m] for m in mine])
    if world > 1:
        for m in mine:
            _modulation_process(m, cfg, data_mat, device=local_rank)
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.barrier()
    else:
        n_dev = max(1, torch.cuda.device_count())
        if n_dev == 1:
            for m in mods:
                _modulation_process(m, cfg, data_mat, device=0)
        else:
            with ThreadPoolExecutor(max_workers=n_dev) as pool:
                futs = [pool.submit(_modulation_process, m, cfg, data_mat, i % n_dev) for i, m in enumerate(mods)]
                for fu in futs:
                    fu.result()  # re-raise the first failure (the reference ignores exit codes)
    if rank == 0:
        print("All feature calculations complete!")
