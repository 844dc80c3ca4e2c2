"""Drop-in for `amcpy.feature_extraction` (/root/reference/src/amcpy/feature_extraction.py).

`run_extraction(cfg)` keeps the reference's contract (feature_extraction.py:85-99, :42-82):
reads `<root>/mat-data/all_modulations.mat`, takes the first `frame_size` samples of the first
`num_frames` frames at each of `len(snr_values)` SNRs of every modulation variable, and writes
`<root>/calculated-features/{MOD}_features.mat` holding "Modulation" and a float32
(n_snr, n_frames, 18) matrix under `mat_info[MOD]`.

What replaces the 6 processes x num_threads threads x Queue: the file is parsed ONCE, each
modulation's Fortran-ordered block goes to a GPU in place (sample-major, no host transpose) through
the library's chunked copy/compute pipeline; on one process the modulations overlap (three host threads per
GPU, each with its own pipe of the library) and are spread over the visible GPUs; under torchrun the flattened
(modulation, frame, snr) index space is cut into one contiguous range per rank (`plan_shards`), so all ranks are
busy whatever the number of modulations, and rank 0 assembles and writes the six files.
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
                              device: int = 0, q_range: tuple | None = None) -> np.ndarray:
    """Same result as `extract_modulation` from the memory-mapped planes of a Level-5 variable: element
    (snr, frame, sample) of the column-major (S, F, L) variable is plane[snr + S*frame + S*F*sample], i.e.
    frame q = snr + S*frame of a sample-major block with sample stride S*F.

    q_range=(q0, q1): only the frames q0 <= q < q1 of that block (a rank's shard); returns the float64
    (q1 - q0, 18) rows instead of the assembled matrix."""
    if len(arr.shape) != 3:
        raise ValueError(f"expected a 3-D (snr, frame, sample) array, got shape {arr.shape}")
    S, F, L = arr.shape
    if S < n_snr or F < n_frames or L < frame_size:
        raise ValueError(
            f"data shape {arr.shape} is smaller than the configured "
            f"(n_snr={n_snr}, num_frames={n_frames}, frame_size={frame_size})"
        )
    nq = S * n_frames
    if q_range is not None:
        q0, q1 = q_range
        if not 0 <= q0 <= q1 <= nq:
            raise ValueError(f"q_range {q_range} outside [0, {nq}]")
        im = arr.im[q0:] if arr.im is not None else None
        return ops.extract_features_host_planar(arr.re[q0:], im, q1 - q0, frame_size, S * F, device=device)
    feats = ops.extract_features_host_planar(arr.re, arr.im, nq, frame_size, S * F, device=device)
    return assemble_matrix(feats, 0, S, n_snr, n_frames)


def assemble_matrix(rows: np.ndarray, q0: int, S: int, n_snr: int, n_frames: int, fm: np.ndarray | None = None):
    """Scatter float64 rows of the frames q0, q0+1, ... (q = snr + S*frame) into the float32
    (n_snr, n_frames, 18) matrix of feature_extraction.py:56 (float64 -> float32 on store, :35)."""
    if fm is None:
        fm = np.zeros((n_snr, n_frames, ops.N_FEATURES), dtype=np.float32)
    idx = q0 + np.arange(rows.shape[0])
    si, fi = idx % S, idx // S
    keep = (si < n_snr) & (fi < n_frames)
    fm[si[keep], fi[keep], :] = rows[keep]
    return fm


class _MatSource:
    """The input file, parsed once: memory-mapped planes when the file allows it, scipy.io.loadmat otherwise
    (loaded lazily and only if some variable needs it)."""

    def __init__(self, path, only=None):
        self.path = str(path)
        self.planar = matio.read_planar(self.path, only=only) or {}
        self._mat = None
        self._lock = threading.Lock()   # run_extraction uses several host threads

    def loadmat(self):
        with self._lock:
            if self._mat is None:
                self._mat = scipy.io.loadmat(self.path)
            return self._mat

    def planar_3d(self, key: str):
        arr = self.planar.get(key)
        return arr if arr is not None and len(arr.shape) == 3 else None

    def features(self, key: str, modulation: str, n_snr: int, n_frames: int, frame_size: int, device: int):
        arr = self.planar_3d(key)
        if arr is not None:
            return extract_modulation_planar(arr, n_snr, n_frames, frame_size, device=device)
        mat = self.loadmat()
        if key not in mat:
            raise KeyError(f"variable {key!r} for {modulation} not found in {self.path}")
        return extract_modulation(mat[key], n_snr, n_frames, frame_size, device=device)


def _rank_world():
    """(rank, world, device index).  More ranks than visible GPUs share devices (LOCAL_RANK modulo the device count):
    each rank still drives its own copy/compute pipe of the library."""
    import torch

    n_dev = max(1, torch.cuda.device_count())
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")) % n_dev)


def _save(modulation: str, cfg: Config, fm: np.ndarray) -> None:
    key = cfg.signals.mat_info[modulation]
    out_path = cfg.paths.calculated_features / f"{modulation}_features.mat"
    scipy.io.savemat(str(out_path), {"Modulation": modulation, key: fm})   # feature_extraction.py:77-81


def _modulation_process(modulation: str, cfg: Config, data_mat=None, device: int = 0) -> np.ndarray:
    """One modulation: slice -> GPU features -> savemat (feature_extraction.py:42-82).  Returns the matrix it wrote."""
    t0 = time.perf_counter()
    print(f"[{modulation}] Starting feature extraction on cuda:{device} ...")
    if data_mat is None:
        data_mat = _MatSource(cfg.paths.mat_data / cfg.paths.mat_filename)
    key = cfg.signals.mat_info[modulation]
    fm = data_mat.features(key, modulation, len(cfg.signals.snr_values), cfg.signals.num_frames,
                           cfg.signals.frame_size, device)
    _save(modulation, cfg, fm)
    print(f"[{modulation}] Done in {time.perf_counter() - t0:.2f}s -> {cfg.paths.calculated_features / (modulation + '_features.mat')}")
    return fm


def plan_shards(cfg: Config, sources: dict, world: int):
    """The flattened (modulation, frame, snr) index space cut into one contiguous range per rank
    (sharding.shard_range): [(rank, modulation, q0, q1), ...] with q = snr + S*frame inside a modulation, S taken from
    the data (`sources`: {modulation: S}).  Frames are equal-cost independent units, so the static split is balanced
    and every GPU is busy whatever the number of modulations."""
    from .sharding import shard_range

    mods = list(cfg.signals.modulations_with_noise)
    sizes = [sources[m] * cfg.signals.num_frames for m in mods]
    total = sum(sizes)
    plan = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        base = 0
        for m, sz in zip(mods, sizes):
            q0, q1 = max(lo - base, 0), min(hi - base, sz)
            if q0 < q1:
                plan.append((r, m, q0, q1))
            base += sz
    return plan


def extract_all(cfg: Config) -> dict:
    """{modulation: float32 (n_snr, n_frames, 18)} for every modulation, written to calculated-features/ as well
    (on rank 0 under torchrun; the other ranks return {}).  The in-memory result is what `amcpy full` hands to the
    feature consumer - no .mat round trip."""
    import torch

    from . import _native as nat

    nat.require_cuda()  # no CPU fallback
    cfg.paths.ensure_dirs()
    rank, world, local_rank = _rank_world()
    mods = list(cfg.signals.modulations_with_noise)
    n_snr, n_frames, frame_size = len(cfg.signals.snr_values), cfg.signals.num_frames, cfg.signals.frame_size
    data_mat = _MatSource(cfg.paths.mat_data / cfg.paths.mat_filename, only=[cfg.signals.mat_info[m] for m in mods])
    result: dict = {}
    if world > 1:
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            dist.init_process_group(backend="gloo")           # host-side exchange of the small feature blocks only
        arrs = {m: data_mat.planar_3d(cfg.signals.mat_info[m]) for m in mods}
        if all(a is not None for a in arrs.values()):
            # contiguous shard of the flattened (modulation, frame, snr) space per rank; rank 0 assembles and writes
            plan = plan_shards(cfg, {m: arrs[m].shape[0] for m in mods}, world)
            pieces = []
            for r, m, q0, q1 in plan:
                if r == rank:
                    rows = extract_modulation_planar(arrs[m], n_snr, n_frames, frame_size, device=local_rank, q_range=(q0, q1))
                    pieces.append((m, q0, rows.astype(np.float32)))      # the float32 store of feature_extraction.py:35
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(pieces, gathered, dst=0)
            if rank == 0:
                for m in mods:
                    result[m] = np.zeros((n_snr, n_frames, ops.N_FEATURES), dtype=np.float32)
                for plist in gathered:
                    for m, q0, rows in plist:
                        assemble_matrix(rows, q0, arrs[m].shape[0], n_snr, n_frames, result[m])
                for m in mods:
                    _save(m, cfg, result[m])
        else:
            # a file matio cannot map (scipy.io.loadmat route): whole modulations per rank, gathered the same way
            mine = {m: _modulation_process(m, cfg, data_mat, device=local_rank) for m in mods[rank::world]}
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(mine, gathered, dst=0)
            if rank == 0:
                for d in gathered:
                    result.update(d)
        dist.barrier()
    else:
        n_dev = max(1, torch.cuda.device_count())
        # the modulations overlap on the GPU(s): each host thread drives its own copy/compute pipe of the library
        # (amc_api.cu: HostPipe pool), so one modulation's host-side gather runs while another's chunk is on PCIe
        workers = min(len(mods), 3 * n_dev)
        with ThreadPoolExecutor(max_workers=workers) as pool:
            futs = {m: pool.submit(_modulation_process, m, cfg, data_mat, i % n_dev) for i, m in enumerate(mods)}
            for m, fu in futs.items():
                result[m] = fu.result()  # re-raise the first failure (the reference ignores exit codes)
    if rank == 0:
        print("All feature calculations complete!")
    return result


def run_extraction(cfg: Config) -> None:
    """Feature extraction for all modulation types (feature_extraction.py:85-99)."""
    extract_all(cfg)
