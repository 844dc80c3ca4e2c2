"""Drop-in for `amcpy.features` (/root/reference/src/amcpy/features.py), computed on the GPU.

Same public names, ids and call shapes as the reference:
  * `calculate_features(feature_ids, signal) -> list[float]`   (features.py:214-232)
  * `_FEATURE_FUNCTIONS` : {1..18 -> function(signal) -> float} (features.py:192-211)
  * `_gmax`, `_std_abs_phase`, ..., `_cumulant_63`              (features.py:66-185)
  * `InstantaneousValues`, `MomentValues`                        (features.py:17-31, :39-58)
plus the batched form the reference lacks: `calculate_features_batch`.

Every value comes from the CUDA library (amcpy_b200.ops); nothing here computes features on the
CPU.  A per-frame call costs one small host->device copy and one launch; use the batched entry
(or `feature_extraction.run_extraction`) for throughput.
"""

from __future__ import annotations

import numpy as np

from . import ops

FEATURE_NAMES = (
    "_gmax", "_std_abs_phase", "_std_direct_phase", "_std_abs_cna", "_std_cnf", "_mean_magnitude",
    "_norm_sqrt_sum_amp", "_kurtosis_cna", "_kurtosis_cnf", "_cumulant_20", "_cumulant_21",
    "_cumulant_40", "_cumulant_41", "_cumulant_42", "_cumulant_60", "_cumulant_61", "_cumulant_62",
    "_cumulant_63",
)


def _to_device_frame(signal):
    """One frame (any real/complex dtype, any length) -> (1, N) complex CUDA tensor."""
    import torch

    if isinstance(signal, torch.Tensor):
        t = signal
        if not t.is_complex():
            t = t.to(torch.complex128)
        return t.reshape(1, -1).cuda()
    a = np.asarray(signal)
    if a.dtype != np.complex64:
        a = a.astype(np.complex128, copy=False)
    return torch.from_numpy(np.ascontiguousarray(a).reshape(1, -1)).cuda()


def _all18(signal, feature_mask: int = ops.nat.AMC_ALL_FEATURES) -> np.ndarray:
    return ops.extract_features(_to_device_frame(signal), feature_mask=feature_mask).cpu().numpy().reshape(18)


def _make(fid: int, name: str):
    def fn(signal) -> float:
        return float(_all18(signal, 1 << (fid - 1))[fid - 1])

    fn.__name__ = name
    fn.__qualname__ = name
    fn.__doc__ = f"Feature {fid} of one frame (reference: features.py `{name}`), computed on the GPU."
    return fn


# id -> per-frame function, ids and names as in features.py:192-211
_FEATURE_FUNCTIONS = {i + 1: _make(i + 1, n) for i, n in enumerate(FEATURE_NAMES)}
globals().update({fn.__name__: fn for fn in _FEATURE_FUNCTIONS.values()})


def calculate_features(feature_ids, signal) -> list:
    """`[F[fid](signal) for fid in feature_ids]` (features.py:232): values follow the order of
    `feature_ids`; an unknown id raises KeyError.  All requested ids share one GPU launch."""
    ids = list(feature_ids)
    for fid in ids:
        _FEATURE_FUNCTIONS[fid]  # KeyError on unknown id, like the reference's dict lookup
    if not ids:
        return []
    row = _all18(signal, ops.feature_mask_of(ids))   # feature groups nobody asked for may be skipped
    return [float(row[fid - 1]) for fid in ids]


def calculate_features_batch(feature_ids, frames):
    """Batched form: frames (..., N) complex (CUDA tensor, or numpy array which is copied to the
    GPU) -> float64 (..., len(feature_ids)) of the same kind."""
    import torch

    ids = list(feature_ids)
    for fid in ids:
        _FEATURE_FUNCTIONS[fid]
    cols = [fid - 1 for fid in ids]
    mask = ops.feature_mask_of(ids) if ids else ops.nat.AMC_ALL_FEATURES
    if isinstance(frames, torch.Tensor):
        return ops.extract_features(frames, feature_mask=mask)[..., cols]
    a = np.asarray(frames)
    lead = a.shape[:-1]
    res = ops.extract_features_host(a.reshape(-1, a.shape[-1]), feature_mask=mask)
    return res[:, cols].reshape(lead + (len(cols),))


class InstantaneousValues:
    """abs / phase / unwrapped_phase / frequency / cn_amplitude of one frame (features.py:17-31)."""

    def __init__(self, signal) -> None:
        res = ops.instantaneous_batch(_to_device_frame(signal))
        self.abs = res["abs"][0].cpu().numpy()
        self.phase = res["phase"][0].cpu().numpy()
        self.unwrapped_phase = res["unwrapped_phase"][0].cpu().numpy()
        self.frequency = res["frequency"][0].cpu().numpy()
        self.cn_amplitude = res["cn_amplitude"][0].cpu().numpy()


class MomentValues:
    """Mixed moments M_pq = E[x^(p-q) conj(x)^q] of one frame (features.py:39-58);
    m21, m42, m62 are real floats exactly as in the reference."""

    def __init__(self, signal) -> None:
        m = ops.moments_batch(_to_device_frame(signal))[0].cpu().numpy()
        for name, v in zip(ops.MOMENT_NAMES, m):
            setattr(self, name, float(v.real) if name in ("m21", "m42", "m62") else complex(v))
