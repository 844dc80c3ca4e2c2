"""Batched operators over the C ABI: the reference's per-frame operator
(/root/reference/src/amcpy/features.py:214-232) applied to whole (n_snr, n_frames, frame_size)
tensors at once.  torch is used for device memory and streams only."""

from __future__ import annotations

from contextlib import nullcontext as _nullcontext

import numpy as np

from . import _native as nat

N_FEATURES = 18
FUSED_SIZES = (256, 512, 1024, 2048, 4096, 8192, 16384)   # frame sizes with a fused kernel (amc_api.cu: fused_size)
MOMENT_NAMES = ("m20", "m21", "m22", "m40", "m41", "m42", "m43", "m60", "m61", "m62", "m63")


def _torch():
    import torch

    return torch


def _dtype_code(t) -> int:
    torch = _torch()
    if t.dtype == torch.complex128:
        return nat.AMC_C128
    if t.dtype == torch.complex64:
        return nat.AMC_C64
    raise TypeError(f"expected a complex64/complex128 tensor, got {t.dtype}")


def _as_frames_2d(iq):
    """(..., N) CUDA tensor -> 2-D view (frames, N) without copying when the leading dims collapse."""
    torch = _torch()
    if not iq.is_cuda:
        raise ValueError("extract_features needs a CUDA tensor (use extract_features_host for host arrays)")
    if iq.dim() == 1:
        iq = iq.unsqueeze(0)
    n = iq.shape[-1]
    if iq.dim() == 2:
        return iq
    try:
        return iq.view(-1, n)
    except RuntimeError:
        return iq.reshape(-1, n)  # copies only when the leading dims are not collapsible


def _stream_ptr(stream):
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def feature_mask_of(feature_ids) -> int:
    """C-ABI feature_mask (bit k = feature id k+1) of a list of ids 1..18."""
    mask = 0
    for fid in feature_ids:
        if not 1 <= int(fid) <= N_FEATURES:
            raise KeyError(fid)
        mask |= 1 << (int(fid) - 1)
    return mask


def extract_features(iq, out=None, stream=None, force_general: bool = False, feature_mask: int = nat.AMC_ALL_FEATURES,
                     direct_dft: bool = False, relayout: bool = True, extra_flags: int = 0):
    """All 18 features of every frame of a device-resident complex tensor.

    feature_mask (default: all 18): the features the caller will read.  The library may skip the work of
    feature groups outside the mask (FFT / phase+frequency / amplitude / moments; frame sizes 256..4096); the
    columns of a skipped group hold NaN, requested columns are bitwise what the full call returns.

    iq  : CUDA tensor (..., frame_size), complex128 or complex64 (north_star's batched entry takes
          (n_snr, n_frames, frame_size)).
    out : optional CUDA float64 tensor (..., 18), C-contiguous.
    extra_flags: raw C-ABI flag bits OR-ed in (A/B experiment kernels of a -DAMC_EXPERIMENTS build only).
    relayout: sample-major device tensors of a fused frame size are re-laid-out into a scratch tensor first
          (False: the general kernel reads them in place).
    direct_dft: cross-check switch - frame sizes that are not powers of two use the float64 direct DFT (O(N^2))
          instead of the float32 Bluestein FFT of the general kernel.
    Returns float64 (..., 18); column k = feature id k+1.  Enqueued on `stream`
    (default: torch's current stream); does not synchronise.
    """
    torch = _torch()
    lead = tuple(iq.shape[:-1]) if iq.dim() > 1 else (1,)
    x = _as_frames_2d(iq)
    n_frames, n = x.shape
    if (relayout and not force_general and n in FUSED_SIZES and n_frames > 1 and x.stride(0) == 1
            and x.stride(1) >= n_frames):
        # sample-major on the device (a loadmat-ordered block): one pass of the library's re-layout kernel, then the
        # fused kernel, instead of the general kernel in place (17x slower); costs a scratch copy of the batch
        x = frames_from_sample_major(x, n_frames, n, x.stride(1), stream=stream)
    if out is None:
        # allocated on the stream the kernels run on: the caching allocator then orders any reuse of the block
        # behind them (a block from another stream could be handed out again while the kernels are still queued)
        with torch.cuda.stream(stream) if stream is not None else _nullcontext():
            out = torch.empty((n_frames, N_FEATURES), dtype=torch.float64, device=x.device)
    else:
        if out.dtype != torch.float64 or not out.is_contiguous() or out.numel() != n_frames * N_FEATURES:
            raise ValueError("out must be a contiguous float64 tensor with 18 values per frame")
        if stream is not None:
            out.record_stream(stream)
    if stream is not None:
        x.record_stream(stream)          # (covers the re-layout scratch tensor, which is dropped on return)
    with torch.cuda.device(x.device):
        rc = nat.lib().amc_extract_batch(
            x.data_ptr(), _dtype_code(x), n_frames, n, x.stride(0) if n_frames > 1 else n, x.stride(1) if n > 1 else 1,
            out.data_ptr(), N_FEATURES, feature_mask,
            (nat.AMC_FLAG_FORCE_GENERAL if force_general else 0) | (nat.AMC_FLAG_DIRECT_DFT if direct_dft else 0)
            | int(extra_flags),
            _stream_ptr(stream),
        )
    nat.check(rc)
    return out.view(*lead, N_FEATURES) if iq.dim() > 1 else out.view(N_FEATURES)


def _np_dtype_code(a: np.ndarray) -> int:
    if a.dtype == np.complex128:
        return nat.AMC_C128
    if a.dtype == np.complex64:
        return nat.AMC_C64
    raise TypeError(f"expected complex64/complex128, got {a.dtype}")


def extract_features_host(frames: np.ndarray, device: int = 0, out: np.ndarray | None = None,
                          force_general: bool = False, feature_mask: int = nat.AMC_ALL_FEATURES) -> np.ndarray:
    """Host arrays through the library's chunked copy/compute pipeline (amc_extract_host).

    frames: (n_frames, N) complex array.  Row-per-frame (C order, rows may be padded) and
    sample-major (Fortran order - what scipy.io.loadmat returns) are consumed in place; any other
    striding is made C-contiguous first.  Returns float64 (n_frames, 18)."""
    a = np.asarray(frames)
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2:
        raise ValueError("frames must be 2-D (n_frames, frame_size)")
    if not np.iscomplexobj(a):
        a = a.astype(np.complex128)
    code = _np_dtype_code(a)
    nf, n = a.shape
    es = a.itemsize
    s0, s1 = a.strides[0] // es, a.strides[1] // es
    ok_row = (n == 1 or s1 == 1) and (nf == 1 or s0 >= n) and a.strides[0] % es == 0
    ok_col = (nf == 1 or s0 == 1) and (n == 1 or s1 >= nf) and a.strides[1] % es == 0 and not ok_row
    if not (ok_row or ok_col):
        a = np.ascontiguousarray(a)
        s0, s1, ok_row = n, 1, True
    if ok_row:
        s0, s1 = (s0 if nf > 1 else n), 1
    else:
        s0, s1 = 1, (s1 if n > 1 else nf)
    if out is None:
        out = np.empty((nf, N_FEATURES), dtype=np.float64)
    elif out.dtype != np.float64 or out.shape != (nf, N_FEATURES) or not out.flags.c_contiguous:
        raise ValueError("out must be C-contiguous float64 (n_frames, 18)")
    rc = nat.lib().amc_extract_host(
        a.ctypes.data, code, nf, n, s0, s1, out.ctypes.data, N_FEATURES, feature_mask,
        nat.AMC_FLAG_FORCE_GENERAL if force_general else 0, device,
    )
    nat.check(rc)
    return out


def extract_features_host_planar(re: np.ndarray, im: np.ndarray | None, n_frames: int, frame_size: int,
                                 sample_stride: int, device: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    """Planar sample-major host data (amc_extract_host_planar): `re` / `im` are flat float64 or float32 planes
    (e.g. numpy.memmap views of a .mat file, see matio.py), plane element (f, n) at f + n*sample_stride.
    Returns float64 (n_frames, 18)."""
    re = np.asarray(re)
    if re.dtype not in (np.float64, np.float32) or re.ndim != 1 or not re.flags.c_contiguous:
        raise TypeError("re must be a flat contiguous float64/float32 array")
    if im is not None:
        im = np.asarray(im)
        if im.dtype != re.dtype or im.shape != re.shape or not im.flags.c_contiguous:
            raise TypeError("im must match re")
    need = (frame_size - 1) * sample_stride + n_frames
    if n_frames > 0 and re.size < need:
        raise ValueError(f"planes hold {re.size} elements, layout needs {need}")
    if out is None:
        out = np.empty((n_frames, N_FEATURES), dtype=np.float64)
    elif out.dtype != np.float64 or out.shape != (n_frames, N_FEATURES) or not out.flags.c_contiguous:
        raise ValueError("out must be C-contiguous float64 (n_frames, 18)")
    rc = nat.lib().amc_extract_host_planar(
        re.ctypes.data, im.ctypes.data if im is not None else None,
        nat.AMC_C128 if re.dtype == np.float64 else nat.AMC_C64, n_frames, frame_size, sample_stride,
        out.ctypes.data, N_FEATURES, nat.AMC_ALL_FEATURES, 0, device,
    )
    nat.check(rc)
    return out


def frames_from_sample_major(src, n_frames: int, frame_size: int, sample_stride: int, stream=None):
    """Device re-layout of a flat sample-major block (element (f, n) at f + n*sample_stride) into
    a (n_frames, frame_size) C-contiguous tensor."""
    torch = _torch()
    with torch.cuda.stream(stream) if stream is not None else _nullcontext():
        dst = torch.empty((n_frames, frame_size), dtype=src.dtype, device=src.device)
    if stream is not None:
        src.record_stream(stream)
    with torch.cuda.device(src.device):
        rc = nat.lib().amc_frames_from_sample_major(
            src.data_ptr(), _dtype_code(src), n_frames, frame_size, sample_stride, dst.data_ptr(), _stream_ptr(stream)
        )
    nat.check(rc)
    return dst


def instantaneous_batch(iq, stream=None):
    """dict of float64 CUDA tensors: abs, phase, unwrapped_phase, cn_amplitude (frames, N) and
    frequency (frames, N-1) - the arrays of the reference's InstantaneousValues (features.py:27-31)."""
    torch = _torch()
    x = _as_frames_2d(iq)
    nf, n = x.shape
    mk = lambda m: torch.empty((nf, m), dtype=torch.float64, device=x.device)  # noqa: E731
    res = {"abs": mk(n), "phase": mk(n), "unwrapped_phase": mk(n), "frequency": mk(max(n - 1, 0)), "cn_amplitude": mk(n)}
    with torch.cuda.device(x.device):
        rc = nat.lib().amc_instantaneous_batch(
            x.data_ptr(), _dtype_code(x), nf, n, x.stride(0) if nf > 1 else n, x.stride(1) if n > 1 else 1,
            res["abs"].data_ptr(), res["phase"].data_ptr(), res["unwrapped_phase"].data_ptr(),
            res["frequency"].data_ptr() if n > 1 else None, res["cn_amplitude"].data_ptr(), _stream_ptr(stream),
        )
    nat.check(rc)
    return res


def moments_batch(iq, stream=None):
    """complex128 CUDA tensor (frames, 11): m20 m21 m22 m40 m41 m42 m43 m60 m61 m62 m63
    (features.py:46-58; m21, m42, m62 carry a zero imaginary part)."""
    torch = _torch()
    x = _as_frames_2d(iq)
    nf, n = x.shape
    out = torch.empty((nf, 22), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = nat.lib().amc_moments_batch(
            x.data_ptr(), _dtype_code(x), nf, n, x.stride(0) if nf > 1 else n, x.stride(1) if n > 1 else 1,
            out.data_ptr(), _stream_ptr(stream),
        )
    nat.check(rc)
    return torch.view_as_complex(out.view(nf, 11, 2))
