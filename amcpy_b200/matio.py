"""Zero-copy reader for uncompressed MATLAB Level-5 .mat files (what `scipy.io.savemat` writes by default).

Why: `scipy.io.loadmat` of the reference's `mat-data/all_modulations.mat`
(/root/reference/src/amcpy/feature_extraction.py:46-48) has to interleave the separately stored real and
imaginary planes of every complex variable into one complex128 array on a single thread - 1.13 s of the
1.29 s the whole extraction stage takes on BASELINE config 1, and the reference pays it once per process (6x).
A Level-5 file stores each numeric array as one contiguous column-major real plane followed by one imaginary
plane; this module only parses the element headers and returns `numpy.memmap` views of the two planes, which
the library uploads as they are (`amc_extract_host_planar`: the interleave + transpose happen on the GPU).

Compressed elements (MATLAB's `save -v7` default, `savemat(do_compression=True)`) are one zlib stream per
top-level variable: they are inflated in parallel (zlib releases the GIL; scipy inflates and interleaves them one
after the other) and then parsed the same way, the planes being views of the inflated buffers.

Anything it does not understand (sparse / cell / struct / integer-compressed numeric data, big-endian files)
makes `read_planar` return None or skip that variable; the caller then uses `scipy.io.loadmat`.  That is an
I/O fallback - the features are computed on the GPU either way.
"""

from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

MI_INT8, MI_INT32, MI_UINT32, MI_SINGLE, MI_DOUBLE, MI_MATRIX, MI_COMPRESSED = 1, 5, 6, 7, 9, 14, 15
MX_DOUBLE, MX_SINGLE = 6, 7


@dataclass
class PlanarArray:
    """One numeric variable: column-major planes of `shape`; `im` is None for a real array."""

    shape: tuple
    re: np.ndarray
    im: np.ndarray | None
    dtype: np.dtype   # float64 or float32 (of each plane)


def _tag(buf, off):
    """(type, nbytes, data_offset, next_offset) of the data element at `off`."""
    w0, w1 = struct.unpack_from("<II", buf, off)
    if w0 >> 16:  # small data element: type in the low half, byte count in the high half, data in word 1
        return w0 & 0xFFFF, w0 >> 16, off + 4, off + 8
    return w0, w1, off + 8, off + 8 + ((w1 + 7) & ~7)


def _peek_name(mm, data: int, nbytes: int):
    """Variable name of a compressed element from the first bytes of its stream (None if it cannot be read)."""
    try:
        head = zlib.decompressobj().decompress(mm[data:data + min(nbytes, 4096)], 512)
        typ, _, d, off = _tag(head, 0)
        if typ != MI_MATRIX:
            return None
        _, _, _, off = _tag(head, d)                               # array flags
        _, _, _, off = _tag(head, off)                             # dimensions
        typ, n, d, _ = _tag(head, off)                             # name
        return bytes(head[d:d + n]).decode("latin-1") if typ == MI_INT8 else None
    except (zlib.error, struct.error):
        return None


def _inflate(mm, data: int, nbytes: int, only=None):
    """One compressed top-level element -> (name, PlanarArray) or None (views of the inflated buffer)."""
    if only is not None and _peek_name(mm, data, nbytes) not in only:
        return None
    try:
        raw = zlib.decompress(mm[data:data + nbytes])
    except zlib.error:
        return None
    buf = np.frombuffer(raw, dtype=np.uint8)
    if buf.size < 8:
        return None
    typ, n, d, _ = _tag(buf, 0)
    if typ != MI_MATRIX or d + n > buf.size:
        return None
    return _matrix(buf, d, d + n)


def read_planar(path, max_workers: int = 8, only=None) -> dict | None:
    """{name: PlanarArray} for every double/single array of the file (other variables are skipped), or None
    when the file is not a little-endian Level-5 file.  `only`: names to keep (compressed variables outside
    it are not inflated - a rank of a multi-GPU run reads its own modulations only)."""
    if only is not None:
        only = set(only)
    mm = np.memmap(str(path), dtype=np.uint8, mode="r")
    if mm.size < 128 or bytes(mm[126:128]) != b"IM" or bytes(mm[:4]) == b"\x00\x00\x00\x00":
        return None
    out = {}
    compressed = []
    off, end = 128, mm.size
    while off + 8 <= end:
        typ, nbytes, data, nxt = _tag(mm, off)
        if typ == MI_COMPRESSED:
            compressed.append((data, nbytes))
            nxt = data + nbytes                      # compressed elements are not padded to 8 bytes
        elif typ == MI_MATRIX and nbytes > 0:
            item = _matrix(mm, data, data + nbytes)
            if item is not None and (only is None or item[0] in only):
                out[item[0]] = item[1]
        if nxt <= off:
            return None
        off = nxt
    if compressed:
        with ThreadPoolExecutor(max_workers=max(1, min(max_workers, len(compressed)))) as pool:
            for item in pool.map(lambda c: _inflate(mm, *c, only=only), compressed):
                if item is not None:
                    out[item[0]] = item[1]
    return out


def _matrix(mm, off, end):
    typ, nbytes, data, off = _tag(mm, off)                     # array flags
    if typ != MI_UINT32 or nbytes != 8:
        return None
    flags = struct.unpack_from("<I", mm, data)[0]
    cls, is_complex = flags & 0xFF, bool(flags & 0x0800)
    typ, nbytes, data, off = _tag(mm, off)                     # dimensions
    if typ != MI_INT32:
        return None
    shape = tuple(int(v) for v in np.frombuffer(mm, dtype="<i4", count=nbytes // 4, offset=data))
    typ, nbytes, data, off = _tag(mm, off)                     # name
    if typ != MI_INT8:
        return None
    name = bytes(mm[data:data + nbytes]).decode("latin-1")
    if cls not in (MX_DOUBLE, MX_SINGLE):
        return None
    want, dt = (MI_DOUBLE, np.dtype("<f8")) if cls == MX_DOUBLE else (MI_SINGLE, np.dtype("<f4"))
    count = int(np.prod(shape)) if shape else 0
    planes = []
    for _ in range(2 if is_complex else 1):
        if off + 8 > end:
            return None
        typ, nbytes, data, off = _tag(mm, off)
        if typ != want or nbytes != count * dt.itemsize or data % dt.itemsize:
            return None                                        # e.g. MATLAB's integer-compressed numeric data
        if data + nbytes > end:
            return None
        planes.append(mm[data:data + nbytes].view(dt))
    return name, PlanarArray(shape, planes[0], planes[1] if is_complex else None, dt)
