"""ctypes binding of the C-ABI library (include/amcpy_b200.h) and its in-tree build.

The library is built by nvcc straight into amcpy_b200/_lib/ (git-ignored, ships to the GPU box
with the snapshot).  There is NO CPU fallback: if the library is missing or no CUDA device is
visible, every compute entry raises.
"""

from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "_lib"
LIB_PATH = Path(os.environ.get("AMCPY_B200_LIB", LIB_DIR / "libamcpy_b200.so"))   # override: A/B builds only
HEADER = PKG.parent / "include" / "amcpy_b200.h"

AMC_C64, AMC_C128 = 0, 1
AMC_FLAG_FORCE_GENERAL = 1
AMC_FLAG_DIRECT_DFT = 8
AMC_FLAG_FUSED_SPT8 = 2   # only understood by a library built with -DAMC_EXPERIMENTS (tools/exp/build_variants.py)
AMC_FLAG_FUSED_WS = 4     # idem
AMC_ALL_FEATURES = 0x3FFFF

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
]


class AmcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"amcpy_b200 native error {code}: {msg}")
        self.code = code


def _sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.rglob("*.cuh")) + [HEADER]


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    return any(s.stat().st_mtime > built for s in _sources())


def build(force: bool = False, verbose: bool = False, defines=(), out: Path | None = None) -> Path:
    """Compile csrc/*.cu for sm_100a into _lib/libamcpy_b200.so (nvcc cross-compiles without a GPU).
    `defines` / `out`: A/B builds only (e.g. defines=("AMC_EXPERIMENTS",), out=_lib/exp/libamcpy_b200_exp.so,
    selected at run time with AMCPY_B200_LIB)."""
    if out is not None:
        return _compile(Path(out), verbose, defines)
    if not force and not needs_build():
        return LIB_PATH
    path = _compile(LIB_PATH, verbose, defines)
    global _LIB
    _LIB = None
    return path


def _compile(target: Path, verbose: bool, defines) -> Path:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build the amcpy_b200 CUDA library")
    target.parent.mkdir(parents=True, exist_ok=True)
    tmp = target.parent / f".{target.stem}.{os.getpid()}.so"
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", str(tmp)] + [str(p) for p in sorted(CSRC.glob("*.cu"))]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    os.replace(tmp, target)
    return target


def build_if_missing() -> Path:
    """Build only when the library file is absent (a checkout that received sources but no build products).
    Unlike `build()` it never rebuilds a library that merely looks older than a source file."""
    if os.environ.get("AMCPY_B200_LIB") or LIB_PATH.exists():
        return LIB_PATH
    return build(force=True)


_LIB = None

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_INT = ctypes.c_int
_U32 = ctypes.c_uint32

# name -> (restype, argtypes); must list every symbol include/amcpy_b200.h declares
SIGNATURES = {
    "amc_version": (_INT, []),
    "amc_init": (_INT, [_INT]),
    "amc_workspace_bytes": (_I64, [_INT, _I64, _I64, _INT]),
    "amc_last_error_string": (ctypes.c_char_p, []),
    "amc_device_count": (_INT, []),
    "amc_launch_count": (_I64, []),
    "amc_extract_batch": (_INT, [_P, _INT, _I64, _I64, _I64, _I64, _P, _I64, _U32, _INT, _P]),
    "amc_extract_host": (_INT, [_P, _INT, _I64, _I64, _I64, _I64, _P, _I64, _U32, _INT, _INT]),
    "amc_extract_host_planar": (_INT, [_P, _P, _INT, _I64, _I64, _I64, _P, _I64, _U32, _INT, _INT]),
    "amc_frames_from_sample_major": (_INT, [_P, _INT, _I64, _I64, _I64, _P, _P]),
    "amc_instantaneous_batch": (_INT, [_P, _INT, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P]),
    "amc_moments_batch": (_INT, [_P, _INT, _I64, _I64, _I64, _I64, _P, _P]),
    "amc_generate_frames": (_INT, [_P, _INT, _INT, _I64, _I64, _I64, _P, _P, _P, ctypes.c_uint64, _P]),
}


def lib() -> ctypes.CDLL:
    """Load the library (never builds implicitly; fails loudly when it is missing)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback."
        )
    handle = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = handle
    return handle


def check(code: int) -> None:
    if code < 0:
        raise AmcError(code, lib().amc_last_error_string().decode("utf-8", "replace"))


def require_cuda() -> int:
    """Number of devices; raises (no silent fallback) when there is none."""
    n = lib().amc_device_count()
    check(n)
    return n


def bind_host_thread_to_gpu(device_index: int) -> list[int]:
    """Pin the calling process to the CPUs NVML reports as local to `device_index` (its NUMA node), so
    that pinned staging buffers allocated afterwards are first-touched next to that GPU's PCIe root.
    With one rank per GPU on a two-socket host this is what keeps N ranks from sharing one socket's
    memory controllers during host -> device streaming.  Best effort: returns the CPU list it set,
    or [] when NVML / affinity are unavailable (nothing is changed then)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return []
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001 - affinity is an optimisation, never a requirement
        return []
