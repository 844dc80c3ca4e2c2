"""Multi-GPU sharding of the extraction path.

Frames are independent units (each frame's 18 outputs depend on that frame only -
/root/reference/src/amcpy/features.py:214-232; the reference already treats (modulation) and
(snr, frame) as independent work items - feature_extraction.py:64-72, :89-92), so the flattened
(modulation, snr, frame) index space is cut into one contiguous range per rank and there is NO
collective on the data path.  The only exchange is the optional gather of the (frames, 18) feature
matrix to every rank / rank 0 for the downstream classifier:
  * `gather_features`          - NCCL `all_gather_into_tensor` over NVLink (gloo in the CPU tests);
  * `extract_sharded_to_root`  - no collective call at all: every rank's extraction kernel stores its
    144-byte rows straight into the root's matrix through the NVLink/NVSwitch peer mapping of a
    symmetric-memory buffer, so the gather rides on the kernel's own epilogue.
"""

from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the contiguous shard of `total` equal-cost units owned by `rank` of `world`
    (sizes differ by at most one; empty shards are legal when world > total)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def unflatten(index: int, n_snr: int, n_frames: int) -> tuple[int, int, int]:
    """Flat unit index -> (modulation, snr, frame) with frame fastest."""
    mod, rest = divmod(index, n_snr * n_frames)
    snr, frame = divmod(rest, n_frames)
    return mod, snr, frame


def gather_features(local, total: int, group=None):
    """All ranks receive the full (total, 18) matrix assembled from contiguous shards.

    `local` is this rank's (hi - lo, 18) float64 tensor (CUDA under NCCL, CPU under gloo).  Shards
    may differ by one row, so they are padded to the largest shard for `all_gather_into_tensor`."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(total, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: local has {local.shape[0]} rows, shard is [{lo},{hi})")
    width = local.shape[1]
    per = -(-total // world)
    padded = torch.zeros((per, width), dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    gathered = torch.empty((world * per, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    out = torch.empty((total, width), dtype=local.dtype, device=local.device)
    for r in range(world):
        rlo, rhi = shard_range(total, r, world)
        out[rlo:rhi] = gathered[r * per : r * per + (rhi - rlo)]
    return out


def extract_sharded(frames_local, total: int, gather: bool = True, group=None):
    """Features of this rank's shard (device tensor (hi-lo, N)); optionally gathered to all ranks."""
    from . import ops

    local = ops.extract_features(frames_local)
    return gather_features(local, total, group) if gather else local


_SYMM_CACHE: dict = {}


def extract_sharded_to_root(frames_local, total: int, root: int = 0, group=None):
    """Features of this rank's shard, written by the extraction kernel DIRECTLY into `root`'s
    (total, 18) matrix over NVLink (peer stores into a torch symmetric-memory buffer).

    The kernel's finalisation stores each frame's 144-byte row to `out`; here `out` is the peer
    mapping of the root's buffer at this rank's row offset, so compute and gather are one kernel and
    the only extra work is one device-side barrier.  Returns the (total, 18) float64 tensor on
    `root` (valid after the call in stream order) and None on the other ranks.  NCCL/CUDA only."""
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm

    from . import ops

    group = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(total, rank, world)
    if frames_local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: {frames_local.shape[0]} local frames, shard is [{lo},{hi})")
    key = (total, frames_local.device.index, group.group_name)
    if key not in _SYMM_CACHE:   # allocation + rendezvous are collective and slow: once per shape
        buf = symm.empty((total, ops.N_FEATURES), dtype=torch.float64, device=frames_local.device)
        _SYMM_CACHE[key] = (buf, symm.rendezvous(buf, group))
    buf, hdl = _SYMM_CACHE[key]
    dst = hdl.get_buffer(root, (total, ops.N_FEATURES), torch.float64)   # root's buffer, peer-mapped here
    hdl.barrier()                    # the root is done reading the previous result
    if hi > lo:
        ops.extract_features(frames_local, out=dst[lo:hi])
    hdl.barrier()                    # every rank's rows have landed in the root's memory
    return buf if rank == root else None
