"""Multi-GPU sharding of the extraction path.

Frames are independent units (each frame's 18 outputs depend on that frame only -
/root/reference/src/amcpy/features.py:214-232; the reference already treats (modulation) and
(snr, frame) as independent work items - feature_extraction.py:64-72, :89-92), so the flattened
(modulation, snr, frame) index space is cut into one contiguous range per rank and there is NO
collective on the data path.  The only exchange is the optional gather of the (frames, 18) feature
matrix to every rank / rank 0 for the downstream classifier (NCCL over NVLink on GPUs, gloo in the
CPU tests).
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
