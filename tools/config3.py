"""BASELINE config 3 at its real size: 6 modulations x 21 SNRs x 20,000 frames x 2048 complex128 samples
(2.52 M frames, 82.6 GB) generated on the device(s) and extracted, on 1 GPU or sharded over the ranks of a
torchrun launch (every rank takes a contiguous slice of the frame axis of every (modulation, SNR) cell; the
counter-based generator makes any shard bit-identical to the same frames of the full set).
usage: python tools/config3.py            |  python -m torch.distributed.run --nproc-per-node 8 ... tools/config3.py
Prints one JSON line (rank 0): frames/s of the extraction kernel over the whole set, max over ranks, and a SHA-256 of
the gathered (2.52 M, 18) float64 matrix - it must not depend on the number of GPUs."""
import argparse
import hashlib
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from amcpy_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=20000, help="frames per (modulation, SNR) cell")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
snrs = [-20.0 + 2.0 * i for i in range(21)]
n_cells = 6 * len(snrs)
assert args.frames % world == 0
per = args.frames // world
x = synth.dataset_device(6, snrs, per, 2048, dev, seed=3, first_frame=rank * per)      # (126 * per, 2048)
out = torch.empty((x.shape[0], 18), dtype=torch.float64, device=dev)
for _ in range(2):
    ops.extract_features(x, out=out)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    ops.extract_features(x, out=out)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    parts = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
    dist.gather(out, parts, dst=0)
else:
    parts = [out]
if rank == 0:
    # (rank, cell, frame-in-shard) -> (cell, frame): the single-GPU order
    full = torch.stack([p.view(n_cells, per, 18) for p in parts], dim=1).reshape(n_cells * args.frames, 18)
    h = hashlib.sha256(full.cpu().numpy().tobytes()).hexdigest()
    frames = n_cells * args.frames
    t = float(ms.item()) * 1e-3
    print(json.dumps({"workload": f"BASELINE config 3: 6 x 21 x {args.frames} x 2048 complex128", "n_gpus": world,
                      "frames": frames, "input_GB": round(frames * 2048 * 16 / 1e9, 1), "ms": round(t * 1e3, 3),
                      "frames_per_s": round(frames / t), "GBps": round(frames * 32912 / t / 1e9, 1),
                      "finite": bool(torch.isfinite(full).all()), "sha256_of_features": h}), flush=True)
if world > 1:
    dist.destroy_process_group()
