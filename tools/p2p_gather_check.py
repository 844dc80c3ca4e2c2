"""2+ GPU check of sharding.extract_sharded_to_root (kernel epilogue stores into the root's matrix over
NVLink peer memory) against the NCCL gather and the single-GPU result; also times both.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_gather_check.py"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from amcpy_b200 import ops, sharding, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
snrs = [-10.0 + 2.0 * i for i in range(16)]
full = synth.dataset_device(6, snrs, 50, 2048, dev, seed=7)          # every rank regenerates the same 4800 frames
total = full.shape[0]
lo, hi = sharding.shard_range(total, rank, world)
mine = full[lo:hi].contiguous()
want = ops.extract_features(full)                                    # single-GPU result of the whole set
got_nccl = sharding.gather_features(ops.extract_features(mine), total)
got_p2p = sharding.extract_sharded_to_root(mine, total, root=0)
torch.cuda.synchronize()
ok_nccl = bool(torch.equal(got_nccl, want))
ok_p2p = bool(torch.equal(got_p2p, want)) if rank == 0 else True


def timed(fn, n=20):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


t_local = timed(lambda: ops.extract_features(mine))
t_nccl = timed(lambda: sharding.gather_features(ops.extract_features(mine), total))
t_p2p = timed(lambda: sharding.extract_sharded_to_root(mine, total, root=0))
flags = torch.tensor([int(ok_nccl), int(ok_p2p)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: nccl gather bitwise == single GPU: {bool(flags[0])}; peer-store gather bitwise == single GPU: "
          f"{bool(flags[1])}; ms per call: extract only {t_local:.3f}, + NCCL all_gather {t_nccl:.3f}, "
          f"fused peer stores to root {t_p2p:.3f}")
dist.destroy_process_group()
sys.exit(0 if bool(flags.min()) else 1)
