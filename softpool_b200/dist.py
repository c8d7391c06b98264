"""Multi-GPU plumbing for the batch-sharded hot path (one process per GPU, torch.distributed).

The op is per-sample independent (reference softpool.py:140-147 indexes [:, region, :];
chamfer.cu:15 loops over the batch), so ranks shard the batch and exchange NOTHING on the data
path; the only collectives are the reporting ones here (max-over-ranks time, summed units) and,
in a training job, DDP's gradient all-reduce of the surrounding model's parameters.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    """Initialise the default process group from the torchrun environment (no-op for world 1)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                **({"device_id": torch.device("cuda", local_rank)} if backend == "nccl" else {}))
    return rank, local_rank, world


def shard_batch(global_batch, rank, world):
    """Contiguous [lo, hi) slice of the global batch owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _reduce(value, op, device):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value, device="cpu"):
    return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value, device="cpu"):
    return _reduce(value, dist.ReduceOp.SUM, device)


def gather_over_ranks(value, device="cpu"):
    """[value of rank 0, ..., value of rank world-1] on every rank (reporting only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(value)]
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
