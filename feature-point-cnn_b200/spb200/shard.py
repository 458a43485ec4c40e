"""Batch sharding across the GPUs of one box: images are independent, so rank r of `world`
takes a contiguous slice of the batch and nothing is exchanged on the data path.  The only collective a multi-GPU
run needs is the MAX of the per-rank step time for reporting; it goes through torch.distributed (NCCL on the GPU
box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [lo, hi) slice of `total` images for `rank`; sizes differ by at most one, earlier ranks larger."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError('bad rank/world')
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(images, rank, world):
    """The slice of a B x ... tensor (or array) this rank processes."""
    lo, hi = shard_range(len(images), rank, world)
    return images[lo:hi]


def max_over_ranks(value, device='cpu'):
    """MAX of a per-rank scalar (step time) over the process group; identity without a group."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(units_per_rank, seconds, device='cpu'):
    """units all ranks processed / max-over-ranks time (the bench contract's `value`)."""
    total = torch.tensor([float(units_per_rank)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return float(total.item()) / max_over_ranks(seconds, device)
