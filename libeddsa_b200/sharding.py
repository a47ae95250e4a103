"""Host-side plumbing for the one-process-per-GPU launch (torchrun): contiguous index shards and the
max-over-ranks reduction used for every reported time.  No data-path collective exists — operations
are independent, so a rank only ever touches its own shard.  Works on any
torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Items [lo, hi) of a batch of n that rank `rank` of `world` processes — the same contiguous
    partition host.c uses across devices (run_job: lo = n*g/G, hi = n*(g+1)/G)."""
    if not 0 <= rank < world:
        raise ValueError("rank outside world")
    return n * rank // world, n * (rank + 1) // world


def max_over_ranks(value, device=None):
    """Largest `value` over all ranks (every reported multi-GPU time is the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
