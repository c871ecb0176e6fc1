"""One-process-per-GPU helpers (torch.distributed / NCCL is plumbing only).

Accumulation shards examples and needs exactly one all-reduce of the factors (done inside
``Laplace.infer(shard=...)``).  The predictive shards TEST INPUTS: the posterior is replicated,
each rank evaluates a contiguous block, and the only exchange is the gather of the outputs.
"""
import torch


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous block [lo, hi) of n items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_predictive(fn, X, n_samples, gather=True, **kw):
    """Run ``fn(X_local, n_samples, **kw) -> [S, n_local, K]`` on this rank's block of X and
    (optionally) all-gather the blocks into [S, N, K] on every rank."""
    import torch.distributed as dist
    rank, ws = world()
    lo, hi = shard_bounds(len(X), rank, ws)
    out = fn(X[lo:hi], n_samples, **kw)
    if ws == 1 or not gather:
        return out
    sizes = [shard_bounds(len(X), r, ws) for r in range(ws)]
    parts = [torch.empty(out.shape[0], b - a, out.shape[2], dtype=out.dtype, device=out.device)
             for a, b in sizes]
    dist.all_gather(parts, out.contiguous())
    return torch.cat(parts, dim=1)
