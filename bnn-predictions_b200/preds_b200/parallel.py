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
    # all_gather needs equal shapes: pad every block to the largest one, trim after the exchange
    sizes = [b - a for a, b in (shard_bounds(len(X), r, ws) for r in range(ws))]
    width = max(sizes)
    mine = out.new_zeros(out.shape[0], width, out.shape[2])
    mine[:, :out.shape[1]] = out
    parts = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(parts, mine)
    return torch.cat([p[:, :n] for p, n in zip(parts, sizes)], dim=1)


# ----------------------------------------------------------------------------------------------
# accumulation side: which loader batches a rank owns and the single exchange step.  These helpers
# are device-agnostic (they are exercised with the gloo backend on CPU tensors in the tests and
# with NCCL on the GPU box).
# ----------------------------------------------------------------------------------------------
def owns_batch(i, shard, rank, world_size):
    """True if this rank processes loader batch i under the given sharding mode."""
    return not (shard == 'batches' and world_size > 1 and i % world_size != rank)


def global_count(n_local, device, enabled=True):
    """Sum of the per-rank example counts (the N of the KFAC batch weight M/N)."""
    import torch.distributed as dist
    if not enabled or world()[1] == 1:
        return int(n_local)
    cnt = torch.tensor([float(n_local)], device=device, dtype=torch.float64)
    dist.all_reduce(cnt)
    return int(round(cnt.item()))


def reduce_sum_(acc, prior=None):
    """ONE all-reduce(sum) of a rank-local accumulator; the prior precision is added once, after
    the reduce (never per rank)."""
    import torch.distributed as dist
    if world()[1] > 1:
        dist.all_reduce(acc)
    if prior is not None:
        acc += prior
    return acc


def reduce_factors_(qhs):
    """All Kronecker factors of all parameters in ONE flat bucket (one launch, one all-reduce)."""
    import torch.distributed as dist
    if world()[1] == 1:
        return qhs
    flat = torch.cat([m.reshape(-1) for qh in qhs for m in qh])
    dist.all_reduce(flat)
    off = 0
    for qh in qhs:
        for m in qh:
            m.copy_(flat[off:off + m.numel()].reshape(m.shape))
            off += m.numel()
    return qhs
