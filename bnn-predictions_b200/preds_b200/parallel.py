"""One-process-per-GPU helpers (torch.distributed / NCCL is plumbing only).

Accumulation shards examples and needs exactly one exchange of the factors (done inside
``Laplace.infer(shard=...)``).  The predictive shards TEST INPUTS: the posterior is replicated,
each rank evaluates a contiguous block, and the only exchange is the gather of the outputs.

The full-covariance exchange (``LowerExchange``) is the library's own kernel over NVLink peer memory
(``csrc/exchange.cu``): the accumulator lives in a symmetric allocation, only its lower triangle is exchanged, row
chunk by row chunk on a side stream while the tensor-core kernel still accumulates later chunks.
"""
import ctypes as C

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


def reduce_factors_(qhs, timing=None):
    """All Kronecker factors of all parameters in ONE flat bucket (one launch, one all-reduce).  ``timing``: optional
    dict that receives CUDA events around the collective ('e0', 'e1') and the bucket size in bytes."""
    import torch.distributed as dist
    if world()[1] == 1:
        return qhs
    flat = torch.cat([m.reshape(-1) for qh in qhs for m in qh])
    if timing is not None:
        timing['e0'], timing['e1'] = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        timing['bytes'] = flat.numel() * 4
        timing['e0'].record()
    dist.all_reduce(flat)
    if timing is not None:
        timing['e1'].record()
    off = 0
    for qh in qhs:
        for m in qh:
            m.copy_(flat[off:off + m.numel()].reshape(m.shape))
            off += m.numel()
    return qhs


# ----------------------------------------------------------------------------------------------
# full covariance: exchange of the lower triangle of H over peer memory (csrc/exchange.cu)
# ----------------------------------------------------------------------------------------------
EXCHANGE_CHUNKS = 6          # row chunks of the last super-batch (equal triangle area); the last one is not overlapped
EXCHANGE_CTAS = 0            # 0: the kernel's default (NVLS 16, P2P 64): few SMs, the rest keep accumulating


_SYMMETRIC_POOL = {}         # (P, device index) -> [(tensor, handle)] handed back by Laplace.release_posterior()


def symmetric_release(t, hdl):
    """Hand a symmetric accumulator back for the next sharded infer of the same size (allocation + rendezvous of a
    14 GB symmetric buffer costs far more than zeroing it).  Only the owner may call this, once it holds no other
    reference to ``t``."""
    _SYMMETRIC_POOL.setdefault((t.shape[0], t.device.index), []).append((t, hdl))


def symmetric_zeros(P, device):
    """([P,P] fp32 zeros in symmetric memory, rendezvous handle) over the default group, or (None, None) when this
    torch / system cannot provide peer mappings (then the packed NCCL exchange is used).  Collective: every rank must
    call it (and ``symmetric_release``) in the same order."""
    import torch.distributed as dist
    from ._cabi import lib, check, ptr, stream
    pool = _SYMMETRIC_POOL.get((P, torch.device(device).index))
    if pool:
        t, hdl = pool.pop()
        check(lib.bnnp_zero_lower(ptr(t), P, stream()), 'bnnp_zero_lower')     # the upper triangle is rewritten by the mirror
        return t, hdl
    try:
        import warnings
        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD
        try:
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                symm.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        t = symm.empty((P, P), dtype=torch.float32, device=device)
        hdl = symm.rendezvous(t, group.group_name)
        t.zero_()
        return t, hdl
    except Exception as exc:                       # no peer access / no VMM export in this environment
        import logging
        logging.warning('symmetric memory unavailable (%s): full-covariance exchange falls back to packed NCCL', exc)
        return None, None


class LowerExchange:
    """Sum the lower triangle of the rank-local accumulators ``H`` ([P,P], row-major) in place, chunk of rows by chunk
    of rows, and add the prior precision (scalar or [P]) to the diagonal once.

    mode 'nvls'  own kernel, sum formed inside NVSwitch (multimem.ld_reduce / multimem.st) -- needs ``hdl`` with a
                 multicast mapping;
         'p2p'   own kernel over the peer mappings of ``hdl`` (loads from every peer in rank order, stores to every peer);
         'nccl'  pack the rows' triangle -> ``dist.all_reduce`` -> unpack (half the bytes of the square).
    ``rows_done(r0, r1)`` is called on the accumulation stream as soon as rows [r0, r1) are final on this rank; the
    exchange of those rows runs on a side stream.  ``finish()`` joins the side stream.  ``stats`` holds CUDA events
    around every chunk's exchange (bench.py reports allreduce_ms and the bus bandwidth from them)."""

    def __init__(self, H, hdl, prior, mode='auto'):
        import torch.distributed as dist
        from ._cabi import BnnpError
        self.H, self.hdl, self.P = H, hdl, H.shape[0]
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if mode == 'auto':
            # bytes per GPU and link direction: NVLS (1 + 1/W) x triangle (every copy is read through the switch, own copy
            # included), P2P 2 (W-1)/W x triangle -> NVLS from W = 4 (measured at W = 2: NVLS 20.3 ms, P2P 11.8 ms)
            nvls = hdl is not None and int(hdl.multicast_ptr) != 0 and self.world >= 4
            mode = 'nccl' if hdl is None else ('nvls' if nvls else 'p2p')
        if mode in ('nvls', 'p2p') and hdl is None:
            raise BnnpError("exchange mode '%s' needs the accumulator in symmetric memory" % mode)
        if mode == 'nvls' and int(hdl.multicast_ptr) == 0:
            raise BnnpError("exchange mode 'nvls' needs a multicast mapping (NVSwitch multicast is unavailable here)")
        if mode not in ('nvls', 'p2p', 'nccl'):
            raise ValueError("exchange must be 'auto', 'nvls', 'p2p' or 'nccl'")
        self.mode = mode
        self.prior_vec, self.prior_scalar = None, 0.0
        if torch.is_tensor(prior) and prior.dim() == 1:
            self.prior_vec = prior.to(device=H.device, dtype=torch.float32).contiguous()
        elif prior is not None:
            self.prior_scalar = float(prior)
        self.comm = torch.cuda.Stream(device=H.device)
        self.events = []
        self._unmirrored = None      # rows exchanged but not mirrored yet (their broadcasts land by the next barrier)
        self.mirrors = True          # the exchange also mirrors the lower triangle into the upper one, chunk by chunk
        if hdl is not None:
            off = H.data_ptr() - int(hdl.buffer_ptrs[self.rank])
            self.peers = (C.c_void_p * self.world)(*[int(b) + off for b in hdl.buffer_ptrs])
            self.mc = int(hdl.multicast_ptr) + off if int(hdl.multicast_ptr) else None

    def agree(self, can_chunk):
        """True iff EVERY rank can exchange row chunk by row chunk (the chunk boundaries depend on the model only, but a
        rank without examples, or one whose last super-batch runs on the SIMT engine, accumulates H in one piece).  The
        number of ``rows_done`` calls must be the same on every rank: each one contains a cross-rank barrier."""
        import torch.distributed as dist
        flag = torch.tensor([1 if can_chunk else 0], device=self.H.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    def rows_done(self, r0, r1):
        import torch.distributed as dist
        from ._cabi import lib, check, ptr
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ready)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if self.mode == 'nccl':
                n = int(lib.bnnp_packed_lower_floats(r0, r1))
                buf = torch.empty(max(n, 1), dtype=torch.float32, device=self.H.device)
                check(lib.bnnp_pack_lower(ptr(self.H), self.P, r0, r1, ptr(buf), self.comm.cuda_stream), 'bnnp_pack_lower')
                e0.record()
                dist.all_reduce(buf)
                e1.record()
                check(lib.bnnp_unpack_lower(ptr(buf), self.P, r0, r1, ptr(self.H), ptr(self.prior_vec),
                                            self.prior_scalar, self.comm.cuda_stream), 'bnnp_unpack_lower')
                check(lib.bnnp_symmetrize_lower_rows(ptr(self.H), self.P, r0, r1, self.comm.cuda_stream),
                      'bnnp_symmetrize_lower_rows')
                buf.record_stream(self.comm)
            else:
                # every rank has finished accumulating these rows (and the previous chunk's broadcasts have landed)
                self.hdl.barrier(channel=0)
                e0.record()
                check(lib.bnnp_exchange_lower_allreduce(self.peers, self.mc if self.mode == 'nvls' else None, self.rank,
                                                        self.world, self.P, r0, r1, ptr(self.prior_vec),
                                                        self.prior_scalar, EXCHANGE_CTAS, self.comm.cuda_stream),
                      'bnnp_exchange_lower_allreduce')
                e1.record()
                # the barrier above also means the PREVIOUS chunk's broadcasts have landed here: mirror it now, beside
                # the accumulation of the next chunk (only the last chunk's mirror is left for finish())
                if self._unmirrored is not None:
                    check(lib.bnnp_symmetrize_lower_rows(ptr(self.H), self.P, self._unmirrored[0], self._unmirrored[1],
                                                         self.comm.cuda_stream), 'bnnp_symmetrize_lower_rows')
                self._unmirrored = (r0, r1)
            self.events.append((r0, r1, e0, e1))

    def finish(self):
        from ._cabi import lib, check, ptr
        with torch.cuda.stream(self.comm):
            if self.mode != 'nccl':
                self.hdl.barrier(channel=0)          # every owner's broadcasts have landed in this rank's H
                if self._unmirrored is not None:
                    check(lib.bnnp_symmetrize_lower_rows(ptr(self.H), self.P, self._unmirrored[0], self._unmirrored[1],
                                                         self.comm.cuda_stream), 'bnnp_symmetrize_lower_rows')
                    self._unmirrored = None
        torch.cuda.current_stream().wait_stream(self.comm)

    def stats(self):
        """(total exchange ms summed over the chunks, bytes of the triangle exchanged)."""
        ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in self.events)
        nbytes = 4 * sum(r1 * (r1 + 1) // 2 - r0 * (r0 + 1) // 2 for r0, r1, _, _ in self.events)
        return ms, nbytes
