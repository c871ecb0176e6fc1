"""Multi-GPU probe of the full-covariance exchange (run under torchrun): which modes exist on this box (symmetric
memory, NVLS multicast), correctness of each against NCCL, and time / bus bandwidth at P = 59 635 for
  (a) NCCL all-reduce of the unpacked square (round-1 behaviour), (b) packed NCCL, (c) own P2P kernel, (d) own NVLS kernel.
busbw = 2 (W-1)/W x bytes / time (the all-reduce convention), bytes = the lower triangle (or the square for (a))."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bnn-predictions_b200')):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from preds_b200 import parallel
    from preds_b200._cabi import lib, check, ptr
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 59635
    out = {'world': world, 'P': P}
    H, hdl = parallel.symmetric_zeros(P, dev)
    out['symmetric_memory'] = hdl is not None
    out['multicast'] = bool(hdl is not None and int(hdl.multicast_ptr))
    if H is None:
        H = torch.zeros(P, P, device=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)

    def fill():
        H.copy_(torch.tril(torch.randn(P, P, device=dev, generator=g)) if P <= 16384 else H.normal_(generator=g))
        torch.cuda.synchronize()
        dist.barrier()

    def timed(fn, reps=3):
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t) if best is None else min(best, float(t))
        return best

    tri_bytes = 4 * P * (P + 1) // 2
    bus = lambda nbytes, ms: 2 * (world - 1) / world * nbytes / (ms * 1e-3) / 1e9

    # reference result for correctness (small P only: keeps a copy)
    modes = ['nccl'] + (['p2p'] if hdl is not None else []) + (['nvls'] if out['multicast'] else [])
    if P <= 16384:
        fill()
        ref = H.clone()
        dist.all_reduce(ref)
        ref = torch.tril(ref)
        ref.diagonal().add_(0.25)
        src = H.clone()
        for mode in modes:
            H.copy_(src)
            torch.cuda.synchronize()
            dist.barrier()
            ex = parallel.LowerExchange(H, hdl, 0.25, mode=mode)
            third = P // 3 // 4 * 4 + 1
            for r0, r1 in ((0, third), (third, 2 * third), (2 * third, P)):
                ex.rows_done(r0, r1)
            ex.finish()
            torch.cuda.synchronize()
            err = float((torch.tril(H) - ref).abs().max() / ref.abs().max())
            out['err_' + mode] = err
    # timing
    fill()
    t = timed(lambda: dist.all_reduce(H))
    out['nccl_square'] = {'ms': t, 'busbw_GBs': bus(4 * P * P, t)}
    for mode in modes:
        for ctas in ((16, 32, 64, 128) if mode != 'nccl' else (0,)):
            parallel.EXCHANGE_CTAS = ctas

            def run():
                ex = parallel.LowerExchange(H, hdl, 0.0, mode=mode)
                ex.rows_done(0, P)
                ex.finish()
                run.ex = ex
            t = timed(run)
            kms, nb = run.ex.stats()
            out['%s_ctas%d' % (mode, ctas)] = {'ms_total': t, 'kernel_ms': kms, 'busbw_GBs_kernel': bus(tri_bytes, kms),
                                               'busbw_GBs_total': bus(tri_bytes, t)}
    if rank == 0:
        print(json.dumps(out))
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'exchange_probe_w%d_P%d.json' % (world, P)), 'w') as fh:
            json.dump(out, fh, indent=1)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
