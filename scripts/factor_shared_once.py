"""Time the factorisation chain shared between the ranks of a torchrun job against the single-rank chain (same H everywhere).
torchrun --nproc-per-node N scripts/factor_shared_once.py [P]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
import torch.distributed as dist
from preds_b200 import factor
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0))))
P = int(sys.argv[1]) if len(sys.argv) > 1 else 59635
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.randn(P, 2048, device='cuda', generator=g)
H = A @ A.T
del A
H.diagonal().add_(float(H.diagonal().mean()) * 0.02)
if world > 1:
    dist.broadcast(H, 0)
    w = torch.zeros(1 << 28, device='cuda'); dist.all_reduce(w); del w          # warm the communicator
def t(fn):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); return out, time.perf_counter() - t0
for shared in ((False, True) if world > 1 else (False,)):
    for rep in range(2):
        L, t1 = t(lambda: factor.cholesky(H, shared=shared))
        S, t2 = t(lambda: factor.inverse_from_chol(L, shared=shared))
        del L
        C_, t3 = t(lambda: factor.cholesky(S, shared=shared))
        chk = float(C_.diagonal().double().sum())
        del S, C_
    if rank == 0:
        print('world %d shared %s: cholesky %.3f s, inverse %.3f s, cholesky(Sigma) %.3f s, total %.3f s (trace %.6f)' % (
            world, shared, t1, t2, t3, t1 + t2 + t3, chk), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
