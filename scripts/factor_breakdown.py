"""Where the time of factor.inverse_from_chol / cholesky_lower goes (synchronising timers around the phases)."""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import factor
P = int(sys.argv[1]) if len(sys.argv) > 1 else 59635
if len(sys.argv) > 2: factor.KSEG = int(sys.argv[2])
if len(sys.argv) > 3: factor.NB = int(sys.argv[3])
print('P', P, 'KSEG', factor.KSEG, 'NB', factor.NB)
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.randn(P, 2048, device='cuda', generator=g)
H = A @ A.T
del A
H.diagonal().add_(float(H.diagonal().mean()) * 0.02)
acc = collections.OrderedDict()
def wrap(obj, name, label):
    fn = getattr(obj, name)
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0
        return out
    setattr(obj, name, w)
orig_gemm3 = factor._gemm3
def gemm3(M, N, K, alpha, A_, a, B_, b, beta, c, ldc, lower=0, ktri=0, koff=0):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    orig_gemm3(M, N, K, alpha, A_, a, B_, b, beta, c, ldc, lower, ktri, koff)
    torch.cuda.synchronize()
    label = 'gemm3 lower=%d ktri=%d N%s' % (lower, ktri, '=nb' if N <= factor.NB else '>nb')
    acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0
factor._gemm3 = gemm3
wrap(factor, '_tri_inv', 'tri_inv (library)')
wrap(factor._Planes, 'fill', 'split3 fills')
wrap(torch.linalg, 'cholesky_ex', 'cholesky_ex (library)')
def run(name, fn):
    acc.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print('%s total %.3f s (with synchronising timers)' % (name, tot))
    for k, v in acc.items(): print('    %-32s %.3f s' % (k, v))
    print('    %-32s %.3f s' % ('other (allocs, zeroing, tril, python)', tot - sum(acc.values())), flush=True)
    return out
L = run('cholesky_lower', lambda: factor.cholesky_lower(H))
S = run('inverse_from_chol', lambda: factor.inverse_from_chol(L))
