"""Accuracy and time of the tensor-core factorisation chain (preds_b200.factor) against cuSOLVER through torch:
chol(H), Sigma = H^-1, chol(Sigma) on a synthetic SPD matrix with a prior-dominated small end of the spectrum."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import factor
P = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
ref = len(sys.argv) > 2 and sys.argv[2] == 'ref'
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.randn(P, 2048, device='cuda', generator=g)
H = A @ A.T
del A
H.diagonal().add_(float(H.diagonal().mean()) * 0.02)
def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); return out, time.perf_counter() - t0
idx = torch.randperm(P, device='cuda', generator=g)[:256]
E = torch.zeros(P, 256, device='cuda', dtype=torch.float64); E[idx, torch.arange(256, device='cuda')] = 1
def resid(S):
    return float((H.double() @ S[:, idx].double() - E).abs().max())
def chol_resid(L, M):
    # |L L^T - M| on a probe of columns, relative to max |M|
    return float(((L.double() @ L[idx].double().T) - M[:, idx].double()).abs().max() / M.abs().max())
for rep in range(2):
    L1, t1 = t(lambda: factor.cholesky_lower(H.clone()))
print('P %d  cholesky_lower %.3f s  |L L^T - H| / max|H| = %.3g' % (P, t1, chol_resid(L1, H)), flush=True)
S1, t2 = t(lambda: factor.inverse_from_chol(L1))
print('inverse_from_chol %.3f s  |H S - I|_max = %.3g  symmetric %s' % (t2, resid(S1), bool(torch.equal(S1, S1.T))), flush=True)
C1, t3 = t(lambda: factor.cholesky_lower(S1.clone()))
print('cholesky_lower(Sigma) %.3f s  |C C^T - S| / max|S| = %.3g   chain total %.3f s' % (t3, chol_resid(C1, S1), t1 + t2 + t3), flush=True)
if ref:
    L0, u1 = t(lambda: torch.linalg.cholesky(H))
    print('torch cholesky %.3f s  |L L^T - H| / max|H| = %.3g   max|L1 - L0| / max|L0| = %.3g' % (
        u1, chol_resid(L0, H), float((L1 - L0).abs().max() / L0.abs().max())), flush=True)
    del L1
    S0, u2 = t(lambda: torch.cholesky_inverse(L0, upper=False))
    S0 = S0.T if not S0.is_contiguous() else S0
    print('torch cholesky_inverse %.3f s  |H S - I|_max = %.3g   max|S1 - S0| / max|S0| = %.3g' % (
        u2, resid(S0), float((S1 - S0).abs().max() / S0.abs().max())), flush=True)
    del L0, S1
    C0, u3 = t(lambda: torch.linalg.cholesky(S0))
    print('torch cholesky(Sigma) %.3f s  |C C^T - S| / max|S| = %.3g   max|C1 - C0| / max|C0| = %.3g   chain total %.3f s' % (
        u3, chol_resid(C0, S0), float((C1 - C0).abs().max() / C0.abs().max()), u1 + u2 + u3), flush=True)
