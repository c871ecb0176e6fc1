"""Accuracy and time of preds_b200.dense.spd_inverse_from_chol against torch.cholesky_inverse."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import dense
P = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.randn(P, 2048, device='cuda', generator=g)
H = A @ A.T
del A
H.diagonal().add_(float(H.diagonal().mean()) * 0.02)
L = torch.linalg.cholesky(H)
def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); return out, time.perf_counter() - t0
dense.ENGINE = sys.argv[2] if len(sys.argv) > 2 else 'cublas'
S1, t1 = t(lambda: dense.spd_inverse_from_chol(L))
# residual ||H S - I||_max on a random probe of columns (the full product is P^3)
idx = torch.randperm(P, device='cuda', generator=g)[:256]
E = torch.zeros(P, 256, device='cuda'); E[idx, torch.arange(256, device='cuda')] = 1
r1 = float((H.double() @ S1[:, idx].double() - E.double()).abs().max())
print('P', P, 'engine', dense.ENGINE, 'blocked inverse %.2f s, residual |H S - I|_max %.3g, symmetric %s' % (t1, r1, bool(torch.equal(S1, S1.T))), flush=True)
if P <= 40000 or (len(sys.argv) > 3 and sys.argv[3] == 'ref'):
    S0, t0 = t(lambda: torch.cholesky_inverse(L, upper=False))
    r0 = float((H.double() @ S0[:, idx].double() - E.double()).abs().max())
    rel = float((S1 - S0).abs().max() / S0.abs().max())
    print('torch.cholesky_inverse %.2f s, residual %.3g; max |S1 - S0| / max |S0| = %.3g' % (t0, r0, rel))
