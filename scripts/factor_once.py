"""One factorisation chain at P = 59 635 (for launch lists: ncu --metrics gpu__time_duration.sum ... python scripts/factor_once.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import factor
P = int(sys.argv[1]) if len(sys.argv) > 1 else 59635
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.randn(P, 2048, device='cuda', generator=g)
H = A @ A.T
del A
H.diagonal().add_(float(H.diagonal().mean()) * 0.02)
torch.cuda.synchronize(); t0 = time.perf_counter()
chol, Sigma = factor.posterior_from_precision(H)
torch.cuda.synchronize()
print('ok chain %.3f s, trace(Sigma_chol) %.6f' % (time.perf_counter() - t0, float(chol.diagonal().double().sum())))
