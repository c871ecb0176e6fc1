"""Accuracy of the full-GGN engines vs an fp64 reference as a function of the launch size N
(number of sequential tensor-core accumulator updates grows with N)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
import preds_b200.laplace as L
from preds_b200 import Laplace, CategoricalLh, MLPS, Jacobians

torch.manual_seed(71)
model = MLPS(64, [32], 10, 'tanh').cuda()
P = sum(p.numel() for p in model.parameters())
lh = CategoricalLh()
for N in (1024, 4096, 16384, 65536):
    g = torch.Generator().manual_seed(15)
    X = torch.randn(N, 64, generator=g).cuda()
    y = torch.randint(10, (N,), generator=g).cuda()
    # fp64 reference from materialised Jacobians, chunked
    H64 = torch.zeros(P, P, dtype=torch.float64, device='cuda')
    for s in range(0, N, 1024):
        Js, f = Jacobians(model, X[s:s + 1024])
        Lam = lh.Hessian(f).double()
        Jd = Js.double()
        H64 += torch.einsum('mpk,mkl,mql->pq', Jd, Lam, Jd)
    out = {}
    for eng in (1, 2):
        lap = Laplace(model, 0.0, lh); lap.ggn_engine = eng
        lap.infer([(X, y)], cov_type='full', factorise=False)
        H = lap.precision.double()
        d = (H - H64)
        out[eng] = (float(d.abs().max() / H64.abs().max()), float(d.norm() / H64.norm()),
                    float(((H.diagonal() - H64.diagonal()) / H64.diagonal()).mean()))
    print('N=%6d  simt: max %.2e fro %.2e diagbias %+.2e | tcgen05: max %.2e fro %.2e diagbias %+.2e'
          % (N, *out[1], *out[2]), flush=True)
