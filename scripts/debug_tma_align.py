import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import MLPS, Jacobians
from preds_b200.predictives import glm_moments
for (din, hid, K, N) in [(16, [24], 4, 70), (64, [40], 8, 300), (16, [24], 5, 70)]:
    torch.manual_seed(3)
    model = MLPS(din, hid, K, 'tanh').cuda()
    P = sum(p.numel() for p in model.parameters())
    X = torch.randn(N, din, device='cuda')
    A = torch.randn(P, P, device='cuda') / P ** 0.5
    Sigma = (A @ A.T + 0.05 * torch.eye(P, device='cuda')).contiguous()
    mu = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    Js, f = Jacobians(model, X)
    Jt = Js.transpose(1, 2).double()
    ref = torch.einsum('nkp,pq,ncq->nkc', Jt, Sigma.double(), Jt)
    try:
        fmu, fvar = glm_moments(X, model, mu, Sigma, engine=3)
        torch.cuda.synchronize()
        print(din, hid, K, 'P', P, 'err', float((fvar.double() - ref).abs().max() / ref.abs().max()), flush=True)
    except Exception as e:
        print(din, hid, K, 'P', P, 'FAILED', repr(e)[:100], flush=True)
        break
