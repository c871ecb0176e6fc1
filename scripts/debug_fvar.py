import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import MLPS, Jacobians
from preds_b200.predictives import glm_moments

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())

for (din, hid, K, N) in [(2,[10,10],2,6), (6,[16,12],3,10), (6,[16,12],3,5), (6,[16],3,10), (6,[30],3,10), (6,[50],3,10),
                         (6,[16,12],2,10), (6,[16,12],4,10), (20,[12],3,10), (6,[16,12],3,21), (6,[16,12],3,22)]:
    torch.manual_seed(0)
    model = MLPS(din, hid, K, 'tanh').cuda()
    P = sum(p.numel() for p in model.parameters())
    X = torch.randn(N, din).cuda()
    A = torch.randn(P, P, device='cuda') / P ** 0.5
    Sigma = (A @ A.T + 0.1 * torch.eye(P, device='cuda')).contiguous()
    mu = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    Js, f = Jacobians(model, X)
    Jt = Js.transpose(1, 2) if Js.dim() == 3 else Js.unsqueeze(1)
    ref = torch.einsum('nkp,pq,ncq->nkc', Jt, Sigma, Jt)
    fmu, fvar = glm_moments(X, model, mu, Sigma)
    d = (fvar - ref).abs().amax(dim=(1, 2)) / ref.abs().max()
    print(din, hid, K, 'N', N, 'P', P, 'NK', N * K, 'err %.2e' % rel(fvar, ref), 'per-n', ['%.0e' % float(x) for x in d])
