import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import MLPS, Jacobians
from preds_b200.predictives import glm_moments
from preds_b200._engine import net_for

din, hid, K, N = 6, [30], 3, 4
torch.manual_seed(0)
model = MLPS(din, hid, K, 'tanh').cuda()
P = sum(p.numel() for p in model.parameters())
X = torch.randn(N, din).cuda()
A = torch.randn(P, P, device='cuda') / P ** 0.5
Sigma = (A @ A.T + 0.1 * torch.eye(P, device='cuda')).contiguous()
mu = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
Js, f = Jacobians(model, X)
Jt = Js.transpose(1, 2).contiguous()
fmu, fvar = glm_moments(X, model, mu, Sigma)
torch.cuda.synchronize()
net = net_for(model)
Ppad = (P + 31) // 32 * 32
sc = net._scratch
pu = sc[:Ppad].view(torch.int32)[:P].cpu()
pc = sc[Ppad:2 * Ppad].view(torch.int32)[:P].cpu()
T = sc[2 * Ppad: 2 * Ppad + N * K * P].reshape(N * K, P)
Tref = Jt.reshape(N * K, P) @ Sigma
d = (T - Tref).abs()
print('P', P, 'T err', float(d.max() / Tref.abs().max()))
bad_cols = (d.amax(0) > 1e-4 * Tref.abs().max()).nonzero().flatten().tolist()
bad_rows = (d.amax(1) > 1e-4 * Tref.abs().max()).nonzero().flatten().tolist()
print('bad cols', bad_cols[:20], '... n=', len(bad_cols)); print('bad rows', bad_rows)
print('pu', pu.tolist()[:8], pu.tolist()[175:190], pu.tolist()[-8:])
print('pc', pc.tolist()[:8], pc.tolist()[175:190], pc.tolist()[-8:])
ref = torch.einsum('nkp,pq,ncq->nkc', Jt, Sigma, Jt)
f2 = torch.einsum('rq,rcq->rc', Tref, Jt.repeat_interleave(K, 0).reshape(N * K, K, P))
print('fvar err', float((fvar - ref).abs().max() / ref.abs().max()))
# contribution check: fvar from device T with torch
f3 = torch.einsum('nkq,ncq->nkc', T.reshape(N, K, P), Jt)
print('fvar(torch from device T) err', float((f3 - ref).abs().max() / ref.abs().max()))
