import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import MLPS
from preds_b200.predictives import glm_moments
torch.manual_seed(71)
model = MLPS(784, [75], 10, 'tanh', flatten=True).cuda()
P = 59635
NTE = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
Sigma = torch.zeros(P, P, device='cuda'); Sigma.diagonal().fill_(1.0); Sigma[:2000, :2000] += 0.001
Xte = torch.randn(NTE, 1, 28, 28, device='cuda')
mu = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
for _ in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fmu, fvar = glm_moments(Xte, model, mu, Sigma, engine=int(sys.argv[2]) if len(sys.argv) > 2 else 0); e1.record(); torch.cuda.synchronize()
    print('moments ms', e0.elapsed_time(e1), 'TF/s executed', 6 * NTE * P * P / e0.elapsed_time(e1) / 1e9)
