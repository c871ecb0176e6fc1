"""One GLM predictive (cfg 5 shape: 10 000 inputs x 1000 samples, P = 59 635) for the ncu launch list."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import MLPS, CategoricalLh
from preds_b200.predictives import functional_sampling_predictive
torch.manual_seed(71)
model = MLPS(784, [75], 10, 'tanh', flatten=True).cuda()
P = 59635
NTE = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
Sigma = torch.zeros(P, P, device='cuda'); Sigma.diagonal().fill_(1.0); Sigma[:2000, :2000] += 0.001
Xte = torch.randn(NTE, 1, 28, 28, device='cuda')
mu = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
for it in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    out = functional_sampling_predictive(Xte, model, CategoricalLh(), mu, Sigma, mc_samples=1000, seed=it)
    e1.record(); torch.cuda.synchronize()
    print('predictive ms', e0.elapsed_time(e1))
