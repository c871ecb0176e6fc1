"""cfg 5 BNN predictive (1000 inputs x 100 weight samples, P = 59 635, full covariance factor) for the launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import MLPS, CategoricalLh
from preds_b200.predictives import nn_sampling_predictive
torch.manual_seed(71)
model = MLPS(784, [75], 10, 'tanh', flatten=True).cuda()
P = 59635
L = torch.tril(torch.randn(P, P, device='cuda') * 0.001); L.diagonal().fill_(0.05)
mu = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
X = torch.randn(1000, 1, 28, 28, device='cuda')
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for it in range(3):
    if it == 2: torch.cuda.profiler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = nn_sampling_predictive(X, model, CategoricalLh(), mu, L, mc_samples=S); e1.record()
    torch.cuda.synchronize()
    if it == 2: torch.cuda.profiler.stop()
    print('bnn predictive ms', e0.elapsed_time(e1), tuple(out.shape))
