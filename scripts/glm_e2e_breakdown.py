"""Wall-clock pieces of the end-to-end GLM predictive (cfg 5 shape) around the device-timed kernel chain."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import MLPS, CategoricalLh
from preds_b200.predictives import functional_sampling_predictive, functional_mean_predictive, glm_moments
torch.manual_seed(71)
model = MLPS(784, [75], 10, 'tanh', flatten=True).cuda()
P = 59635
Sigma = torch.zeros(P, P, device='cuda'); Sigma.diagonal().fill_(1.0); Sigma[:2000, :2000] += 0.001
Xh = torch.randn(10000, 1, 28, 28).pin_memory()
mu = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
lh = CategoricalLh()
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = sync(); Xd = Xh.to('cuda'); t1 = sync()
    fmu, fvar = glm_moments(Xd, model, mu, Sigma); t2 = sync()
    out = functional_sampling_predictive(Xd, model, lh, mu, Sigma, mc_samples=1000, seed=it); t3 = sync()
    pm = out.mean(0); t4 = sync()
    pmc = pm.cpu(); t5 = sync()
    del out
    t6 = sync(); out2 = functional_sampling_predictive(Xh, model, lh, mu, Sigma, mc_samples=1000, seed=it).mean(0).cpu(); t7 = sync()
    t8 = sync(); pm2 = functional_mean_predictive(Xh, model, lh, mu, Sigma, mc_samples=1000, seed=it).cpu(); t9 = sync()
    print('fused-mean e2e call %.1f ms, max diff to materialised mean %.2g' % (1e3 * (t9 - t8), float((pm2 - out2).abs().max())))
    print('H2D %.1f | moments %.1f | moments+sampling %.1f | mean(0) %.1f | D2H %.1f | whole e2e call %.1f ms' %
          tuple(1e3 * d for d in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t7 - t6)), flush=True)
