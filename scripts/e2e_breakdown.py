"""Wall-clock breakdown of one Laplace.infer(..., 'full', factorise=False) call on a pinned host batch (cfg 5)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, MLPS
import preds_b200.laplace as L
torch.manual_seed(71)
model = MLPS(784, [75], 10, 'tanh', flatten=True).cuda()
B = 16384
X = torch.randn(B, 1, 28, 28).pin_memory(); y = torch.randint(10, (B,))
lap = Laplace(model, 1.0, CategoricalLh())
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = sync()
    lap.infer([(X, y)], cov_type='full', factorise=False)
    t1 = sync()
    v = float(lap.precision.diagonal().sum())
    t2 = sync()
    print('infer %.1f ms (accumulate event %.1f ms) readback %.1f ms' % ((t1 - t0) * 1e3, lap.timings['accumulate_ms'], (t2 - t1) * 1e3))
# pieces
t0 = sync(); prior = lap.expand_prior_precision('full'); t1 = sync(); print('expand_prior_precision %.1f ms' % ((t1 - t0) * 1e3))
t0 = sync(); Xd = X.to('cuda'); t1 = sync(); print('H2D %.1f ms' % ((t1 - t0) * 1e3))
H = prior
t0 = sync(); lap._accumulate_full(iter([(Xd, y)]), H); t1 = sync(); print('_accumulate_full (device batch) %.1f ms' % ((t1 - t0) * 1e3))
