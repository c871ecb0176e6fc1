"""Config 3 (BASELINE.json): MNIST-shape MLP 784-1024-512-256-128-10, kron / diag posterior, GLM predictive on
10 000 inputs x 1000 samples ([1000,10000,10] fp32 = 400 MB), device-timed; also the fused-mean variant."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, MLPS

def timed(fn):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), out

torch.manual_seed(71)
model = MLPS(784, [1024, 512, 256, 128], 10, flatten=True).cuda()
g = torch.Generator().manual_seed(15)
N, bs, nte, S = 4096, 256, 10000, 1000
X = torch.randn(N, 1, 28, 28, generator=g); y = torch.randint(10, (N,), generator=g)
batches = [(X[i:i + bs].cuda(), y[i:i + bs].cuda()) for i in range(0, N, bs)]
Xte = torch.randn(nte, 1, 28, 28, generator=g).cuda()
for cov in ('kron', 'diag'):
    lap = Laplace(model, 1.0, CategoricalLh())
    lap.infer(batches[:2], cov_type=cov)
    ms_acc, _ = timed(lambda: lap.infer(batches, cov_type=cov, factorise=False))
    lap.infer(batches, cov_type=cov)
    lap.predictive_samples_glm(Xte[:64], n_samples=8, seed=1)
    ms, out = timed(lambda: lap.predictive_samples_glm(Xte, n_samples=S, seed=2))
    del out
    lap.predictive_mean_glm(Xte[:64], n_samples=8, seed=1)
    ms_f, pm = timed(lambda: lap.predictive_mean_glm(Xte, n_samples=S, seed=2))
    print(json.dumps({'cfg': 'cfg3 MLP P=1494154', 'cov': cov, 'accum_ex_per_s': round(N / ms_acc * 1e3),
                      'glm_ms': round(ms, 2), 'glm_units_per_s': round(nte * S / ms * 1e3),
                      'glm_fused_mean_ms': round(ms_f, 2), 'glm_fused_units_per_s': round(nte * S / ms_f * 1e3),
                      'rows_sum_to_one': float((pm.sum(1) - 1).abs().max())}), flush=True)
