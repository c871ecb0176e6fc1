"""Function-space inference on the config-5 network (M training x nte test inputs): both kernel routes, timed."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import FunctionaLaplace, CategoricalLh, MLPS, gps
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1600
nte = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
torch.manual_seed(71)
model = MLPS(784, [75], 10, 'tanh', flatten=True).cuda()
g = torch.Generator().manual_seed(15)
Xtr, Xte = torch.randn(M, 1, 28, 28, generator=g), torch.randn(nte, 1, 28, 28, generator=g)
lap = FunctionaLaplace(model, 60000.0 / M, CategoricalLh())
res = {}
for route in (True, False):
    gps.STRUCTURED = route
    lap.infer([(Xtr[:64], None)], [(Xte[:64], None)])
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lap.infer([(Xtr, None)], [(Xte, None)])
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    res[route] = {k: v.clone() for k, v in lap.gp_quantities.items() if torch.is_tensor(v)}
    print('route %-12s infer %.1f ms' % ('structured' if route else 'materialised', best * 1e3), flush=True)
for k in res[True]:
    a, b = res[True][k], res[False][k]
    print('  %-14s max rel diff between the routes %.2e' % (k, float((a - b).abs().max() / b.abs().max())))
lap.predictive_samples(n_samples=8, batch_size=8, seed=1)
for indep in (False, True):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = lap.predictive_samples(n_samples=1000, batch_size=1000, indep=indep, seed=2)
    torch.cuda.synchronize()
    print('predictive indep=%s %.1f ms, rows sum to 1: %s' % (indep, (time.perf_counter() - t0) * 1e3, bool((p.sum(1) - 1).abs().max() < 1e-5)))
