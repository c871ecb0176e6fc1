"""cfg5 end to end at full size: infer (accumulation | one-off factorisation), GLM 10k x 1000, BNN."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, MLPS
from preds_b200.predictives import glm_moments

def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e

torch.manual_seed(71)
model = MLPS(784, [75], 10, 'tanh', flatten=True).cuda()
NTR = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
NTE = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
g = torch.Generator().manual_seed(15)
X = torch.randn(NTR, 1, 28, 28, generator=g).cuda(); y = torch.randint(10, (NTR,), generator=g).cuda()
Xte = torch.randn(NTE, 1, 28, 28, generator=torch.Generator().manual_seed(16)).cuda()
lap = Laplace(model, 1.0, CategoricalLh())
t0 = time.time()
lap.infer([(X, y)], cov_type='full')
torch.cuda.synchronize()
print(json.dumps({'infer_wall_s': round(time.time() - t0, 2), **{k: round(v, 1) for k, v in lap.timings.items()},
                  'mem_GB': round(torch.cuda.max_memory_allocated() / 1e9, 1)}), flush=True)
# warm-up + timed GLM predictive
lap.predictive_samples_glm(Xte[:256], n_samples=8)
torch.cuda.synchronize()
a = ev(); fmu, fvar = glm_moments(Xte, model, lap.mu, lap.Sigma); b = ev()
out = lap.predictive_samples_glm(Xte, n_samples=1000, seed=1); c = ev()
torch.cuda.synchronize()
mom_ms, tot_ms = a.elapsed_time(b), b.elapsed_time(c)
print(json.dumps({'glm_inputs': NTE, 'samples': 1000, 'moments_ms': round(mom_ms, 1), 'predictive_total_ms': round(tot_ms, 1),
                  'units_per_s': round(NTE * 1000 / tot_ms * 1e3), 'fvar_min_diag': float(fvar.diagonal(dim1=1, dim2=2).min()),
                  'mean_prob_sum': float(out.mean(0).sum(1).mean())}), flush=True)
a = ev(); o2 = lap.predictive_samples_bnn(Xte[:1000], n_samples=100); b = ev(); torch.cuda.synchronize()
print(json.dumps({'bnn_inputs': 1000, 'samples': 100, 'ms': round(a.elapsed_time(b), 1)}), flush=True)
