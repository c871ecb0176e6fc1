"""A/B of the cfg3 GLM predictive (10 000 inputs x 1000 samples, MLP 784-1024-512-256-128-10), kron and diag posteriors:
TMA-fed GEMM on pre-split operands (default) against the generated-operand GEMM (option gemm_generated)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, MLPS
from preds_b200._cabi import lib
torch.manual_seed(71)
model = MLPS(784, [1024, 512, 256, 128], 10, flatten=True).cuda()
g = torch.Generator().manual_seed(15)
X = torch.randn(512, 1, 28, 28, generator=g).cuda(); y = torch.randint(10, (512,), generator=g).cuda()
Xte = torch.randn(10000, 1, 28, 28, generator=g).cuda()
for cov in ('kron', 'diag'):
    lap = Laplace(model, 1.0, CategoricalLh())
    lap.infer([(X, y)], cov_type=cov)
    outs = {}
    for name, opt in (('tma', 0), ('generated', 1)):
        lib.bnnp_set_option(b'gemm_generated', opt)
        best = 1e9
        for it in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = lap.predictive_samples_glm(Xte, n_samples=1000, seed=2)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        outs[name] = out.mean(0)
        print('%s %-9s %.2f ms -> %.0fM (inputs x samples)/s' % (cov, name, best, 1e7 / best / 1e3))
    print('%s max abs diff of the predictive means %.2e' % (cov, float((outs['tma'] - outs['generated']).abs().max())))
lib.bnnp_set_option(b'gemm_generated', 0)
