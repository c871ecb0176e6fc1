"""A/B of the Linear-layer products / KFAC Grams: TMA-fed GEMM on pre-split operands (default) against the
generated-operand GEMM (option gemm_generated) on cfg3 (MLP 784-1024-512-256-128-10), kron and diag accumulation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, MLPS
from preds_b200._cabi import lib
torch.manual_seed(71)
model = MLPS(784, [1024, 512, 256, 128], 10, flatten=True).cuda()
g = torch.Generator().manual_seed(15)
N, bs = 4096, 256
X = torch.randn(N, 1, 28, 28, generator=g); y = torch.randint(10, (N,), generator=g)
batches = [(X[i:i + bs].cuda(), y[i:i + bs].cuda()) for i in range(0, N, bs)]
res = {}
for cov in ('kron', 'diag'):
    for name, opt in (('tma', 0), ('generated', 1)):
        lib.bnnp_set_option(b'gemm_generated', opt)
        lap = Laplace(model, 1.0, CategoricalLh())
        ms = []
        for it in range(4):
            lap.infer(batches, cov_type=cov, factorise=False)
            torch.cuda.synchronize()
            ms.append(lap.timings['accumulate_ms'])
        out = lap.precision if cov == 'diag' else torch.cat([m.reshape(-1) for qh in lap.Sigma_chol.qhs for m in qh])
        res[(cov, name)] = out.clone()
        print('%s %-9s min %.2f ms for %d examples -> %.0f examples/s' % (cov, name, min(ms), N, N / min(ms) * 1e3))
    a, b = res[(cov, 'tma')], res[(cov, 'generated')]
    print('%s tma vs generated: max rel diff %.2e' % (cov, float((a - b).abs().max() / b.abs().max())))
lib.bnnp_set_option(b'gemm_generated', 0)
