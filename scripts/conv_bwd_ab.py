"""A/B of the conv backward-data: implicit GEMM on 4-D TMA boxes (default) against product + col2im gather (option
conv_bwd_explicit) on cfg4 (CIFAR10Net(3,10), 32x32, 512 examples), kron and diag accumulation, with the difference of the
results."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, CIFAR10Net
from preds_b200._cabi import lib
torch.manual_seed(71)
model = CIFAR10Net(3, 10, use_tanh=True).cuda()
g = torch.Generator().manual_seed(15)
N = 512
X = torch.randn(N, 3, 32, 32, generator=g).cuda(); y = torch.randint(10, (N,), generator=g).cuda()
res = {}
for cov in ('kron', 'diag'):
    for name, opt in (('implicit', 0), ('explicit', 1)):
        assert lib.bnnp_set_option(b'conv_bwd_explicit', opt) == 0
        lap = Laplace(model, 1.0, CategoricalLh())
        ms = []
        for it in range(4):
            lap.infer([(X, y)], cov_type=cov, factorise=False)
            torch.cuda.synchronize()
            ms.append(lap.timings['accumulate_ms'])
        out = lap.precision if cov == 'diag' else torch.cat([m.reshape(-1) for qh in lap.Sigma_chol.qhs for m in qh])
        res[(cov, name)] = out.clone()
        print('%s %-9s min %.2f ms for %d examples -> %.0f examples/s' % (cov, name, min(ms), N, N / min(ms) * 1e3), flush=True)
    a, b = res[(cov, 'implicit')], res[(cov, 'explicit')]
    print('%s implicit vs explicit: max rel diff %.2e' % (cov, float((a - b).abs().max() / b.abs().max())), flush=True)
lib.bnnp_set_option(b'conv_bwd_explicit', 0)
