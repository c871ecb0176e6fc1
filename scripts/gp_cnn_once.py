"""Function-space inference on CIFAR10Net (conv route: materialised one-column Jacobians, P = 895 210 walked in 219 K segments)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import FunctionaLaplace, CategoricalLh, CIFAR10Net
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nte = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.manual_seed(71)
model = CIFAR10Net(3, 10, use_tanh=True).cuda()
g = torch.Generator().manual_seed(15)
Xtr, Xte = torch.randn(M, 3, 32, 32, generator=g), torch.randn(nte, 3, 32, 32, generator=g)
lap = FunctionaLaplace(model, 50000.0 / M, CategoricalLh())
lap.infer([(Xtr[:16], None)], [(Xte[:16], None)])
torch.cuda.synchronize(); t0 = time.perf_counter()
lap.infer([(Xtr, None)], [(Xte, None)])
torch.cuda.synchronize(); dt = time.perf_counter() - t0
P = lap.P
print('CIFAR10Net P=%d, M=%d x %d test inputs, 10 outputs: infer %.2f s (Gram flops alone at %.0f TFLOP/s bf16)' % (
    P, M, nte, dt, 10.0 * (M * M + M * nte) * P * 6 / dt / 1e12), flush=True)
K = lap.gp_quantities['K_train']
print('K_train symmetric to %.2e, diag min %.3e' % (float((K - K.transpose(1, 2)).abs().max() / K.abs().max()), float(K.diagonal(dim1=1, dim2=2).min())))
t0 = time.perf_counter()
p = lap.predictive_samples(n_samples=1000, batch_size=1000, seed=1)
torch.cuda.synchronize()
print('predictive %.1f ms, rows sum to 1: %s, max prob mean %.3f' % ((time.perf_counter() - t0) * 1e3, bool((p.sum(1) - 1).abs().max() < 1e-5), float(p.max(1).values.mean())))
