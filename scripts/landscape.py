"""Rates of the secondary configurations with the current kernels (device-timed, CUDA events)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, MLPS, CIFAR10Net

def timed(fn, reps=1):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def run(name, model, in_shape, N, bs, covs, nte, S, K=10):
    g = torch.Generator().manual_seed(15)
    X = torch.randn(N, *in_shape, generator=g); y = torch.randint(K, (N,), generator=g)
    batches = [(X[i:i + bs].cuda(), y[i:i + bs].cuda()) for i in range(0, N, bs)]
    Xte = torch.randn(nte, *in_shape, generator=g).cuda()
    for cov in covs:
        lap = Laplace(model, 1.0, CategoricalLh())
        lap.infer(batches[:1], cov_type=cov, factorise=False)
        ms = min(timed(lambda: lap.infer(batches, cov_type=cov, factorise=False)) for _ in range(4))
        lap.infer(batches, cov_type=cov)
        fact = lap.timings['factorise_ms']
        lap.predictive_samples_glm(Xte[:4], n_samples=S)
        gms = min(timed(lambda: lap.predictive_samples_glm(Xte, n_samples=S)) for _ in range(3))
        bms = timed(lambda: lap.predictive_samples_bnn(Xte[:64], n_samples=32))
        print(json.dumps({'cfg': name, 'cov': cov, 'N': N, 'accum_ex_per_s': round(N / ms * 1e3), 'accum_ms': round(ms, 1),
                          'factorise_ms': round(fact, 1), 'glm_units_per_s': round(nte * S / gms * 1e3), 'glm_ms': round(gms, 1),
                          'nte': nte, 'S': S, 'bnn_fwd_per_s(64x32)': round(64 * 32 / bms * 1e3)}), flush=True)

torch.manual_seed(71)
which = sys.argv[1:] or ['cfg1', 'cfg2', 'cfg3', 'cfg4']
if 'cfg1' in which:
    run('cfg1 banana MLPS(2,[50,50],2)', MLPS(2, [50, 50], 2).cuda(), (2,), 5300, 5300, ['full', 'kron', 'diag'], 5300, 1000, K=2)
if 'cfg2' in which:
    run('cfg2 satellite MLPS(36,[50,50],6)', MLPS(36, [50, 50], 6).cuda(), (36,), 4504, 4504, ['full', 'kron', 'diag'], 965, 1000, K=6)
    run('cfg2 waveform MLPS(21,[50,50],3)', MLPS(21, [50, 50], 3).cuda(), (21,), 700, 700, ['full', 'kron', 'diag'], 150, 1000, K=3)
if 'cfg3' in which:
    run('cfg3 MNIST MLP 784-1024-512-256-128-10', MLPS(784, [1024, 512, 256, 128], 10, flatten=True).cuda(), (1, 28, 28), 2048, 256, ['kron', 'diag'], 512, 1000)
if 'cfg4' in which:
    run('cfg4 CIFAR10Net(3,10)', CIFAR10Net(3, 10, use_tanh=True).cuda(), (3, 32, 32), 512, 256, ['kron', 'diag'], 40, 100)
