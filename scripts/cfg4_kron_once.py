import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, CIFAR10Net
torch.manual_seed(71)
model = CIFAR10Net(3, 10, use_tanh=True).cuda()
cov = sys.argv[1] if len(sys.argv) > 1 else 'kron'
g = torch.Generator().manual_seed(15)
X = torch.randn(512, 3, 32, 32, generator=g).cuda(); y = torch.randint(10, (512,), generator=g).cuda()
lap = Laplace(model, 1.0, CategoricalLh())
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ms = []
for it in range(n):
    if it == n - 1: torch.cuda.profiler.start()       # ncu --profile-from-start off: the steady-state pass only
    lap.infer([(X, y)], cov_type=cov, factorise=False)
    torch.cuda.synchronize()
    if it == n - 1: torch.cuda.profiler.stop()
    ms.append(lap.timings['accumulate_ms'])
torch.cuda.synchronize()
print('ok min %.2f ms over %d infers of 512 examples -> %.0f examples/s' % (min(ms), n, 512 / min(ms) * 1e3), lap.timings)
