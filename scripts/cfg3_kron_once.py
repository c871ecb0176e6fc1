import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, MLPS
torch.manual_seed(71)
model = MLPS(784, [1024, 512, 256, 128], 10, flatten=True).cuda()
g = torch.Generator().manual_seed(15)
N, bs = 4096, 256
X = torch.randn(N, 1, 28, 28, generator=g); y = torch.randint(10, (N,), generator=g)
batches = [(X[i:i + bs].cuda(), y[i:i + bs].cuda()) for i in range(0, N, bs)]
cov = sys.argv[1] if len(sys.argv) > 1 else 'kron'
lap = Laplace(model, 1.0, CategoricalLh())
ms = []
for it in range(4):
    if it == 3: torch.cuda.profiler.start()
    lap.infer(batches, cov_type=cov, factorise=False)
    torch.cuda.synchronize()
    if it == 3: torch.cuda.profiler.stop()
    ms.append(lap.timings['accumulate_ms'])
print('ok min %.2f ms for %d examples -> %.0f examples/s' % (min(ms), N, N / min(ms) * 1e3))
