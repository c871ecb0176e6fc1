import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, MLPS
torch.manual_seed(71)
model = MLPS(784, [1024, 512, 256, 128], 10, flatten=True).cuda()
g = torch.Generator().manual_seed(15)
X = torch.randn(512, 1, 28, 28, generator=g).cuda(); y = torch.randint(10, (512,), generator=g).cuda()
Xte = torch.randn(10000, 1, 28, 28, generator=g).cuda()
lap = Laplace(model, 1.0, CategoricalLh())
lap.infer([(X, y)], cov_type=sys.argv[1] if len(sys.argv) > 1 else 'kron')
out = lap.predictive_samples_glm(Xte, n_samples=1000, seed=2)
torch.cuda.synchronize()
torch.cuda.profiler.start()            # ncu --profile-from-start off: only the steady-state predictive is listed
out = lap.predictive_samples_glm(Xte, n_samples=1000, seed=2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('ok')
