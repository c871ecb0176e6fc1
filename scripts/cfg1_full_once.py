import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, MLPS
torch.manual_seed(71)
model = MLPS(2, [50, 50], 2).cuda()
g = torch.Generator().manual_seed(15)
X = torch.randn(5300, 2, generator=g).cuda(); y = torch.randint(2, (5300,), generator=g).cuda()
for eng in (0, 1, 2):
    lap = Laplace(model, 1.0, CategoricalLh()); lap.ggn_engine = eng
    ms = []
    for it in range(6):
        lap.infer([(X, y)], cov_type='full', factorise=False)
        ms.append(lap.timings['accumulate_ms'])
    print('engine', eng, 'min %.2f ms -> %.0f examples/s' % (min(ms), 5300 / min(ms) * 1e3))
