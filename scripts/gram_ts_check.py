"""A/B of the TMEM-A gram kernel (BNNP_GRAM_TS=1) against the CTA-pair kernel: bitwise H and time per launch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200 import Laplace, CategoricalLh, MLPS
archs = [(300, [40, 24], 7, 1100), (784, [75], 10, 4096)]
for din, hid, K, N in archs:
    torch.manual_seed(11)
    model = MLPS(din, hid, K, 'tanh').cuda()
    g = torch.Generator().manual_seed(12)
    X, y = torch.randn(N, din, generator=g).cuda(), torch.randint(K, (N,), generator=g).cuda()
    res = {}
    for mode in ('pair', '208', '160', '128', '96'):
        if mode != 'pair': os.environ['BNNP_GRAM_TS'] = mode
        else: os.environ.pop('BNNP_GRAM_TS', None)
        lap = Laplace(model, 1.0, CategoricalLh()); lap.ggn_engine = 2
        lap.infer([(X, y)], cov_type='full', factorise=False)
        torch.cuda.synchronize()
        lap.infer([(X, y)], cov_type='full', factorise=False)
        res[mode] = (lap.precision.clone(), lap.timings['accumulate_ms'])
    for mode in ('208', '160', '128', '96'):
        same = torch.equal(res['pair'][0], res[mode][0])
        print('arch', (din, hid, K, N), 'ts width', mode, 'bitwise', same, 'ms pair %.2f ts %.2f' % (res['pair'][1], res[mode][1]), flush=True)
