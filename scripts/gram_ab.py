"""A/B of the full-GGN tensor-core kernels at cfg5 (P = 59 635): time per 4096-example launch of the dominant
(layer 1, layer 1) block and bitwise equality of H.  kernels: wide (default), single (option gram_kernel=1), pair
(gram_kernel=2, only in BNNP_EXPERIMENTS builds)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bnn-predictions_b200')):
    sys.path.insert(0, p)
import torch
from preds_b200 import MLPS, CategoricalLh
from preds_b200._cabi import lib, check, ptr, stream
from preds_b200._engine import net_for

torch.manual_seed(71)
model = MLPS(784, [75], 10, 'tanh', flatten=True).cuda()
net = net_for(model)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
X = torch.randn(N, 1, 28, 28, generator=torch.Generator().manual_seed(15)).cuda()
P = net.P
H = torch.zeros(P, P, device='cuda')
lh = CategoricalLh()
out = {}
ref = None
for name, opt in (('wide', 0), ('pair', 2), ('single', 1), ('wide2', 0)):
    lib.bnnp_set_option(b'gram_kernel', opt)
    X2d, pl, ws, f = net.jacobian_factors(X)
    Lam = lh._hessian_nkk(f)
    sc = net.scratch(lib.bnnp_ggn_full_scratch_floats(C.byref(pl), N, 2))
    for rep in range(2):
        H.zero_()
        lib.bnnp_prof_enable(1)
        check(lib.bnnp_ggn_full_accum(net.ops, net.n_ops, C.byref(pl), N, ptr(ws), ptr(Lam), ptr(H), ptr(sc), 2, stream()))
        torch.cuda.synchronize()
        cap = 64
        ms, ex, us = (C.c_float * cap)(), (C.c_double * cap)(), (C.c_double * cap)()
        n = lib.bnnp_prof_read(ms, ex, us, cap)
        lib.bnnp_prof_enable(0)
    top = max(us[i] for i in range(n))
    dom = [i for i in range(n) if us[i] >= 0.5 * top]
    t = sum(ms[i] for i in dom) / len(dom)
    out[name] = {'ms_per_launch': t, 'useful_tflops': top / t / 1e9, 'executed_tflops': ex[dom[0]] / t / 1e9,
                 'checksum': float(H.diagonal().double().sum())}
    if ref is None:
        ref = H.clone()
    else:
        out[name]['bitwise_equal_to_wide'] = bool(torch.equal(torch.tril(H), torch.tril(ref)))
lib.bnnp_set_option(b'gram_kernel', 0)
print(json.dumps(out, indent=1))
