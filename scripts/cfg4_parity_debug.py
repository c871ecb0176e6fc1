"""Per-layer parity of the conv-net factor paths at cfg4 size (CIFAR10Net(3,10), 32x32, batch 256) against the oracle:
sqrt-GGN gradients at every parameterised layer's output, diagonal GGN and KFLR factors, with the contractions on the
tensor cores (default) and forced onto the fp32 SIMT kernels (option net_simt) -- bisects which kernel a mismatch comes from."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bnn-predictions_b200')):
    sys.path.insert(0, p)
import torch
from oracle import port
from preds_b200 import Laplace, CategoricalLh, CIFAR10Net
from preds_b200 import _cabi as cabi
from preds_b200._cabi import lib
from preds_b200._engine import net_for

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(71)
model = CIFAR10Net(3, 10, use_tanh=True)
omodel = port.make_cifar10net(3, 10, use_tanh=True)
port.set_theta(omodel, port.get_theta(model))
g = torch.Generator().manual_seed(15)
X = torch.randn(N, 3, 32, 32, generator=g)
y = torch.randint(10, (N,), generator=g)
olh = port.CategoricalLh()
model = model.cuda()
names = [n for n, _ in model.named_parameters()]
from tests import routing
INJECT = len(sys.argv) > 2 and sys.argv[2] == 'inject'


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


for simt in (0, 1):
    lib.bnnp_set_option(b'net_simt', simt)
    print('==== net_simt = %d' % simt)
    ref_model = omodel
    if INJECT:
        dec, _ = routing.device_decisions(model, X)
        print('decisions', routing.decision_report(omodel, dec, X))
        ref_model = routing.routed_oracle(omodel, dec)
    rec, Sls = port._sqrt_ggn_backprop(ref_model, olh, X)          # S_l [V, N, out, (H, W)]
    dref = port.diag_ggn_batch(ref_model, olh, X)
    qref = port.kflr_batch(ref_model, olh, X)
    net = net_for(model)
    X2d, pl, ws, f = net.sqrt_factors(X.cuda(), 0)
    with torch.no_grad():
        print('logits', rel(f.cpu(), omodel(X)))
    li = 0
    for i, (op, mod) in enumerate(zip(net.ops, net.mods)):
        if op.kind not in (cabi.OP_LINEAR, cabi.OP_CONV2D):
            continue
        Sl = Sls[li]                                          # [V, N, out, ...]
        V = Sl.shape[0]
        osz = op.out_c * op.out_h * op.out_w
        q = pl.op[i]
        if op.kind == cabi.OP_CONV2D:
            G = torch.as_strided(ws, (N * V, osz), (q.grad_ld, 1), q.grad_off).cpu().reshape(N, V, osz)
        else:
            G = torch.as_strided(ws, (N * V, op.out_c), (pl.Mtot, 1), pl.gcat_off + q.lin_unit).cpu().reshape(N, V, op.out_c)
        ref = Sl.reshape(V, N, -1).transpose(0, 1)
        print('  grad at output of %-10s rel err %.2e' % (names[2 * li].rsplit('.', 1)[0], rel(G, ref)))
        li += 1
    lap = Laplace(model, 0.0, CategoricalLh())
    lap.infer([(X, y)], cov_type='diag', factorise=False)
    d = lap.precision.cpu()
    off = 0
    for name, p in model.named_parameters():
        print('  diag %-18s rel err %.2e' % (name, rel(d[off:off + p.numel()], dref[off:off + p.numel()])))
        off += p.numel()
    lap = Laplace(model, 1.0, CategoricalLh())
    lap.infer([(X, y)], cov_type='kron', factorise=False)
    for pi, (qh, qr) in enumerate(zip(lap.Sigma_chol.qhs, qref)):
        for fi, (m, r) in enumerate(zip(qh, qr)):
            r = r * (1.0 if fi == 0 else 1.0)
            print('  kron %-18s factor %d rel err %.2e' % (names[pi], fi, rel(m.cpu(), r)))
lib.bnnp_set_option(b'net_simt', 0)
