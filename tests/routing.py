"""Test infrastructure: the oracle evaluated with the device path's DISCRETE DECISIONS injected.

A conv net takes non-smooth decisions -- which inputs a ReLU lets through, which window element a max-pool routes its
gradient to.  Two fp32 evaluations of the same network (this library, MKL on the host, the reference's own cuDNN run)
round their convolutions differently at the 1e-6..1e-5 level, so among the millions of decisions of a 256-example batch
a handful flips: a pre-activation of +-1e-7, two window elements 1e-7 apart.  Each flip moves one gradient value and
shows up as a 1e-4..1e-3 outlier in the per-layer second-order quantities -- in ANY pair of implementations, including
reference-vs-reference.  For parity at BASELINE sizes the decisions are therefore treated like the standard-normal
draws: taken from one side (the device path) and injected into the other (the oracle), after which everything left is
smooth arithmetic and must agree to the stated tolerance in max-norm.  ``decision_report`` checks the other half: the
device's decisions differ from the oracle's own only where the oracle's margin is within rounding of zero (on identical
inputs the arg-max is bit-exact: tests/test_gpu_edge_cases.py, test_gpu_reference_suite.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import port


class MaskedReLU(nn.Module):
    def __init__(self, mask):
        super().__init__()
        self.mask = mask

    def forward(self, x):
        return x * self.mask.reshape(x.shape).to(x.dtype)


class RoutedPool(nn.Module):
    """Max-pool with given winners: idx[n,c,oh,ow] = ih*W+iw in the UNPADDED input plane, -1 = a padding zero."""

    def __init__(self, idx):
        super().__init__()
        self.idx = idx.long()

    def forward(self, x):
        N, C, H, W = x.shape
        xz = torch.cat([x.reshape(N, C, H * W), x.new_zeros(N, C, 1)], dim=2)
        ii = torch.where(self.idx < 0, torch.full_like(self.idx, H * W), self.idx)
        return torch.gather(xz, 2, ii.reshape(N, C, -1)).reshape(self.idx.shape)


def device_decisions(model_gpu, X):
    """[(kind, tensor)] per ReLU / max-pool op of the device path's forward on X, in execution order."""
    from preds_b200 import _cabi as cabi
    from preds_b200._engine import net_for
    net = net_for(model_gpu)
    X2d = net.prepare(X.cuda())
    N = X2d.shape[0]
    pl, ws, f = net.forward(X2d, 1)
    out = []
    for i, op in enumerate(net.ops):
        q = pl.op[i]
        osz = op.out_c * op.out_h * op.out_w
        if op.kind == cabi.OP_RELU:
            act = torch.as_strided(ws, (N, osz), (q.act_ld, 1), q.act_off)
            out.append(('relu', (act > 0).reshape(N, op.out_c, op.out_h, op.out_w).cpu()))
        elif op.kind == cabi.OP_MAXPOOL2D:
            idx = ws.view(torch.int32)[q.idx_off: q.idx_off + N * osz]
            out.append(('pool', idx.reshape(N, op.out_c, op.out_h, op.out_w).cpu().clone()))
    return out, f.cpu()


def routed_oracle(omodel, decisions):
    """Copy of the (nn.Sequential) oracle model that shares its parameters and takes the given decisions."""
    mods = list(omodel.children())
    it = iter(decisions)
    new = []
    for k, m in enumerate(mods):
        if isinstance(m, nn.ReLU):
            kind, t = next(it)
            assert kind == 'relu'
            new.append(MaskedReLU(t))
        elif isinstance(m, port.SamePad2d) and k + 1 < len(mods) and isinstance(mods[k + 1], nn.MaxPool2d):
            continue                                   # folded into the routed pool (idx refers to the unpadded plane)
        elif isinstance(m, nn.MaxPool2d):
            kind, t = next(it)
            assert kind == 'pool'
            new.append(RoutedPool(t))
        else:
            new.append(m)
    net = nn.Sequential(*new)
    net.output_size = omodel.output_size
    return net


def decision_report(omodel, decisions, X):
    """Compares the injected decisions with the ones the oracle takes by itself on X.  Returns
    {'relu_flips', 'relu_total', 'relu_worst_margin', 'pool_flips', 'pool_total', 'pool_worst_margin', 'scale'}: margins are
    how far the ORACLE's own activations are from making the device's decision (|pre-activation| for a ReLU, winner minus
    the device-chosen element for a pool), i.e. what rounding would have to explain."""
    mods = list(omodel.children())
    it = iter(decisions)
    rep = dict(relu_flips=0, relu_total=0, relu_worst_margin=0.0, pool_flips=0, pool_total=0, pool_worst_margin=0.0,
               scale=0.0)
    x = X
    with torch.no_grad():
        for k, m in enumerate(mods):
            if isinstance(m, nn.ReLU):
                kind, t = next(it)
                t = t.reshape(x.shape)
                own = x > 0
                flip = own != t
                rep['relu_flips'] += int(flip.sum())
                rep['relu_total'] += flip.numel()
                if flip.any():
                    rep['relu_worst_margin'] = max(rep['relu_worst_margin'], float(x[flip].abs().max()))
                rep['scale'] = max(rep['scale'], float(x.abs().max()))
                x = m(x)
            elif isinstance(m, nn.MaxPool2d):
                kind, idx = next(it)
                own = m(x)                              # x is already padded by the preceding SamePad2d
                chosen = RoutedPool(idx)(unpadded)
                gap = (own - chosen).abs()
                flip = gap > 0
                rep['pool_flips'] += int(flip.sum())
                rep['pool_total'] += flip.numel()
                rep['pool_worst_margin'] = max(rep['pool_worst_margin'], float(gap.max()))
                x = own
            else:
                if isinstance(m, port.SamePad2d):
                    unpadded = x
                x = m(x)
                if not isinstance(m, port.SamePad2d):
                    unpadded = x
    return rep


def oracle_decisions(omodel, X):
    """The oracle's own decisions in the format of ``device_decisions`` (used to test this module on the CPU)."""
    mods = list(omodel.children())
    out = []
    x = X
    with torch.no_grad():
        for k, m in enumerate(mods):
            if isinstance(m, nn.ReLU):
                out.append(('relu', x > 0))
                x = m(x)
            elif isinstance(m, nn.MaxPool2d):
                N, C, H, W = unpadded.shape
                l, r, t, b = pads
                y, pidx = F.max_pool2d(x, m.kernel_size, m.stride, return_indices=True)      # indices into the padded plane
                Wp = W + l + r
                ih, iw = pidx // Wp - t, pidx % Wp - l
                inside = (ih >= 0) & (ih < H) & (iw >= 0) & (iw < W)
                out.append(('pool', torch.where(inside, ih * W + iw, torch.full_like(pidx, -1)).to(torch.int32)))
                x = y
            else:
                if isinstance(m, port.SamePad2d):
                    unpadded, pads = x, m.pads(x.shape[2], x.shape[3])
                x = m(x)
                if not isinstance(m, port.SamePad2d):
                    unpadded, pads = x, (0, 0, 0, 0)
    return out
