"""Plain-autograd stand-in for BackPACK 1.1.1 (``requirements.txt:4`` of the reference).

TEST INFRASTRUCTURE ONLY.  BackPACK's source is not under /root/reference and the package is
not installed, so the quantities the reference reads off the parameters are recomputed here
from their published definitions:

* ``BatchGrad``    -> ``p.grad_batch``      per-sample gradients
* ``DiagGGNExact`` -> ``p.diag_ggn_exact``  exact GGN diagonal
* ``KFLR``         -> ``p.kflr``            Kronecker factors [G (sum over batch), A (batch mean)]

The reference runs its forward passes *outside* the ``with backpack(...)`` block, so the
forward hooks record layer inputs/outputs whenever a module is extended.  Second-order
quantities are computed on ``__enter__`` (graph still alive) rather than re-entrantly inside
a tensor hook.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .context import CTX

_PARAM_LAYERS = (nn.Linear, nn.Conv2d)
_LOSSES = (nn.CrossEntropyLoss, nn.MSELoss)


def _unfold(module, x):
    return F.unfold(x, module.kernel_size, dilation=module.dilation,
                    padding=module.padding, stride=module.stride)


def _first_order(module, grad_out):
    if not any(type(e).__name__ == 'BatchGrad' for e in CTX.active_exts):
        return
    g = grad_out.detach()
    a = module._bp_in.detach()
    N = g.shape[0]
    if isinstance(module, nn.Linear):
        module.weight.grad_batch = torch.einsum('no,ni->noi', g, a)
        if module.bias is not None:
            module.bias.grad_batch = g
    else:
        X = _unfold(module, a)
        go = g.reshape(N, g.shape[1], -1)
        module.weight.grad_batch = torch.einsum('nil,nol->noi', X, go).reshape(
            N, *module.weight.shape)
        if module.bias is not None:
            module.bias.grad_batch = go.sum(2)


def _fwd_hook(module, inp, out):
    if not getattr(module, '_backpack_extend', False):
        return
    if isinstance(module, _PARAM_LAYERS):
        module._bp_in = inp[0]
        module._bp_out = out
        CTX.layer_log.append(module)
        if out.requires_grad:
            out.register_hook(lambda g, m=module: _first_order(m, g))
    elif isinstance(module, _LOSSES):
        CTX.pending = (module, inp[0], list(CTX.layer_log))
        CTX.layer_log = []


def extend(module, debug=False):
    for child in module.children():
        extend(child)
    if getattr(module, '_bp_gen', None) != CTX.generation:
        CTX.handles.append(module.register_forward_hook(_fwd_hook))
        module._bp_gen = CTX.generation
    module._backpack_extend = True
    return module


def memory_cleanup(module):
    for name in ('_bp_in', '_bp_out'):
        if hasattr(module, name):
            delattr(module, name)


def _loss_hessian_sqrt(loss_module, f):
    """S[n, k, v] with S_n S_n^T = Hessian of the (sum-reduced) loss wrt the logits."""
    N, K = f.shape
    if isinstance(loss_module, nn.CrossEntropyLoss):
        p = torch.softmax(f.detach(), dim=1)
        sq = p.sqrt()
        S = torch.diag_embed(sq) - torch.einsum('nk,nv->nkv', p, sq)
    else:
        S = (2.0 ** 0.5) * torch.eye(K, dtype=f.dtype).expand(N, K, K).clone()
    if loss_module.reduction == 'mean':
        S = S / (N ** 0.5)
    return S


def _second_order(exts):
    loss_module, f, layers = CTX.pending
    if not layers:
        return
    N = f.shape[0]
    S = _loss_hessian_sqrt(loss_module, f)
    V = S.shape[2]
    per_v = []
    for v in range(V):
        gs = torch.autograd.grad(f, [m._bp_out for m in layers], grad_outputs=S[:, :, v],
                                 retain_graph=True, allow_unused=True)
        per_v.append(gs)
    want_diag = any(type(e).__name__ == 'DiagGGNExact' for e in exts)
    want_kflr = any(type(e).__name__ == 'KFLR' for e in exts)
    for li, m in enumerate(layers):
        Sl = torch.stack([per_v[v][li] for v in range(V)]).detach()   # [V,N,out,(H,W)]
        a = m._bp_in.detach()
        if isinstance(m, nn.Linear):
            if want_diag:
                m.weight.diag_ggn_exact = torch.einsum('vno,ni->oi', Sl ** 2, a ** 2)
                if m.bias is not None:
                    m.bias.diag_ggn_exact = (Sl ** 2).sum((0, 1))
            if want_kflr:
                G = torch.einsum('vni,vnj->ij', Sl, Sl)
                m.weight.kflr = [G, torch.einsum('ni,nj->ij', a, a) / N]
                if m.bias is not None:
                    m.bias.kflr = [G]
        else:
            X = _unfold(m, a)
            So = Sl.reshape(V, N, Sl.shape[2], -1)
            if want_diag:
                JS = torch.einsum('nil,vnol->vnoi', X, So)
                m.weight.diag_ggn_exact = (JS ** 2).sum((0, 1)).reshape(m.weight.shape)
                if m.bias is not None:
                    m.bias.diag_ggn_exact = (So.sum(3) ** 2).sum((0, 1))
            if want_kflr:
                s = So.sum(3)
                G = torch.einsum('cbi,cbl->il', s, s)
                m.weight.kflr = [G, torch.einsum('bik,bjk->ij', X, X) / N]
                if m.bias is not None:
                    m.bias.kflr = [G]


class backpack:
    def __init__(self, *exts, debug=False):
        self.exts = exts
        self._prev = ()

    def __enter__(self):
        self._prev = CTX.active_exts
        CTX.active_exts = self.exts
        second = [e for e in self.exts if getattr(e, 'second_order', False)]
        if second and CTX.pending is not None:
            _second_order(second)
        return self

    def __exit__(self, *a):
        CTX.active_exts = self._prev
        return False
