"""Kronecker-factored posterior precision with the public surface of ``preds/kron.py``:
``Kron(model, delta, dampen)`` with ``qhs, Ws, Lams, deltas, dampen`` and
``update / decompose / sample / functional_variance / change_delta``.

Storage is torch tensors on the model's device; the arithmetic (factor axpy, the rotations of
``sample`` and ``functional_variance``) runs in the CUDA library.  The eigendecomposition of each
factor is the one-off posterior factorisation: it calls ``torch.linalg.eigh`` (cuSOLVER syevd) on
the device -- a deliberately chosen library call, timed separately from the accumulation.
"""
import logging

import numpy as np
import torch

from . import _cabi as cabi
from ._cabi import lib, check, ptr, stream


def symeig(M):
    """Eigenvalues (ascending, clamped at 0, NaN -> 0) and eigenvectors from the UPPER triangle,
    with the reference's +I jitter retry (preds/kron.py:146-173).

    On the device the decomposition runs in float64 and is rounded to float32 afterwards: cuSOLVER's float32 syevd leaves
    the eigenvectors of a 1152 x 1152 factor orthogonal only to 3e-4 (measured, CIFAR10Net dense1), which shows up one to
    one as a 3e-4 relative error of the Kron GLM variance against the reference (LAPACK float32 is good to 1e-6).  The
    factorisation is one-off and small next to the accumulation."""
    dt = M.dtype
    Md = M.double() if M.is_cuda else M
    try:
        L, W = torch.linalg.eigh(Md, UPLO='U')
    except RuntimeError:
        logging.info('SYMEIG: adding jitter, did not converge.')
        Mj = Md + torch.eye(M.shape[0], device=M.device, dtype=Md.dtype)
        L, W = torch.linalg.eigh(Mj, UPLO='U')     # a second failure propagates
        L = L - 1.0
    L, W = L.to(dt), W.to(dt)
    L = L.clamp(min=0.0)
    L[torch.isnan(L)] = 0.0
    W[torch.isnan(W)] = 0.0
    return L.contiguous(), W.contiguous()


class Kron:
    def __init__(self, model, delta, dampen=False):
        self.qhs = self.init_factors(model)
        self.Ws = None
        self.Lams = None
        self.deltas = None
        self.change_delta(delta)
        self.dampen = dampen

    @staticmethod
    def init_factors(model):
        """Zero factors per parameter: [QH] for vectors, [Q(out), H(in)] otherwise (:25-43)."""
        qhs = []
        for p in model.parameters():
            if p.ndim == 1:
                qhs.append([torch.zeros(p.shape[0], p.shape[0], device=p.device)])
            else:
                m, n = p.shape[0], int(np.prod(p.shape[1:]))
                qhs.append([torch.zeros(m, m, device=p.device), torch.zeros(n, n, device=p.device)])
        return qhs

    def change_delta(self, new_delta):
        """One prior precision per parameter group, or one for all (:45-56).  Kept on the CPU."""
        if np.isscalar(new_delta):
            new_delta = torch.tensor(new_delta)
        if new_delta.ndim == 0 or (new_delta.ndim == 1 and len(new_delta) == 1):
            self.deltas = new_delta.repeat(len(self.qhs))
        elif new_delta.ndim == 1:
            if len(new_delta) != len(self.qhs):
                raise ValueError('Deltas do not match with parameter groups.')
            self.deltas = new_delta
        else:
            raise ValueError('Invalid dimensionality of delta.')

    def update(self, krons, factor=1., batch_factor=1.):
        """QH[0] += k[0]*factor, QH[1] += k[1]*batch_factor (:58-69)."""
        for kron, QH in zip(krons, self.qhs):
            if len(kron) not in (1, 2):
                raise ValueError('..')
            for k, Q, c in zip(kron, QH, (factor, batch_factor)):
                if Q.is_cuda:
                    k = cabi.f32(k, Q.device)
                    check(lib.bnnp_axpy(Q.numel(), float(c), ptr(k), ptr(Q), stream()), 'bnnp_axpy')
                else:                       # CPU mock models (reference tests): plain torch axpy
                    Q += k * c

    def decompose(self):
        """Eigendecompose every factor (:71-83)."""
        Ws, Lams = [], []
        for qh in self.qhs:
            eig = [symeig(m) for m in qh]
            Lams.append([e[0] for e in eig])
            Ws.append([e[1] for e in eig])
        self.Ws, self.Lams = Ws, Lams

    def _delta(self, i):
        return float(self.deltas[i])

    def sample(self, s, eps=None, mu=None):
        """[s, P] draws from N(0, (Q (x) H + delta I)^-1) (:85-110).  ``eps`` optionally injects
        the standard-normal draws, one tensor per parameter in ``parameters()`` order: [p, s] for
        1-factor groups and [s, m, n] for 2-factor groups.  ``mu`` [P] is added if given."""
        dev = self.Ws[0][0].device
        cabi.require_device(self.Ws[0][0])
        P = sum(len(l[0]) * (len(l[1]) if len(l) == 2 else 1) for l in self.Lams)
        out = torch.empty(s, P, dtype=torch.float32, device=dev)
        off = 0
        for i, (l, w) in enumerate(zip(self.Lams, self.Ws)):
            m = len(l[0])
            if len(l) == 1:
                e = torch.randn(m, s, device=dev) if eps is None else cabi.f32(eps[i], dev)
                assert tuple(e.shape) == (m, s)
                scratch = torch.empty(m * s + 64, device=dev)
                check(lib.bnnp_kron_sample_group(ptr(w[0]), ptr(l[0]), None, None, m, 0, self._delta(i),
                                                 int(self.dampen), ptr(e), s, ptr(mu), ptr(out), P, off,
                                                 ptr(scratch), stream()), 'bnnp_kron_sample_group')
                off += m
            else:
                n = len(l[1])
                e = torch.randn(s, m, n, device=dev) if eps is None else cabi.f32(eps[i], dev)
                assert tuple(e.shape) == (s, m, n)
                done = 0
                chunk = max(1, min(s, 65535, (1 << 28) // (m * n)))
                while done < s:           # bound the scratch (2 * chunk * m * n floats)
                    c = min(chunk, s - done)
                    scratch = torch.empty(2 * c * m * n + m * n + 64, device=dev)
                    ec = e[done:done + c].contiguous()
                    check(lib.bnnp_kron_sample_group(
                        ptr(w[0]), ptr(l[0]), ptr(w[1]), ptr(l[1]), m, n, self._delta(i), int(self.dampen),
                        ptr(ec), c, ptr(mu), out.data_ptr() + done * P * 4, P, off, ptr(scratch), stream()),
                        'bnnp_kron_sample_group')
                    done += c
                off += m * n
        return out

    def functional_variance(self, Js, hess_factor=1.):
        """J Sigma J^T for a MATERIALISED Js [B, K, P] (:112-143) -- API compatibility; the
        predictive path uses the structured ``bnnp_glm_fvar_kron`` instead."""
        B, K, P = Js.size()
        Js = cabi.f32(Js, self.Ws[0][0].device)
        cabi.require_device(Js)
        f_var = torch.zeros(B, K, K, dtype=torch.float32, device=Js.device)
        cur = 0
        for i, (l, w) in enumerate(zip(self.Lams, self.Ws)):
            m = len(l[0])
            l1 = (l[0] * hess_factor).contiguous() if hess_factor != 1. else l[0]
            if len(l) == 1:
                scratch = torch.empty(B * K * m + m + 64, device=Js.device)
                check(lib.bnnp_kron_fvar_block(Js.data_ptr() + cur * 4, P, K, m, 0, ptr(w[0]), ptr(l1), None,
                                               None, self._delta(i), int(self.dampen), ptr(f_var),
                                               ptr(scratch), B * K, stream()), 'bnnp_kron_fvar_block')
                cur += m
            else:
                n = len(l[1])
                scratch = torch.empty(2 * B * K * m * n + m * n + 64, device=Js.device)
                check(lib.bnnp_kron_fvar_block(Js.data_ptr() + cur * 4, P, K, m, n, ptr(w[0]), ptr(l1),
                                               ptr(w[1]), ptr(l[1]), self._delta(i), int(self.dampen),
                                               ptr(f_var), ptr(scratch), B * K, stream()),
                      'bnnp_kron_fvar_block')
                cur += m * n
        return f_var
