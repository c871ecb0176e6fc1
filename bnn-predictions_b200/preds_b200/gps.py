"""Function-space (GP) quantities of the linearised network, API of ``preds/gps.py:664-927``.

``gp_quantities`` builds, per network output c, the NTK Gram matrices of the Jacobian ``J_c`` [n, P]

    K_train[c]      = J_c(X_train) J_c(X_train)^T           [n, n]
    K_train_test[c] = J_c(X_train) J_c(X_test)^T            [n, m]
    K_test[c]       = diag(J_c(X_test) J_c(X_test)^T)       [m]      (the full [m, m] block with only_pred_var=False)
    f_gp[c]         = J_c theta,       f_star[c] = f_c(x)

-- the dual of the GGN contraction: the reduction runs over the P parameters instead of the examples.  The reference
scatters Jacobian batches over ``nn.parallel`` replicas and contracts with ``opt_einsum`` (``Parallel_calc_kernel``,
``preds/gps.py:56-88``).  Here one forward + ONE back-propagated column per output gives ``J_c`` for a chunk of inputs
(``bnnp_jacobians_cols``, V = 1), the chunk is split once into bf16 hi/lo copies and every kernel block is an NT product
``J_row J_col^T`` on the TMA-fed tensor-core GEMM (``bnnp_gemm_nt_split``), the P axis walked in 4096-wide segments that
add through C.  The Jacobian of the training inputs of one output stays resident as bf16 planes (4 bytes per entry: 11 GB
for CIFAR10Net and 3200 inputs), the test inputs stream through in chunks, so nothing of size [n, P, K] ever exists.

**Linear-only networks take a structured route** (``_structured``): the Jacobian of a Linear layer is rank one per output,
``J_c[n, (i, j)] = g_c[n, i] a~[n, j]`` (a~ = layer input with a constant 1 for the bias), so the layer's term of the kernel is
a Hadamard product of two SHORT Grams, ``K_c += (G_c G_c^T) o (A~ A~^T)`` -- reductions over the layer's fan-out and fan-in
instead of over its m x n parameters (63x fewer flops on MLPS(784,[75],10)) -- evaluated as ``A~ A~^T`` once per layer and
one product per output whose epilogue multiplies by it and adds into K_c (``bnnp_gemm_nt_split_had``).  The factors of
ALL outputs come from one forward + one K-column backward per chunk; ``f_gp = J theta`` from ``bnnp_glm_fmu``; J never exists.

Prior-variance groups (``p.log_prior_var``; ``optimize_params`` puts a group's kernel into its own slot, the rest is
summed into the last slot and scaled by its prior variance when ``with_prior_prec``) follow the reference; every group is
a set of contiguous parameter ranges, i.e. of K ranges of the same product.

Not carried over: ``train_jacobians`` / ``save_jacobians`` (host-pinned Jacobian caches of the multi-GPU scatter loop),
``device_ids`` (one process per GPU here; the argument is accepted and ignored).  ``jacobian_method`` selects nothing: both
of the reference's methods give the same Jacobian.
"""
import ctypes as ct

import numpy as np
import torch

from . import _cabi as cabi
from ._cabi import lib, check, ptr, stream
from ._engine import net_for

KSEG = 4096                    # one tensor-core accumulation never spans more (gemm_tma.cu)
STRUCTURED = True              # Linear-only networks: Hadamard-of-Grams kernels (False: materialised one-column Jacobians)
JAC_CHUNK_FLOATS = 1 << 29     # fp32 Jacobian of one chunk of inputs (2 GB)


class _SplitRows:
    """bf16 hi / lo copies of a [rows, P] matrix (rows k-contiguous, pitch a multiple of 64 elements = 128-byte lines)."""

    def __init__(self, rows, P, device):
        self.rows, self.P, self.pitch = rows, P, (P + 63) // 64 * 64
        self.hi = torch.empty(max(rows, 1), self.pitch, dtype=torch.bfloat16, device=device)
        self.lo = torch.empty(max(rows, 1), self.pitch, dtype=torch.bfloat16, device=device)

    def fill(self, J, r0):
        check(lib.bnnp_split_bf16_rows(ptr(J), J.stride(0), J.shape[0], J.shape[1], self.pitch,
                                       self.hi.data_ptr() + 2 * r0 * self.pitch, self.lo.data_ptr() + 2 * r0 * self.pitch,
                                       stream()), 'bnnp_split_bf16_rows')


SPLITK_SCRATCH_FLOATS = 1 << 28       # 1 GB of deterministic split-K partials per product call


def _nt(A, an, B, bn, out):
    """out [an, bn] (fp32, row stride out.stride(0)) = A[:an] B[:bn]^T over all P columns.  One tensor-core accumulation
    never spans more than KSEG columns; a small output with a long P (a few output tiles, hundreds of segments) is ONE call
    whose segments run in parallel as split-K partials (summed in a fixed order), a large output walks P in as many
    segments per call as the partial buffer allows."""
    per_call = max(1, min((A.P + KSEG - 1) // KSEG, SPLITK_SCRATCH_FLOATS // max(1, an * bn)))
    kcall = per_call * KSEG
    scratch = None
    if per_call > 1:
        scratch = torch.empty(int(lib.bnnp_gemm_nt_scratch_floats(an, bn, min(kcall, A.P))) + 64, dtype=torch.float32,
                              device=out.device)
    for s in range(0, A.P, kcall):
        kk = min(kcall, A.P - s)
        check(lib.bnnp_gemm_nt_split(an, bn, kk, 1.0, A.hi.data_ptr() + 2 * s, A.lo.data_ptr() + 2 * s, A.pitch,
                                     B.hi.data_ptr() + 2 * s, B.lo.data_ptr() + 2 * s, B.pitch,
                                     0.0 if s == 0 else 1.0, out.data_ptr(), out.stride(0), ptr(scratch), stream()),
              'bnnp_gemm_nt_split')


def _masked(J, ranges):
    """J with the columns outside ``ranges`` zeroed (one prior-variance slot of several)."""
    M = torch.zeros_like(J)
    for lo_, hi_ in ranges:
        M[:, lo_:hi_] = J[:, lo_:hi_]
    return M


def _groups(model, optimize_params):
    """Parameter ranges per kernel slot, as preds/gps.py:700-735 lays them out: one slot per optimised prior-variance
    tensor, everything else in the last slot (with its prior variance as a scale)."""
    params = list(model.parameters())
    if not all(p.requires_grad for p in params):
        raise cabi.BnnpError('gp_quantities: every parameter must require gradients on the device path')
    assert np.any([hasattr(p, 'log_prior_var') for p in params])
    all_vars = []
    for p in params:
        if hasattr(p, 'log_prior_var') and not any(p.log_prior_var is v for v in all_vars):
            all_vars.append(p.log_prior_var)
    if isinstance(optimize_params, torch.Tensor):
        optimize_params = [optimize_params]
    opt_vars = [v for v in all_vars if any(v is o for o in optimize_params)]
    n_slots = min(len(all_vars), len(opt_vars) + 1)
    slots = []                  # (slot, scale, lo, hi)
    off = 0
    for p in params:
        var = p.log_prior_var
        k = next((i for i, v in enumerate(opt_vars) if v is var), None)
        slot = k if k is not None else n_slots - 1
        scale = None if k is not None else float(var.exp().item())
        if slots and slots[-1][0] == slot and slots[-1][1] == scale and slots[-1][3] == off:
            slots[-1] = (slot, scale, slots[-1][2], off + p.numel())
        else:
            slots.append((slot, scale, off, off + p.numel()))
        off += p.numel()
    return all_vars, opt_vars, n_slots, slots, off


class _Factors:
    """bf16 hi / lo copies of the Jacobian factors of a block of inputs of a Linear-only network: per layer the inputs
    a~ [rows, n_l + 1] and, per requested output c, g_c [rows, m_l]."""

    def __init__(self, net, layers, n_out, rows, device):
        self.layers = layers
        self.A = [_SplitRows(rows, d, device) for (_, d, _, _) in layers]
        self.G = [[_SplitRows(rows, m, device) for _ in range(n_out)] for (_, _, _, m) in layers]

    def fill(self, net, pl, ws, n, r0, outputs):
        base = ws.data_ptr()
        K = net.K
        for li, (col, d, unit, m) in enumerate(self.layers):
            a = base + 4 * (int(pl.acat_off) + col)
            check(lib.bnnp_split_bf16_rows(a, int(pl.Dtot), n, d, self.A[li].pitch, self.A[li].hi.data_ptr() + 2 * r0 * self.A[li].pitch,
                                           self.A[li].lo.data_ptr() + 2 * r0 * self.A[li].pitch, stream()), 'bnnp_split_bf16_rows')
            for oi, c in enumerate(outputs):
                g = base + 4 * (int(pl.gcat_off) + c * int(pl.Mtot) + unit)          # rows (n, v = c) of G_cat
                G = self.G[li][oi]
                check(lib.bnnp_split_bf16_rows(g, K * int(pl.Mtot), n, m, G.pitch, G.hi.data_ptr() + 2 * r0 * G.pitch,
                                               G.lo.data_ptr() + 2 * r0 * G.pitch, stream()), 'bnnp_split_bf16_rows')


def _linear_layers(net):
    """[(first A_cat column, fan-in incl. the bias column, first G_cat unit, fan-out)] of a Linear-only network, else None."""
    pl = net.plan(1, net.K)
    layers = []
    for i in range(net.n_ops):
        op = net.ops[i]
        if op.kind == cabi.OP_CONV2D:
            return None
        if op.kind == cabi.OP_LINEAR:
            d = op.in_c + (1 if op.has_bias else 0)
            if d > KSEG or op.out_c > KSEG:
                return None
            layers.append((int(pl.op[i].lin_col), d, int(pl.op[i].lin_unit), int(op.out_c)))
    return layers or None


def _structured(net, model, X_train, X_test, outputs, theta, scale, out):
    """Fill K_train / K_train_test / K_test (diagonal) / f_gp / f_star in ``out`` for a Linear-only network (one kernel slot)."""
    dev = net.device
    n, m = len(X_train), (len(X_test) if X_test is not None else 0)
    C_ = len(outputs)
    chunk = 2048
    layers = None

    def factors(X, s, F, r0):
        nonlocal layers
        X2d, pl, ws, f = net.jacobian_factors(X[s:s + chunk])
        k = X2d.shape[0]
        if layers is None:
            layers = _linear_layers(net)
        if F is None:
            F = _Factors(net, layers, C_, n if r0 is not None else min(chunk, max(m, 1)), dev)
        F.fill(net, pl, ws, k, r0 if r0 is not None else 0, outputs)
        fm = torch.empty(k, net.K, dtype=torch.float32, device=dev)
        check(lib.bnnp_glm_fmu(net.ops, net.n_ops, ct.byref(pl), k, ptr(X2d), ptr(ws), ptr(f), ptr(theta), ptr(fm), stream()),
              'bnnp_glm_fmu')
        Av = ws[int(pl.acat_off):int(pl.acat_off) + k * int(pl.Dtot)].view(k, int(pl.Dtot))
        Gv = ws[int(pl.gcat_off):int(pl.gcat_off) + k * net.K * int(pl.Mtot)].view(k, net.K, int(pl.Mtot))
        diag = None
        if r0 is None:            # test block: the diagonal of its own kernel, sum_l |g_c|^2 |a~|^2
            diag = torch.zeros(C_, k, dtype=torch.float32, device=dev)
            for (col, d, unit, mm) in layers:
                a2 = (Av[:, col:col + d] ** 2).sum(1)
                g2 = (Gv[:, outputs, unit:unit + mm] ** 2).sum(2)
                diag += (g2 * a2.unsqueeze(1)).t()
        return F, k, f, (fm - f), diag

    def products(Fr, nr, Fc, nc, dst, col0, ldc, tmp):
        """dst[c][:, col0:col0+nc] = sum_l (G_c G_c^T) o (A~ A~^T) for the row block Fr and the column block Fc."""
        AA = tmp[:nr * nc].view(nr, nc)
        for li in range(len(layers)):
            d, mm = layers[li][1], layers[li][3]
            Ar, Ac = Fr.A[li], Fc.A[li]
            check(lib.bnnp_gemm_nt_split(nr, nc, d, 1.0, Ar.hi.data_ptr(), Ar.lo.data_ptr(), Ar.pitch, Ac.hi.data_ptr(),
                                         Ac.lo.data_ptr(), Ac.pitch, 0.0, AA.data_ptr(), nc, None, stream()), 'bnnp_gemm_nt_split')
            for oi in range(C_):
                Gr, Gc = Fr.G[li][oi], Fc.G[li][oi]
                check(lib.bnnp_gemm_nt_split_had(nr, nc, mm, scale, Gr.hi.data_ptr(), Gr.lo.data_ptr(), Gr.pitch,
                                                 Gc.hi.data_ptr(), Gc.lo.data_ptr(), Gc.pitch, 0.0 if li == 0 else 1.0,
                                                 dst[oi].data_ptr() + 4 * col0, ldc, AA.data_ptr(), nc, stream()),
                      'bnnp_gemm_nt_split_had')

    Ftr = None
    for s in range(0, n, chunk):
        Ftr, k, f, fgp, _ = factors(X_train, s, Ftr, s)
        out['f_star_train'][:, s:s + k] = f.t()[outputs]
        out['f_gp_train'][:, s:s + k] = fgp.t()[outputs]
    tmp = torch.empty(max(n, 1) * max(n, min(chunk, max(m, 1))), dtype=torch.float32, device=dev)
    products(Ftr, n, Ftr, n, [out['K_train'][c, 0] for c in range(C_)], 0, n, tmp)
    if X_test is None:
        return
    Fte = None
    for s in range(0, m, chunk):
        Fte, k, f, fgp, diag = factors(X_test, s, Fte, None)
        out['f_star_test'][:, s:s + k] = f.t()[outputs]
        out['f_gp_test'][:, s:s + k] = fgp.t()[outputs]
        out['K_test'][:, 0, s:s + k] = scale * diag
        products(Ftr, n, Fte, k, [out['K_train_test'][c, 0] for c in range(C_)], s, m, tmp)


def gp_quantities_device(model, X_train, X_test=None, outputs=-1, optimize_params=(), only_pred_var=True,
                         collapse_groups=False, with_prior_prec=True, max_ram_gb=10):
    """``gp_quantities`` with every tensor left on the model's device (``FunctionaLaplace`` keeps them there)."""
    net = net_for(model)
    dev = next(model.parameters()).device
    cabi.require_device(next(model.parameters()))
    K_out = model.output_size
    if isinstance(outputs, int):
        if 0 <= outputs < K_out:
            outputs = [outputs]
        elif outputs == -1:
            outputs = list(range(K_out))
        else:
            raise ValueError('Incorrect outputs argument value: {}'.format(outputs))
    elif not all(0 <= o < K_out for o in outputs):
        raise ValueError('Incorrect outputs argument value: {}'.format(outputs))
    all_vars, opt_vars, n_slots, slots, P = _groups(model, optimize_params)
    n, m = len(X_train), (len(X_test) if X_test is not None else 0)
    C = len(outputs)
    total = (n * n + n * m + (m if only_pred_var else m * m)) * n_slots * C
    if max_ram_gb * 1024 ** 3 <= 4 * total:
        raise RuntimeError('More memory needed to store kernels!')

    theta = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).to(torch.float32)
    colscale = None
    if with_prior_prec and any(sc is not None and sc != 1.0 for _, sc, _, _ in slots):
        colscale = torch.ones(P, dtype=torch.float32, device=dev)
        for _, sc, lo_, hi_ in slots:
            if sc is not None:
                colscale[lo_:hi_] = sc ** 0.5
    ranges = [[(lo_, hi_) for s, _, lo_, hi_ in slots if s == g] for g in range(n_slots)]

    f32 = dict(dtype=torch.float32, device=dev)
    K_train = torch.zeros(C, n_slots, n, n, **f32)
    f_gp_train, f_star_train = torch.ones(C, n, **f32), torch.ones(C, n, **f32)
    if X_test is not None:
        K_train_test = torch.zeros(C, n_slots, n, m, **f32)
        K_test = torch.zeros(*((C, n_slots, m) if only_pred_var else (C, n_slots, m, m)), **f32)
        f_gp_test, f_star_test = torch.ones(C, m, **f32), torch.ones(C, m, **f32)
    uniform = None            # the prior-variance scale when it is one number for the whole kernel
    if n_slots == 1:
        scs = set(sc for _, sc, _, _ in slots)
        if not with_prior_prec or scs == {None}:
            uniform = 1.0
        elif len(scs) == 1:
            uniform = float(next(iter(scs)))
    net.prepare(cabi.f32(X_train[:1], dev))
    if STRUCTURED and uniform is not None and only_pred_var and _linear_layers(net) is not None:
        res = {'K_train': K_train, 'f_gp_train': f_gp_train, 'f_star_train': f_star_train}
        if X_test is not None:
            res.update({'K_train_test': K_train_test, 'K_test': K_test, 'f_gp_test': f_gp_test, 'f_star_test': f_star_test})
        _structured(net, model, X_train, X_test, outputs, theta, uniform, res)
        net.release()
        return _finish(K_train, K_train_test if X_test is not None else None, K_test if X_test is not None else None,
                       f_gp_train, f_star_train, f_gp_test if X_test is not None else None,
                       f_star_test if X_test is not None else None, all_vars, opt_vars, collapse_groups)
    chunk = max(1, min(1024, JAC_CHUNK_FLOATS // max(P, 1)))
    # one plane set per kernel slot (a single one unless optimize_params separates prior-variance groups)
    Jtr = [_SplitRows(n, P, dev) for _ in range(n_slots)]
    keep_test = X_test is not None and not only_pred_var
    Jte = [_SplitRows(m if keep_test else min(chunk, max(m, 1)), P, dev) for _ in range(n_slots)] if X_test is not None else None
    tmp = torch.empty(max(n, 1) * max(chunk, 1), **f32)

    def jac(X, s, c):
        J, f = net.output_jacobian(X[s:s + chunk], c)
        fgp = J @ theta
        if colscale is not None:
            J.mul_(colscale)
        return J, f, fgp

    def slot(J, g):
        return J if n_slots == 1 else _masked(J, ranges[g])

    for idx, c in enumerate(outputs):
        for s in range(0, n, chunk):
            J, f, fgp = jac(X_train, s, c)
            k = J.shape[0]
            if idx == 0:
                f_star_train[:, s:s + k] = f.t()[outputs]
            f_gp_train[idx, s:s + k] = fgp
            for g in range(n_slots):
                Jtr[g].fill(slot(J, g), s)
            del J
        for g in range(n_slots):
            _nt(Jtr[g], n, Jtr[g], n, K_train[idx, g])
        if X_test is None:
            continue
        for s in range(0, m, chunk):
            J, f, fgp = jac(X_test, s, c)
            k = J.shape[0]
            if idx == 0:
                f_star_test[:, s:s + k] = f.t()[outputs]
            f_gp_test[idx, s:s + k] = fgp
            for g in range(n_slots):
                Jg = slot(J, g)
                if keep_test:
                    Jte[g].fill(Jg, s)
                    continue
                K_test[idx, g, s:s + k] = (Jg * Jg).sum(1)
                Jte[g].fill(Jg, 0)
                out = tmp[:n * k].view(n, k)
                _nt(Jtr[g], n, Jte[g], k, out)
                K_train_test[idx, g, :, s:s + k] = out
            del J
        if keep_test:
            for g in range(n_slots):
                _nt(Jtr[g], n, Jte[g], m, K_train_test[idx, g])
                _nt(Jte[g], m, Jte[g], m, K_test[idx, g])
    net.release()
    return _finish(K_train, K_train_test if X_test is not None else None, K_test if X_test is not None else None,
                   f_gp_train, f_star_train, f_gp_test if X_test is not None else None,
                   f_star_test if X_test is not None else None, all_vars, opt_vars, collapse_groups)


def _finish(K_train, K_train_test, K_test, f_gp_train, f_star_train, f_gp_test, f_star_test, all_vars, opt_vars,
            collapse_groups):
    def collapse(K):
        # preds/gps.py:894-908
        if len(opt_vars) > 0:
            if K.shape[1] > 1:
                kernel = K[:, -1].clone() if len(opt_vars) < len(all_vars) else torch.zeros_like(K[:, -1])
                for i, v in enumerate(opt_vars):
                    kernel += K[:, i] * float(v.exp().detach())
            else:
                kernel = K[:, 0].clone() * float(opt_vars[0].exp().detach())
        else:
            kernel = K[:, 0].clone()
        return kernel

    q = {'K_train': collapse(K_train) if collapse_groups else K_train,
         'f_gp_train': f_gp_train, 'f_star_train': f_star_train,
         'prior_var': (float(all_vars[0].exp().item()) if len(all_vars) == 1 else None) if collapse_groups else None}
    if K_train_test is not None:
        q['K_train_test'] = collapse(K_train_test) if collapse_groups else K_train_test
        q['K_test'] = collapse(K_test) if collapse_groups else K_test
        q['f_gp_test'] = f_gp_test
        q['f_star_test'] = f_star_test
    return q


def gp_quantities(model, X_train, X_test=None, outputs=-1, optimize_params=set(), only_pred_var=True, max_ram_gb=10,
                  jacobian_method='batched', save_jacobians=False, train_jacobians=None, collapse_groups=False,
                  with_prior_prec=True, batch_size_per_gpu=-1, device_ids=None, print_progress=False):
    """Drop-in for ``preds.gps.gp_quantities`` (preds/gps.py:664): same arguments, same dictionary, tensors on the CPU as
    the reference returns them."""
    assert isinstance(X_train, torch.Tensor)
    if X_test is not None:
        assert isinstance(X_test, torch.Tensor)
    assert jacobian_method in ('naive', 'batched')
    if save_jacobians or train_jacobians is not None:
        raise NotImplementedError('save_jacobians / train_jacobians: the device path keeps no host copy of the Jacobians')
    q = gp_quantities_device(model, X_train, X_test, outputs, optimize_params, only_pred_var, collapse_groups,
                             with_prior_prec, max_ram_gb)
    return {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in q.items()}


# ---------------------------------------------------------------------------------------------------------------------
# Predictive distribution of the GP at the test inputs (preds/laplace.py:175-207, preds/gps.py:384-447)
# ---------------------------------------------------------------------------------------------------------------------
def _mm(ta, tb, A, B, out):
    """out = op(A) op(B) on the library's tensor-core GEMM (fp32 operands, split on chip), K walked in segments."""
    from . import dense
    M, N = out.shape
    K = A.shape[0] if ta else A.shape[1]
    dense._gemm(ta, tb, M, N, K, 1.0, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), 0.0, out.data_ptr(), out.stride(0))
    return out


def gp_predictive_moments(q, prior_prec, indep=False):
    """(f_mu [m, C], f_var [m, C, C]) of the function-space posterior at the test inputs for a categorical likelihood.

    indep=True  (preds/laplace.py:186-199): one independent binary-style GP per class, W = p (1 - p) with p the softmax over
                classes of the network outputs; f_var is diagonal, K_** - K_*n (K + W^-1)^-1 K_n*.
    indep=False (``GP_predictive``, preds/gps.py:411-426): the multi-class Laplace GP posterior (Rasmussen & Williams
                alg. 3.4): E_c = D_c^1/2 (I + D_c^1/2 K_c D_c^1/2)^-1 D_c^1/2, M = chol(sum_c E_c), b_c = E_c K_c*,
                Sigma_* = diag(k_** - k_*^T b) + b^T (sum E)^-1 b.
    The mean is the network output itself (``fdist.loc = f_star_test.T``, preds/laplace.py:205).  The n x n factorisations
    go through torch.linalg (one-off, n = #training inputs of the subset); every [n, n] x [n, m] product runs on the
    library's tensor-core GEMM."""
    K_nn = q['K_train'] / prior_prec
    K_nm = q['K_train_test'] / prior_prec
    K_m = q['K_test'] / prior_prec
    m_test, m_train = q['f_star_test'], q['f_star_train']
    C, n, m = K_nm.shape
    dev = K_nn.device
    f_mu = m_test.t().contiguous()
    f_var = torch.zeros(m, C, C, dtype=torch.float32, device=dev)
    T = torch.empty(n, m, dtype=torch.float32, device=dev)
    if indep:
        p = torch.softmax(m_train, dim=0)
        w_sqrt = torch.sqrt(p * (1 - p))
        for c in range(C):
            B = w_sqrt[c].unsqueeze(1) * K_nn[c] * w_sqrt[c].unsqueeze(0)
            B.diagonal().add_(1.0)
            # B = I + W^1/2 K W^1/2 is symmetric positive definite: its inverse through a Cholesky factorisation
            # (the reference calls torch.inverse, an LU with pivoting, preds/laplace.py:193)
            Kinv = w_sqrt[c].unsqueeze(1) * torch.cholesky_inverse(torch.linalg.cholesky(B)) * w_sqrt[c].unsqueeze(0)
            _mm(0, 0, Kinv.contiguous(), K_nm[c], T)
            f_var[:, c, c] = K_m[c] - (K_nm[c] * T).sum(0)
        return f_mu, f_var
    pi = torch.softmax(m_train, dim=0)                   # inv_link((f_gp + (f_star - f_gp)).T).T
    E = torch.empty(C, n, n, dtype=torch.float32, device=dev)
    for c in range(C):
        d = pi[c].sqrt()
        A = K_nn[c] * d.unsqueeze(0) * d.unsqueeze(1)
        A.diagonal().add_(1.0)
        L = torch.linalg.cholesky(A)
        E[c] = torch.cholesky_solve(torch.diag_embed(d), L) * d.unsqueeze(1)
    Minv = torch.cholesky_inverse(torch.linalg.cholesky(E.sum(0))).contiguous()
    b = torch.empty(C, n, m, dtype=torch.float32, device=dev)
    cs = torch.empty(C, n, m, dtype=torch.float32, device=dev)
    for c in range(C):
        _mm(1, 0, E[c], K_nm[c], b[c])
        _mm(0, 0, Minv, b[c], cs[c])
    for c in range(C):
        for d_ in range(C):
            f_var[:, c, d_] = (b[c] * cs[d_]).sum(0)
        f_var[:, c, c] += K_m[c] - (K_nm[c] * b[c]).sum(0)
    return f_mu, f_var


def mc_mean(likelihood, f_mu, f_var, n_samples=100, batch_size=100, eps=None, seed=None):
    """First return value of ``mc_preds`` (preds/gps.py:930-952): the Monte-Carlo mean of softmax(f), f ~ N(f_mu, f_var),
    from the GLM sampler's mean variant (``bnnp_glm_sample_mean``: the samples are summed in registers, [m, C] is all that
    is written).  eps [S, m, C]: injected standard normals, consumed as f = f_mu[i] + chol(f_var[i]) eps[s, i]."""
    assert n_samples % min(n_samples, batch_size) == 0, \
        'Number of samples should be smaller than batch size or divisible by batch size'
    N, K = f_mu.shape
    S = int(n_samples)
    out = torch.empty(N, K, dtype=torch.float32, device=f_mu.device)
    status = torch.empty(N, dtype=torch.int32, device=f_mu.device)
    if eps is not None:
        eps = cabi.f32(eps, f_mu.device)
        assert tuple(eps.shape) == (S, N, K)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    check(lib.bnnp_glm_sample_mean(likelihood.code, ptr(f_mu), ptr(f_var), ptr(eps), seed, S, N, K, ptr(out),
                                   ptr(status), stream()), 'bnnp_glm_sample_mean')
    if bool(status.any()):
        raise RuntimeError('function-space predictive covariance is not positive definite')
    return out
