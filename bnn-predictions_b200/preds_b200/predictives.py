"""GLM and BNN predictives with the signatures of ``preds/predictives.py`` plus the additive
``eps=`` / ``seed=`` arguments used to inject the standard-normal draws (parity with the oracle).
"""
import ctypes as C
import logging

import torch
from torch.nn.utils import parameters_to_vector

from . import _cabi as cabi
from ._cabi import lib, check, ptr, stream
from ._engine import net_for, ptr_array, float_array
from .kron import Kron
from .likelihoods import GaussianLh
from . import dense

# upper bound for per-call scratch; inputs are processed in chunks that respect it
SCRATCH_BUDGET_FLOATS = 1 << 30
# full-covariance GLM engines: 1 SIMT (as written), 2 tcgen05 with generated operands, 3 tcgen05 + TMA
FULL_TMA_MIN_P = 1024


def sigma_split(Sigma):
    """bf16 hi/lo copies of the covariance (row pitch roundup8(P)) for the TMA-fed GLM kernel, made once per
    posterior.  The copies live ON the covariance tensor object (attribute ``_bnnp_split``) together with the tensor
    version they were made from: they die with the posterior (``Laplace.infer`` drops ``_Sigma``), an in-place update of
    Sigma invalidates them, and a NEW covariance that the caching allocator places at a freed one's address can never
    pick up the old split (a cache keyed on ``data_ptr`` could)."""
    ent = getattr(Sigma, '_bnnp_split', None)
    if ent is None or ent[0] != Sigma._version:
        Sigma._bnnp_split = None
        P = Sigma.shape[0]
        pitch = (P + 7) // 8 * 8
        hi = torch.empty(P, pitch, dtype=torch.bfloat16, device=Sigma.device)
        lo = torch.empty(P, pitch, dtype=torch.bfloat16, device=Sigma.device)
        check(lib.bnnp_split_bf16(ptr(Sigma), P, P, P, ptr(hi), ptr(lo), stream()), 'bnnp_split_bf16')
        ent = (Sigma._version, hi, lo)
        Sigma._bnnp_split = ent
    return ent[1], ent[2]


def _kron_tables(net, kron):
    """Per-op eigen pointers {W_G, lam_G, W_A, lam_A} and prior precisions for weight / bias."""
    index = {id(p): i for i, p in enumerate(net.params)}
    eig, dw, db = [], [], []
    for op, mod in zip(net.ops, net.mods):
        if op.kind in (cabi.OP_LINEAR, cabi.OP_CONV2D):
            wi = index[id(mod.weight)]
            eig += [ptr(kron.Ws[wi][0]), ptr(kron.Lams[wi][0]), ptr(kron.Ws[wi][1]), ptr(kron.Lams[wi][1])]
            dw.append(float(kron.deltas[wi]))
            db.append(float(kron.deltas[index[id(mod.bias)]]) if mod.bias is not None else 0.0)
        else:
            eig += [None] * 4
            dw.append(0.0)
            db.append(0.0)
    return ptr_array(eig), float_array(dw), float_array(db)


def glm_moments(X, model, mu, Sigma, engine=0, return_f=False):
    """f_mu [N,K] and f_var [N,K,K] of the linearised model (preds/predictives.py:52-65).
    ``Sigma`` is the full covariance [P,P], the diagonal variances [P], or a ``Kron``.
    ``return_f`` also returns the network output f [N,K] at the linearisation point."""
    net = net_for(model)
    if torch.is_tensor(Sigma) and Sigma.dim() == 2 and not Sigma.is_contiguous() and Sigma.T.is_contiguous():
        Sigma = Sigma.T                # covariance is symmetric: take the contiguous view, no copy
    theta_star = parameters_to_vector(model.parameters()).detach()
    delta = (mu - theta_star).to(torch.float32).contiguous()
    delta_ptr = ptr(delta) if bool((delta != 0).any()) else None
    X = cabi.f32(X, net.device)
    N, K, P = X.shape[0], None, net.P
    net.prepare(X[:1])
    K = net.K
    # chunk size from the scratch the chosen structure needs: a constant part (weight-sized buffers) plus a part per input,
    # separated by evaluating the library's own sizing function at two batch sizes
    def per_input(fn):
        lo, hi = int(fn(1)), int(fn(257))
        per = max(1, (hi - lo) // 256)
        return per, max(0, lo - per)

    if isinstance(Sigma, Kron):
        per, const = per_input(lambda n: lib.bnnp_glm_fvar_kron_scratch_floats(net.ops, net.n_ops, C.byref(net.plan(n, K)), n))
        cap = 65535 // K if not net.linear_only else 1 << 30      # the batched conv kernels index rows with 16 bits
    elif Sigma.dim() == 2:
        if engine == 0:
            engine = 3 if P >= FULL_TMA_MIN_P else 1
        if engine == 3:
            per, const = per_input(lambda n: lib.bnnp_glm_fvar_full_tma_scratch_floats(net.ops, net.n_ops, C.byref(net.plan(n, K)), n))
            split = sigma_split(Sigma)
        else:
            per, const = per_input(lambda n: lib.bnnp_glm_fvar_full_scratch_floats(C.byref(net.plan(n, K)), n, engine))
        cap = 1 << 30
    else:
        per, const = per_input(lambda n: lib.bnnp_glm_fvar_diag_scratch_floats(net.ops, net.n_ops, C.byref(net.plan(n, K)), n))
        cap = 65535 // K if not net.linear_only else 1 << 30
    # the layer-factor workspace of the same chunk counts against the budget too
    wper = max(1, (int(net.plan(257, K).total_floats) - int(net.plan(1, K).total_floats)) // 256)
    per += wper
    chunk = max(1, min(N, cap, max(1, SCRATCH_BUDGET_FLOATS - const) // per))
    f_mu = torch.empty(N, K, dtype=torch.float32, device=net.device)
    f_var = torch.empty(N, K, K, dtype=torch.float32, device=net.device)
    f_all = torch.empty(N, K, dtype=torch.float32, device=net.device) if return_f else None
    tables = _kron_tables(net, Sigma) if isinstance(Sigma, Kron) else None
    for s in range(0, N, chunk):
        Xc = X[s:s + chunk]
        n = Xc.shape[0]
        X2d, pl, ws, f = net.jacobian_factors(Xc)
        fm, fv = f_mu[s:s + n], f_var[s:s + n]
        if return_f:
            f_all[s:s + n] = f
        check(lib.bnnp_glm_fmu(net.ops, net.n_ops, C.byref(pl), n, ptr(X2d), ptr(ws), ptr(f), delta_ptr,
                               ptr(fm), stream()), 'bnnp_glm_fmu')
        if isinstance(Sigma, Kron):
            sc = net.scratch(lib.bnnp_glm_fvar_kron_scratch_floats(net.ops, net.n_ops, C.byref(pl), n))
            check(lib.bnnp_glm_fvar_kron(net.ops, net.n_ops, C.byref(pl), n, ptr(X2d), ptr(ws), tables[0],
                                         tables[1], tables[2], int(Sigma.dampen), ptr(fv), ptr(sc), stream()),
                  'bnnp_glm_fvar_kron')
        elif Sigma.dim() == 2 and engine == 3:
            sc = net.scratch(lib.bnnp_glm_fvar_full_tma_scratch_floats(net.ops, net.n_ops, C.byref(pl), n))
            check(lib.bnnp_glm_fvar_full_tma(net.ops, net.n_ops, C.byref(pl), n, ptr(ws), ptr(Sigma), ptr(split[0]),
                                             ptr(split[1]), ptr(fv), ptr(sc), stream()), 'bnnp_glm_fvar_full_tma')
        elif Sigma.dim() == 2:
            sc = net.scratch(lib.bnnp_glm_fvar_full_scratch_floats(C.byref(pl), n, engine))
            check(lib.bnnp_glm_fvar_full(net.ops, net.n_ops, C.byref(pl), n, ptr(ws), ptr(Sigma), ptr(fv),
                                         ptr(sc), engine, stream()), 'bnnp_glm_fvar_full')
        else:
            sc = net.scratch(lib.bnnp_glm_fvar_diag_scratch_floats(net.ops, net.n_ops, C.byref(pl), n))
            check(lib.bnnp_glm_fvar_diag(net.ops, net.n_ops, C.byref(pl), n, ptr(X2d), ptr(ws), ptr(Sigma),
                                         ptr(fv), ptr(sc), stream()), 'bnnp_glm_fvar_diag')
    return (f_mu, f_var, f_all) if return_f else (f_mu, f_var)


def functional_sampling_predictive(X, model, likelihood, mu, Sigma, mc_samples=1000, no_link=False,
                                   eps=None, seed=None):
    """GLM predictive (preds/predictives.py:50-76): regression returns (f_mu, f_var);
    classification returns link(f_mu + chol(f_var) eps) of shape [S, N, K]."""
    if not isinstance(Sigma, Kron):
        Sigma = cabi.f32(Sigma, mu.device)
    f_mu, f_var = glm_moments(X, model, mu, Sigma)
    if type(likelihood) is GaussianLh:
        return f_mu, f_var
    N, K = f_mu.shape
    S = int(mc_samples)
    lik = cabi.LIK_GAUSSIAN if no_link else likelihood.code
    out = torch.empty(S, N, K, dtype=torch.float32, device=f_mu.device)
    status = torch.empty(N, dtype=torch.int32, device=f_mu.device)
    if eps is not None:
        eps = cabi.f32(eps, f_mu.device)
        assert tuple(eps.shape) == (S, N, K)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())     # advances the global torch RNG
    check(lib.bnnp_glm_sample(lik, ptr(f_mu), ptr(f_var), ptr(eps), seed, S, N, K, ptr(out), ptr(status),
                              stream()), 'bnnp_glm_sample')
    if bool(status.any()):
        # MultivariateNormal would have raised for the whole batch (:70-76)
        logging.warning('functional sampling covariance indefinite - use diagonal')
        check(lib.bnnp_glm_sample_diag_fallback(lik, ptr(f_mu), ptr(f_var), ptr(eps), seed, S, N, K, ptr(out),
                                                stream()), 'bnnp_glm_sample_diag_fallback')
    return out


def functional_mean_predictive(X, model, likelihood, mu, Sigma, mc_samples=1000, eps=None, seed=None):
    """``functional_sampling_predictive(...).mean(0)`` -- what every driver does with the samples
    (experiments/imginference.py:391, classification.py:36) -- with the mean over samples taken inside the
    sampling kernel: [N,K], no [S,N,K] tensor (400 MB at config 3).  Additive API; classification only."""
    if type(likelihood) is GaussianLh:
        raise ValueError('the regression predictive returns moments, not samples (preds/predictives.py:66-67)')
    if not isinstance(Sigma, Kron):
        Sigma = cabi.f32(Sigma, mu.device)
    f_mu, f_var = glm_moments(X, model, mu, Sigma)
    N, K = f_mu.shape
    S = int(mc_samples)
    out = torch.empty(N, K, dtype=torch.float32, device=f_mu.device)
    status = torch.empty(N, dtype=torch.int32, device=f_mu.device)
    if eps is not None:
        eps = cabi.f32(eps, f_mu.device)
        assert tuple(eps.shape) == (S, N, K)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    check(lib.bnnp_glm_sample_mean(likelihood.code, ptr(f_mu), ptr(f_var), ptr(eps), seed, S, N, K, ptr(out),
                                   ptr(status), stream()), 'bnnp_glm_sample_mean')
    if bool(status.any()):      # indefinite f_var somewhere: the reference's whole-batch diagonal fallback (:73-76)
        logging.warning('functional sampling covariance indefinite - use diagonal')
        full = torch.empty(S, N, K, dtype=torch.float32, device=f_mu.device)
        check(lib.bnnp_glm_sample_diag_fallback(likelihood.code, ptr(f_mu), ptr(f_var), ptr(eps), seed, S, N, K,
                                                ptr(full), stream()), 'bnnp_glm_sample_diag_fallback')
        return full.mean(0)
    return out


def linear_regression_predictive(X, model, likelihood, mu, Sigma_chol):
    """preds/predictives.py:79-94: (mu_star, var_f, var_noise) of the linearised model pushed through the inverse
    link by the delta method; ``Sigma_chol`` is a lower factor [P,P] or standard deviations [P]."""
    Sigma_chol = cabi.f32(Sigma_chol, mu.device)
    if Sigma_chol.dim() == 2:
        P = Sigma_chol.shape[0]
        # Sigma = L L^T with the K range walked in 4096-wide segments accumulated through C (no split-K workspace:
        # that would be several extra [P,P] buffers where the reference needs one)
        Sigma = torch.empty(P, P, dtype=torch.float32, device=mu.device)
        dense._gemm(0, 1, P, P, P, 1.0, Sigma_chol.data_ptr(), P, Sigma_chol.data_ptr(), P, 0.0, Sigma.data_ptr(), P)
    else:
        Sigma = (Sigma_chol * Sigma_chol).contiguous()
    f_mu, f_var, f = glm_moments(X, model, cabi.f32(mu, mu.device), Sigma, return_f=True)
    N, K = f.shape
    mu_star = torch.empty_like(f)
    var_f, var_noise = torch.empty_like(f_var), torch.empty_like(f_var)
    check(lib.bnnp_lik_link_moments(likelihood.code, ptr(f), ptr(f_mu), ptr(f_var), float(likelihood.sigma_2_dispersion),
                                    N, K, ptr(mu_star), ptr(var_f), ptr(var_noise), stream()), 'bnnp_lik_link_moments')
    if net_for(model).flat_output:
        mu_star = mu_star.reshape(N)
    return mu_star, var_f.squeeze(), var_noise.squeeze()


def sample_weights(mu, Sigma_chol, mc_samples, eps=None):
    """theta samples [S, P] (preds/predictives.py:14-20)."""
    S, P = int(mc_samples), len(mu)
    mu = cabi.f32(mu, mu.device)
    if isinstance(Sigma_chol, Kron):
        return Sigma_chol.sample(S, eps=eps, mu=mu)
    Sigma_chol = cabi.f32(Sigma_chol, mu.device)
    e = torch.randn(P, S, device=mu.device) if eps is None else cabi.f32(eps, mu.device)
    assert tuple(e.shape) == (P, S)
    out = torch.empty(S, P, dtype=torch.float32, device=mu.device)
    if Sigma_chol.dim() == 2:
        check(lib.bnnp_sample_theta_full(ptr(Sigma_chol), ptr(mu), ptr(e), P, S, ptr(out), stream()),
              'bnnp_sample_theta_full')
    else:
        check(lib.bnnp_sample_theta_diag(ptr(Sigma_chol), ptr(mu), ptr(e), P, S, ptr(out), stream()),
              'bnnp_sample_theta_diag')
    return out


def nn_sampling_predictive(X, model, likelihood, mu, Sigma_chol, mc_samples=100, no_link=False,
                           eps=None):
    """BNN predictive (preds/predictives.py:12-28): S forward passes at sampled weights,
    [S, N, K].  The model's own parameters are never overwritten (the sampled weights are read
    from the sample matrix), so there is nothing to restore."""
    net = net_for(model)
    samples = sample_weights(mu, Sigma_chol, mc_samples, eps=eps)
    X2d = net.prepare(X)
    N, S = X2d.shape[0], samples.shape[0]
    lik = cabi.LIK_GAUSSIAN if no_link else likelihood.code
    out = torch.empty(S, N, net.K, dtype=torch.float32, device=net.device)
    sc = net.scratch(lib.bnnp_bnn_forward_scratch_floats(net.ops, net.n_ops, N, S))
    check(lib.bnnp_bnn_forward(net.ops, net.n_ops, lik, ptr(X2d), N, ptr(samples), S, net.P, ptr(out),
                               ptr(sc), stream()), 'bnnp_bnn_forward')
    return out.reshape(S, N) if net.flat_output else out


def linear_sampling_predictive(X, model, likelihood, mu, Sigma_chol, mc_samples=100, no_link=False,
                               eps=None):
    """GLM predictive by weight sampling (preds/predictives.py:31-47): link(f(x) + J(x)(theta_s - theta*))
    for theta_s = mu + Sigma_chol eps, eps [P,S]; [S, N, K] ([S, N] for a model that flattens its single
    output).  ``Sigma_chol`` is a lower factor [P,P] (also the diagonal *matrix* of ``get_diagonal_ggn``)
    or a vector of standard deviations [P].  All S samples go through one batched layer product per
    layer on the Jacobian factors; J and the [S,P]x[P,N,K] product of the reference are never formed.

    Deviation: for a single-output model that does NOT flatten (``MLPS(..., 1)``: f is [N,1], Js is [N,P]) the
    reference's ``f - Js @ theta`` broadcasts [N,1] - [N] to [N,N] (preds/predictives.py:36); here that case returns
    the intended [S, N, 1]."""
    net = net_for(model)
    theta_star = parameters_to_vector(model.parameters()).detach()
    mu = cabi.f32(mu, net.device)
    dtheta = sample_weights(mu - theta_star, Sigma_chol, mc_samples, eps=eps)       # theta_s - theta* [S,P]
    X = cabi.f32(X, net.device)
    net.prepare(X[:1])
    if not net.linear_only:
        raise cabi.BnnpError('linear_sampling_predictive is implemented for Linear-only networks')
    N, K, S = X.shape[0], net.K, dtheta.shape[0]
    lik = cabi.LIK_GAUSSIAN if no_link else likelihood.code
    out = torch.empty(S, N, K, dtype=torch.float32, device=net.device)
    per = max(1, int(lib.bnnp_glm_linear_samples_scratch_floats(C.byref(net.plan(64, K)), 64, S)) // 64)
    chunk = max(1, min(N, SCRATCH_BUDGET_FLOATS // per))
    for s0 in range(0, N, chunk):
        Xc = X[s0:s0 + chunk]
        n = Xc.shape[0]
        _, pl, ws, f = net.jacobian_factors(Xc)
        sc = net.scratch(lib.bnnp_glm_linear_samples_scratch_floats(C.byref(pl), n, S))
        if n == N:
            dst = out
        else:
            dst = torch.empty(S, n, K, dtype=torch.float32, device=net.device)
        check(lib.bnnp_glm_linear_samples(net.ops, net.n_ops, C.byref(pl), n, ptr(ws), ptr(f), ptr(dtheta), S,
                                          ptr(dst), ptr(sc), stream()), 'bnnp_glm_linear_samples')
        if lik != cabi.LIK_GAUSSIAN:
            check(lib.bnnp_lik_inv_link(lik, ptr(dst), S * n, K, ptr(dst), stream()), 'bnnp_lik_inv_link')
        if dst is not out:
            out[:, s0:s0 + n] = dst
    return out.reshape(S, N) if net.flat_output else out
