"""GLM refinement with the signatures of ``preds/refine.py:9-121`` (``laplace_refine``, ``vi_refine``,
``vi_diag_refine``): iterate on the linearised model  f(theta) = f* + J (theta - theta*)  with J fixed.

The reference materialises ``J [N,P,K]`` once and re-evaluates ``J theta``, ``J^T r`` and ``J^T Lambda J`` with
einsums every iteration.  Here the layer factors of one forward + K-column backward stay resident in the
workspace and every iteration calls the same kernels as the rest of the path on them: ``bnnp_glm_fmu`` (J v),
``bnnp_jtr_accum`` (J^T r), ``bnnp_ggn_full_accum`` / ``bnnp_ggn_diag_lambda`` (J^T Lambda J with the Lambda of that
iteration).  The small dense factorisations, the Adam update of the [P] mean and the scalar loss bookkeeping
(log-likelihood, KL) are torch on the device, as in ``Laplace.infer``.

Additive: ``eps=`` ([n_epochs, P]) injects the standard-normal draws of the VI loops for parity tests.
"""
import ctypes as C

import torch
from torch.distributions import MultivariateNormal, Normal, kl_divergence
from torch.nn.utils import parameters_to_vector
from torch.optim import Adam

from . import _cabi as cabi
from . import dense
from . import factor
from ._cabi import lib, check, ptr, stream
from ._engine import net_for


class _Linearised:
    """Jacobian factors of ``model`` at its current weights for one batch X, kept resident."""

    def __init__(self, model, X):
        self.net = net = net_for(model)
        self.theta_star = parameters_to_vector(model.parameters()).detach().to(torch.float32).contiguous()
        X = cabi.f32(X, net.device)
        net.prepare(X[:1])
        if not net.linear_only:
            raise cabi.BnnpError('GLM refinement is implemented for Linear-only networks')
        self.X2d, self.pl, self.ws, self.f = net.jacobian_factors(X)
        self.N, self.K, self.P = self.X2d.shape[0], net.K, net.P
        self.flat = net.flat_output

    def _shape(self, f):            # as the model returns it (preds/models.py:165)
        return f.reshape(self.N) if self.flat else f

    def f_at(self, theta):
        """f* + J (theta - theta*)  [N,K]  (= f_offset + J @ theta of the reference)."""
        delta = (theta - self.theta_star).contiguous()
        out = torch.empty(self.N, self.K, dtype=torch.float32, device=self.net.device)
        check(lib.bnnp_glm_fmu(self.net.ops, self.net.n_ops, C.byref(self.pl), self.N, ptr(self.X2d), ptr(self.ws),
                               ptr(self.f), ptr(delta), ptr(out), stream()), 'bnnp_glm_fmu')
        return out

    def jtr(self, rs):
        g = torch.zeros(self.P, dtype=torch.float32, device=self.net.device)
        rs = rs.reshape(self.N, self.K).contiguous()
        sc = self.net.scratch(lib.bnnp_jtr_scratch_floats(C.byref(self.pl), self.N))
        check(lib.bnnp_jtr_accum(self.net.ops, self.net.n_ops, C.byref(self.pl), self.N, ptr(self.ws), ptr(rs), ptr(g),
                                 ptr(sc), stream()), 'bnnp_jtr_accum')
        return g

    def ggn(self, Lam):
        H = torch.zeros(self.P, self.P, dtype=torch.float32, device=self.net.device)
        Lam = Lam.reshape(self.N, self.K, self.K).contiguous()
        sc = self.net.scratch(lib.bnnp_ggn_full_scratch_floats(C.byref(self.pl), self.N, 0))
        check(lib.bnnp_ggn_full_accum(self.net.ops, self.net.n_ops, C.byref(self.pl), self.N, ptr(self.ws), ptr(Lam),
                                      ptr(H), ptr(sc), 0, stream()), 'bnnp_ggn_full_accum')
        check(lib.bnnp_symmetrize_lower(ptr(H), self.P, stream()), 'bnnp_symmetrize_lower')
        return H

    def ggn_diag(self, Lam):
        d = torch.zeros(self.P, dtype=torch.float32, device=self.net.device)
        Lam = Lam.reshape(self.N, self.K, self.K).contiguous()
        sc = self.net.scratch(lib.bnnp_jtr_scratch_floats(C.byref(self.pl), self.N))
        check(lib.bnnp_ggn_diag_lambda(self.net.ops, self.net.n_ops, C.byref(self.pl), self.N, ptr(self.ws), ptr(Lam),
                                       ptr(d), ptr(sc), stream()), 'bnnp_ggn_diag_lambda')
        return d


def _chol_inverse_chol(prec):
    """(chol(prec), chol(prec^-1)) -- preds/refine.py:36-38,72,75."""
    pchol = factor.cholesky(prec)
    Sigma = factor.inverse_from_chol(pchol) if factor.use_tc(pchol.shape[0]) else dense.spd_inverse_from_chol(pchol)
    return pchol, factor.cholesky(Sigma)


def laplace_refine(model, X, y, likelihood, prior_prec, n_epochs=1000, lr=1e-3):
    """Laplace in the linearised model (preds/refine.py:9-40): Adam on the GLM's log joint, then the GGN posterior
    at the optimum.  Returns (mu, Sigma_chol, Sigma_chol_diag, losses)."""
    lin = _Linearised(model, X)
    y = y.to(lin.net.device)
    mu = lin.theta_star.clone().requires_grad_(True)
    opt = Adam([mu], lr=lr)
    losses = []
    for _ in range(n_epochs):
        opt.zero_grad()
        with torch.no_grad():
            f = lin.f_at(mu)
            # d(-log p(y|f))/df = -residual for these likelihoods; chain through J with bnnp_jtr_accum
            mu.grad = prior_prec * mu - lin.jtr(likelihood.residual(y, f))
            loss = -likelihood.log_likelihood(y, lin._shape(f)) + 0.5 * (mu @ mu) * prior_prec
        losses.append(loss.item())
        opt.step()
    mu = mu.detach()
    H = lin.ggn(likelihood._hessian_nkk(lin.f_at(mu)))
    H.diagonal().add_(prior_prec)
    _, Sigma_chol = _chol_inverse_chol(H)
    Sigma_chold = torch.diag(torch.sqrt(1 / torch.diag(H)))
    return mu, Sigma_chol, Sigma_chold, losses


def _draw(eps, i, P, device):
    if eps is None:
        return torch.randn(P, device=device)
    return cabi.f32(eps[i], device)


def vi_refine(model, opt, X, y, likelihood, n_epochs=250, lr=1e-2, eps=None):
    """Natural-gradient VI with a full Gaussian in the linearised model (preds/refine.py:42-81), started from the
    ``LaplaceGGN`` posterior in ``opt.state``.  Returns (mu, Sigma_chol, losses)."""
    beta = lr
    lin = _Linearised(model, X)
    y = y.to(lin.net.device)
    mu = lin.theta_star.clone()
    P = lin.P
    Sigma_chol = opt.state['Sigma_chol'].clone()
    prior_prec = opt.state['prior_prec']
    prec = opt.state['precision'].clone()
    sigma_chol_prior = torch.diag(torch.sqrt(1 / torch.diag(prior_prec)))
    p = MultivariateNormal(torch.zeros_like(mu), scale_tril=sigma_chol_prior)
    losses = []
    for i in range(n_epochs):
        theta_s = mu + Sigma_chol @ _draw(eps, i, P, mu.device)
        f_t = lin.f_at(theta_s)
        H = lin.ggn(likelihood._hessian_nkk(f_t))                       # Hessian of E[log p]
        g = lin.jtr(likelihood.residual(y, f_t)) - prior_prec @ mu      # gradient of the ELBO
        prec = (1 - beta) * prec + beta * (H + prior_prec)
        pchol, Sigma_chol = _chol_inverse_chol(prec)
        mu = mu + beta * torch.cholesky_solve(g.reshape(-1, 1), pchol, upper=False).squeeze()
        q = MultivariateNormal(loc=mu, scale_tril=Sigma_chol)
        loss = -likelihood.log_likelihood(y, lin._shape(f_t)) + kl_divergence(q, p)
        losses.append(loss.item())
    return mu, Sigma_chol, losses


def vi_diag_refine(model, opt, X, y, likelihood, n_epochs=250, lr=1e-3, eps=None):
    """Mean-field variant (preds/refine.py:84-121).  Returns (mu, diag-matrix Sigma_chol, losses)."""
    beta = lr
    lin = _Linearised(model, X)
    y = y.to(lin.net.device)
    mu = lin.theta_star.clone()
    P = lin.P
    prec = torch.diag(opt.state['precision']).clone()
    prior_prec = torch.diag(opt.state['prior_prec'])
    sigma_chol = torch.sqrt(1 / prec)
    m_prior = Normal(loc=torch.zeros_like(mu), scale=1 / torch.sqrt(prior_prec))
    losses = []
    for i in range(n_epochs):
        theta_s = mu + sigma_chol * _draw(eps, i, P, mu.device)
        f_t = lin.f_at(theta_s)
        H = lin.ggn_diag(likelihood._hessian_nkk(f_t))
        g = lin.jtr(likelihood.residual(y, f_t)) - prior_prec * mu
        prec = (1 - beta) * prec + beta * (H + prior_prec)
        mu = mu + beta * g / prec
        sigma_chol = torch.sqrt(1 / prec)
        m_post = Normal(loc=mu, scale=sigma_chol)
        loss = -likelihood.log_likelihood(y, lin._shape(f_t)) + kl_divergence(m_post, m_prior).sum()
        losses.append(loss.item())
    return mu, torch.diag(sigma_chol), losses
