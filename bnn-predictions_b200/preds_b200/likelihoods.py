"""Likelihoods behind the names of ``preds/likelihoods.py`` (``Hessian``, ``residual``, ``inv_link``,
``log_likelihood``, ``nn_loss``, ``sigma_2_dispersion``).

All tensor arithmetic of the path -- the K x K Hessian of the negative log-likelihood, the residual, the inverse
link -- runs in ``csrc/lik.cu`` on the device; ``log_likelihood`` (used by MAP training and the refinement loops for
loss bookkeeping) is the sum of torch-distribution log-probabilities, as in the reference.
"""
import torch
from torch import distributions as D

from . import _cabi as cabi
from ._cabi import lib, check, ptr, stream

EPS = 1e-7          # clamp of the class probabilities inside the Hessian (applied in lik.cu; preds/likelihoods.py:4)


def get_Lams_Vys(lh, Hess):
    """(Lambda * s2, Lambda * s2^2): Jacobian of the inverse link and observation variance of the linearised model
    (preds/likelihoods.py:7-12); s2 is 1 except for the Gaussian."""
    scaled = Hess * lh.sigma_2_dispersion
    return scaled, scaled * lh.sigma_2_dispersion


def _rows(f, scalar=False):
    """[rows, K] float32 contiguous device view of network outputs ([N] counts as K = 1).  ``scalar``: the likelihood
    acts elementwise on a single output (Bernoulli / Gaussian: the reference's sigmoid, p(1-p), identity work on any
    shape, e.g. the [S, N] samples of a model that flattens its output), so every element is its own row."""
    cabi.require_device(f)
    f = f.detach().to(torch.float32).contiguous()
    if scalar:
        return f.reshape(-1, 1)
    return f.reshape(-1, f.shape[-1]) if f.dim() > 1 else f.reshape(-1, 1)


class Likelihood:
    """Common machinery: subclasses set ``code`` (the library's likelihood id) and ``_distribution``."""
    code = None
    sigma_2_dispersion = 1.0
    _scalar = False         # True: one output, elementwise on any shape

    def _distribution(self, f):
        raise NotImplementedError

    def log_likelihood(self, y, f):
        return self._distribution(f).log_prob(y).sum()

    def _hessian_nkk(self, f):
        f2 = _rows(f, self._scalar)
        n, k = f2.shape
        out = torch.empty(n, k, k, dtype=torch.float32, device=f2.device)
        check(lib.bnnp_lik_hessian(self.code, ptr(f2), n, k, float(self.sigma_2_dispersion), ptr(out), stream()),
              'bnnp_lik_hessian')
        return out

    def Hessian(self, f):
        return self._hessian_nkk(f)

    def residual(self, y, f):
        f2 = _rows(f, self._scalar)
        n, k = f2.shape
        if self.code == cabi.LIK_CATEGORICAL:           # integer labels index the softmax (:84-88)
            target = y.to(device=f2.device).long().contiguous()
        else:
            target = y.to(device=f2.device, dtype=torch.float32).reshape(n, k).contiguous()
        out = torch.empty_like(f2)
        check(lib.bnnp_lik_residual(self.code, ptr(f2), ptr(target), n, k, float(self.sigma_2_dispersion), ptr(out),
                                    stream()), 'bnnp_lik_residual')
        return out.reshape(f.shape)

    def inv_link(self, f):
        f2 = _rows(f, self._scalar)
        out = torch.empty_like(f2)
        check(lib.bnnp_lik_inv_link(self.code, ptr(f2), f2.shape[0], f2.shape[1], ptr(out), stream()),
              'bnnp_lik_inv_link')
        return out.reshape(f.shape)

    def nn_loss(self):
        raise NotImplementedError


class CategoricalLh(Likelihood):
    """Softmax classification (preds/likelihoods.py:78-99): Hessian diag(p) - p p^T on clamped p."""
    code = cabi.LIK_CATEGORICAL

    def _distribution(self, f):
        return D.Categorical(logits=f)

    def nn_loss(self):
        return torch.nn.CrossEntropyLoss(reduction='sum'), 1


class BernoulliLh(Likelihood):
    """Binary classification on one logit (preds/likelihoods.py:61-75): Hessian p(1-p), returned in the shape of f."""
    code = cabi.LIK_BERNOULLI
    _scalar = True

    def _distribution(self, f):
        return D.Bernoulli(logits=f)

    def Hessian(self, f):
        return self._hessian_nkk(f).reshape(f.shape)

    def nn_loss(self):
        raise ValueError('No extendable nn loss for backpack in Bernoulli case')


class GaussianLh(Likelihood):
    """Regression with fixed noise (preds/likelihoods.py:34-58): Hessian 1/sigma^2, residual (y - f)/sigma^2."""
    code = cabi.LIK_GAUSSIAN
    _scalar = True

    def __init__(self, sigma_noise=1):
        self.sigma_noise = sigma_noise

    @property
    def sigma_2_dispersion(self):
        return self.sigma_noise ** 2

    def _distribution(self, f):
        return D.Normal(f, self.sigma_noise)

    def Hessian(self, f):
        assert f.size(1) == 1
        return self._hessian_nkk(f)

    def nn_loss(self):
        return torch.nn.MSELoss(reduction='sum'), 1 / (2 * self.sigma_2_dispersion)
