"""Likelihoods with the interface of ``preds/likelihoods.py`` (Hessian, residual, inv_link,
log_likelihood, nn_loss, sigma_2_dispersion); the tensor work runs in ``csrc/lik.cu``."""
import torch
from torch.distributions import Bernoulli, Categorical, Normal

from . import _cabi as cabi
from ._cabi import lib, check, ptr, stream

EPS = 1e-7


def get_Lams_Vys(lh, Hess):
    """preds/likelihoods.py:7-12."""
    Lams = Hess * lh.sigma_2_dispersion
    return Lams, Lams * lh.sigma_2_dispersion


def _as2d(f):
    cabi.require_device(f)
    f = f.detach().to(torch.float32).contiguous()
    return f.reshape(-1, f.shape[-1]) if f.dim() > 1 else f.reshape(-1, 1)


class Likelihood:
    code = None

    @property
    def sigma_2_dispersion(self):
        return 1.0

    def log_likelihood(self, y, f):
        raise NotImplementedError

    def _sigma2(self):
        return float(self.sigma_2_dispersion)

    def residual(self, y, f):
        f2 = _as2d(f)
        N, K = f2.shape
        if self.code == cabi.LIK_CATEGORICAL:
            yy = y.to(device=f2.device).long().contiguous()
        else:
            yy = y.to(device=f2.device, dtype=torch.float32).reshape(N, K).contiguous()
        rs = torch.empty_like(f2)
        check(lib.bnnp_lik_residual(self.code, ptr(f2), ptr(yy), N, K, self._sigma2(), ptr(rs), stream()),
              'bnnp_lik_residual')
        return rs.reshape(f.shape)

    def _hessian_nkk(self, f):
        f2 = _as2d(f)
        N, K = f2.shape
        H = torch.empty(N, K, K, dtype=torch.float32, device=f2.device)
        check(lib.bnnp_lik_hessian(self.code, ptr(f2), N, K, self._sigma2(), ptr(H), stream()),
              'bnnp_lik_hessian')
        return H

    def Hessian(self, f):
        return self._hessian_nkk(f)

    def inv_link(self, f):
        f2 = _as2d(f)
        out = torch.empty_like(f2)
        check(lib.bnnp_lik_inv_link(self.code, ptr(f2), f2.shape[0], f2.shape[1], ptr(out), stream()),
              'bnnp_lik_inv_link')
        return out.reshape(f.shape)

    def nn_loss(self):
        raise NotImplementedError


class GaussianLh(Likelihood):
    """preds/likelihoods.py:34-58."""
    code = cabi.LIK_GAUSSIAN

    def __init__(self, sigma_noise=1):
        self.sigma_noise = sigma_noise

    @property
    def sigma_2_dispersion(self):
        return self.sigma_noise ** 2

    def log_likelihood(self, y, f):
        return torch.sum(Normal(f, self.sigma_noise).log_prob(y))

    def Hessian(self, f):
        assert f.size(1) == 1
        return self._hessian_nkk(f)

    def nn_loss(self):
        return torch.nn.MSELoss(reduction='sum'), 1 / (2 * self.sigma_2_dispersion)


class BernoulliLh(Likelihood):
    """preds/likelihoods.py:61-75."""
    code = cabi.LIK_BERNOULLI

    def log_likelihood(self, y, f):
        return torch.sum(Bernoulli(logits=f).log_prob(y))

    def Hessian(self, f):
        return self._hessian_nkk(f).reshape(f.shape)      # p(1-p), shape of f (:67-69)

    def nn_loss(self):
        raise ValueError('No extendable nn loss for backpack in Bernoulli case')


class CategoricalLh(Likelihood):
    """preds/likelihoods.py:78-99."""
    code = cabi.LIK_CATEGORICAL

    def log_likelihood(self, y, f):
        return torch.sum(Categorical(logits=f).log_prob(y))

    def nn_loss(self):
        return torch.nn.CrossEntropyLoss(reduction='sum'), 1
