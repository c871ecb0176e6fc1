"""Minimal stand-in for ``gpytorch.distributions.MultivariateNormal`` as ``preds/gps.py:447`` / ``mc_preds`` use it:
positional (mean, covariance), ``.mean``, a settable ``.loc`` (preds/laplace.py:205) and ``.sample(torch.Size([S]))``
drawing ``loc + chol(cov) eps`` (gpytorch's root decomposition of a small dense covariance is its Cholesky factor)."""
import torch
from torch.distributions import MultivariateNormal as _MVN


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        if hasattr(covariance_matrix, 'evaluate'):
            covariance_matrix = covariance_matrix.evaluate()
        self.loc = mean
        self.covariance_matrix = covariance_matrix

    @property
    def mean(self):
        return self.loc

    def sample(self, sample_shape=torch.Size()):
        return _MVN(self.loc, self.covariance_matrix).sample(sample_shape)


def __getattr__(name):
    return type(name, (), {})
