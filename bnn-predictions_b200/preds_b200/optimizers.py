"""``GGN`` bundle and the ``LaplaceGGN`` optimiser with the interface of ``preds/optimizers.py``.

``LaplaceGGN.step`` is ordinary MAP training (Adam on the log joint) and stays in torch autograd; the
posterior work of ``post_process`` -- ``J^T Lambda J``, ``J^T r`` and the factorisations -- runs through the
CUDA library on the Jacobian factors.
"""
import torch
from torch.nn.utils import parameters_to_vector
from torch.optim import Adam

from . import dense
from .gradients import Jacobians


def GGN(model, likelihood, data, target=None, ret_f=False):
    """preds/optimizers.py:8-25."""
    Js, f = Jacobians(model, data)
    if target is not None:
        rs = likelihood.residual(target.to(f.device), f)
    Hess = likelihood.Hessian(f)
    m, p = Js.shape[:2]
    if Js.dim() == 2:
        Hess = Hess.reshape(m, 1, 1)
        if target is not None:
            rs = rs.reshape(m, 1)
        Js = Js.reshape(m, p, 1)
    if target is not None:
        return (Js, Hess, rs, f) if ret_f else (Js, Hess, rs)
    return Js, Hess, f


def expand_prior_mu(prior_mu, P, device):
    """preds/optimizers.py:28-34."""
    if type(prior_mu) is float:
        return torch.full((P,), prior_mu, device=device)
    if type(prior_mu) is torch.Tensor:
        return prior_mu
    raise ValueError('Invalid shape for prior mean')


def expand_prior_prec(prior_prec, P, device):
    """preds/optimizers.py:37-47: (prior precision matrix, prior covariance matrix)."""
    if type(prior_prec) is float or prior_prec.ndim == 0:
        d = torch.full((P,), float(prior_prec), device=device)
        return torch.diag(d), torch.diag(1 / d)
    if prior_prec.ndim == 1:
        return torch.diag(prior_prec), torch.diag(1 / prior_prec)
    if prior_prec.ndim == 2:
        return prior_prec, torch.inverse(prior_prec)
    raise ValueError('Invalid shape for prior precision')


class LaplaceGGN(Adam):
    """preds/optimizers.py:50-104.  ``state`` keys as in the reference: prior_prec, Sigma_0, prior_mu, mu,
    precision, Sigma, Sigma_chol."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), prior_prec=1.0, prior_mu=0.0, eps=1e-8,
                 amsgrad=False, **kwargs):
        if 'beta1' in kwargs and 'beta2' in kwargs:
            betas = (kwargs['beta1'], kwargs['beta2'])
        # the prior penalty is part of the loss in step(), so no weight decay here
        super().__init__(model.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=amsgrad)
        theta = parameters_to_vector(model.parameters())
        P, device = len(theta), theta.device
        self.defaults['device'] = device
        P_0, S_0 = expand_prior_prec(prior_prec, P, device)
        self.state['prior_prec'] = P_0
        self.state['Sigma_0'] = S_0
        self.state['prior_mu'] = expand_prior_mu(prior_mu, P, device)
        self.state['mu'] = None
        self.state['precision'] = None
        self.state['Sigma_chol'] = None

    def step(self, closure):
        """One Adam step on  -log p(y|f) + 1/2 theta^T P0 theta  (:70-79); returns the loss."""
        log_lik = closure()
        params = parameters_to_vector(self.param_groups[0]['params'])
        loss = -log_lik + 0.5 * params @ self.state['prior_prec'] @ params
        loss.backward()
        super().step()
        return loss.item()

    def post_process(self, model, likelihood, train_loader):
        """Posterior of the linearised model around the trained weights (:81-104):
        precision = sum_n J^T Lambda J + P0, its inverse and Cholesky factors, and the posterior mean
        mu = precision^-1 (sum_n J^T r + JLJ theta* + P0 mu0)."""
        from .laplace import Laplace
        device = self.defaults['device']
        theta_star = parameters_to_vector(self.param_groups[0]['params']).detach().to(device)
        prior_prec, prior_mu = self.state['prior_prec'], self.state['prior_mu']
        P = len(theta_star)
        was_training = model.training
        model.eval()            # the factors are taken in inference mode (no-op for the reference's MLPs)
        JLJ = torch.zeros(P, P, device=device)
        G = torch.zeros(P, device=device)
        Laplace(model, 1.0, likelihood)._accumulate_full(iter(train_loader), JLJ, grad=G)
        model.train(was_training)
        b = G + JLJ @ theta_star + prior_prec @ prior_mu
        precision = JLJ.add_(prior_prec)
        self.state['precision'] = precision
        Chol = torch.linalg.cholesky(precision)
        self.state['mu'] = torch.cholesky_solve(b.reshape(-1, 1), Chol, upper=False).flatten()
        Sigma = dense.spd_inverse_from_chol(Chol)      # potri below dense.MIN_P, blocked inverse above
        del Chol
        self.state['Sigma'] = Sigma
        self.state['Sigma_chol'] = torch.linalg.cholesky(Sigma)
        return self


def get_diagonal_ggn(optimizer):
    """preds/optimizers.py:107-112: (diag covariance matrix, its Cholesky factor) from the precision diagonal."""
    Sigma_diag = 1 / torch.diag(optimizer.state['precision'])
    return torch.diag(Sigma_diag), torch.diag(torch.sqrt(Sigma_diag))
