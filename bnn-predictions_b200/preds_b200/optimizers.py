"""``GGN`` bundle and the ``LaplaceGGN`` optimiser behind the names of ``preds/optimizers.py``.

Only the MAP training step is torch autograd (Adam on the log joint, as in the reference).  Everything
``post_process`` needs -- the GGN ``sum_n J^T Lambda J``, the gradient term ``sum_n J^T r`` and the dense
factorisations -- comes from the CUDA library working on the Jacobian factors of one forward + K-column
backward per super-batch; a Jacobian is never formed.
"""
import torch
from torch.nn.utils import parameters_to_vector
from torch.optim import Adam

from . import dense
from . import factor
from .gradients import Jacobians


def GGN(model, likelihood, data, target=None, ret_f=False):
    """(Js [m,p,k], Hess [m,k,k], rs [m,k] when a target is given, f) -- preds/optimizers.py:8-25.
    Single-output models come back from ``Jacobians`` as [m,p]; they are lifted to k = 1 here."""
    Js, f = Jacobians(model, data)
    single = Js.dim() == 2
    m = Js.shape[0]
    Hess = likelihood.Hessian(f)
    out = [Js.unsqueeze(-1) if single else Js, Hess.reshape(m, 1, 1) if single else Hess]
    if target is None:
        return out[0], out[1], f
    rs = likelihood.residual(target.to(f.device), f)
    out.append(rs.reshape(m, 1) if single else rs)
    if ret_f:
        out.append(f)
    return tuple(out)


def expand_prior_mu(prior_mu, P, device):
    """Prior mean as a [P] vector (preds/optimizers.py:28-34): a float is broadcast, a tensor is taken as is."""
    if isinstance(prior_mu, torch.Tensor):
        return prior_mu
    if type(prior_mu) is not float:
        raise ValueError('Invalid shape for prior mean')
    return torch.full((P,), prior_mu, device=device)


def expand_prior_prec(prior_prec, P, device):
    """(precision matrix, covariance matrix) of the prior from a scalar, a [P] diagonal or a [P,P] matrix
    (preds/optimizers.py:37-47)."""
    if type(prior_prec) is float:
        prior_prec = torch.tensor(prior_prec, device=device)
    nd = prior_prec.ndim
    if nd > 2:
        raise ValueError('Invalid shape for prior precision')
    if nd == 2:
        return prior_prec, torch.inverse(prior_prec)
    diag = prior_prec.to(device).expand(P) if nd == 0 else prior_prec
    return torch.diag(diag), torch.diag(diag.reciprocal())


class LaplaceGGN(Adam):
    """Adam on the log joint during training, Laplace-GGN posterior afterwards (preds/optimizers.py:50-104).
    ``state`` carries the reference's keys: prior_prec, Sigma_0, prior_mu, mu, precision, Sigma, Sigma_chol."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), prior_prec=1.0, prior_mu=0.0, eps=1e-8,
                 amsgrad=False, **kwargs):
        if {'beta1', 'beta2'} <= set(kwargs):
            betas = (kwargs['beta1'], kwargs['beta2'])
        # weight decay stays off: the Gaussian prior enters the loss of step() explicitly
        super().__init__(model.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=amsgrad)
        flat = parameters_to_vector(model.parameters())
        self.defaults['device'] = flat.device
        P_0, S_0 = expand_prior_prec(prior_prec, flat.numel(), flat.device)
        self.state.update(prior_prec=P_0, Sigma_0=S_0, prior_mu=expand_prior_mu(prior_mu, flat.numel(), flat.device),
                          mu=None, precision=None, Sigma_chol=None)

    def _theta(self):
        return parameters_to_vector(self.param_groups[0]['params'])

    def step(self, closure):
        """One Adam step on the negative log joint  -log p(y|f(theta)) + theta^T P0 theta / 2  (:70-79); returns it."""
        theta = self._theta()
        penalty = 0.5 * torch.dot(theta, self.state['prior_prec'] @ theta)
        objective = penalty - closure()
        objective.backward()
        super().step()
        return objective.item()

    def post_process(self, model, likelihood, train_loader):
        """Gaussian posterior of the model linearised at the trained weights (:81-104):
        precision = sum_n J^T Lambda J + P0,  mu = precision^-1 (sum_n J^T r + JLJ theta* + P0 mu0),
        plus Sigma = precision^-1 and the lower Cholesky factor of Sigma."""
        from .laplace import Laplace
        device = self.defaults['device']
        theta_star = self._theta().detach().to(device)
        P0, mu0 = self.state['prior_prec'], self.state['prior_mu']
        n_par = theta_star.numel()
        curvature = torch.zeros(n_par, n_par, device=device)          # JLJ
        grad_term = torch.zeros(n_par, device=device)                 # sum_n J^T r
        was_training = model.training
        model.eval()        # factors in inference mode (no-op for the reference's MLPs)
        Laplace(model, 1.0, likelihood)._accumulate_full(iter(train_loader), curvature, grad=grad_term)
        model.train(was_training)
        rhs = grad_term + curvature @ theta_star + P0 @ mu0
        precision = curvature.add_(P0)
        pchol = factor.cholesky(precision)              # library below factor.MIN_P, own tensor-core chain above
        self.state['precision'] = precision
        self.state['mu'] = torch.cholesky_solve(rhs.unsqueeze(1), pchol, upper=False).squeeze(1)
        Sigma = factor.inverse_from_chol(pchol) if factor.use_tc(pchol.shape[0]) else dense.spd_inverse_from_chol(pchol)
        del pchol
        self.state['Sigma'] = Sigma
        self.state['Sigma_chol'] = factor.cholesky(Sigma)
        return self


def get_diagonal_ggn(optimizer):
    """Mean-field reading of the posterior (preds/optimizers.py:107-112): (diag(1/d), diag(1/sqrt(d))) with d the
    diagonal of the precision."""
    variances = optimizer.state['precision'].diagonal().reciprocal()
    return torch.diag(variances), torch.diag(variances.sqrt())
