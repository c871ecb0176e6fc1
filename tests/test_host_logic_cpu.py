"""Host-side logic of the widened API that needs no GPU: prior expansion, the MAP training step of LaplaceGGN (torch
autograd, checked against the losses / weights the unmodified reference produced), get_diagonal_ggn, the ECE
bookkeeping -- and that everything which needs the CUDA library fails loudly on a CPU-only machine."""
import numpy as np
import pytest
import torch

from oracle import port
from tests import cases


def _pkg():
    import preds_b200
    return preds_b200


def test_expand_priors():
    from preds_b200.optimizers import expand_prior_prec, expand_prior_mu
    P = 5
    P0, S0 = expand_prior_prec(2.0, P, 'cpu')
    assert torch.equal(P0, 2.0 * torch.eye(P)) and torch.allclose(S0, 0.5 * torch.eye(P))
    P0, S0 = expand_prior_prec(torch.tensor(4.0), P, 'cpu')
    assert torch.equal(P0, 4.0 * torch.eye(P)) and torch.allclose(S0, 0.25 * torch.eye(P))
    d = torch.arange(1.0, P + 1)
    P0, S0 = expand_prior_prec(d, P, 'cpu')
    assert torch.equal(P0, torch.diag(d)) and torch.allclose(S0, torch.diag(1 / d))
    M = torch.diag(d) + 0.1
    P0, S0 = expand_prior_prec(M, P, 'cpu')
    assert P0 is M and torch.allclose(S0 @ M, torch.eye(P), atol=1e-5)
    with pytest.raises(ValueError):
        expand_prior_prec(torch.zeros(2, 2, 2), P, 'cpu')
    assert torch.equal(expand_prior_mu(0.5, P, 'cpu'), torch.full((P,), 0.5))
    v = torch.randn(P)
    assert expand_prior_mu(v, P, 'cpu') is v
    with pytest.raises(ValueError):
        expand_prior_mu(1, P, 'cpu')            # an int is not a float (reference quirk, preds/optimizers.py:29)


@pytest.mark.parametrize('name', sorted(cases.GGN_CASES))
def test_laplace_ggn_training_steps_match_reference(name):
    """LaplaceGGN.step is plain torch: on the CPU it must reproduce the reference's losses and weights."""
    pkg = _pkg()
    from preds_b200.models import SiMLP
    g = cases.load_golden(name)
    model = SiMLP(**cases.GGN_CASES[name])
    port.set_theta(model, torch.from_numpy(g['theta0']))
    lh = cases.make_lh(g['lh'], lib=pkg)
    X, y = torch.from_numpy(g['X']), torch.from_numpy(g['y'])
    pp, pm = cases.ggn_prior(g, len(g['theta']))
    opt = pkg.LaplaceGGN(model, lr=1e-2, prior_prec=pp, prior_mu=pm)
    losses = []
    for _ in range(int(g['n_steps'])):
        def closure():
            model.zero_grad()
            return lh.log_likelihood(y, model(X))
        losses.append(opt.step(closure))
    assert np.allclose(losses, g['losses'], rtol=1e-5)
    assert cases.rel_err(port.get_theta(model), g['theta']) < 1e-5
    # the posterior needs the CUDA library: no silent CPU path
    with pytest.raises(pkg._cabi.BnnpError):
        opt.post_process(model, lh, [(X, y)])


def test_get_diagonal_ggn_and_ece_bookkeeping():
    pkg = _pkg()
    from preds_b200.utils import _ece_from

    class Opt:
        state = {'precision': torch.tensor([[4.0, 1.0], [1.0, 16.0]])}
    Sd, Scd = pkg.get_diagonal_ggn(Opt())
    assert torch.equal(Sd, torch.diag(torch.tensor([0.25, 0.0625])))
    assert torch.equal(Scd, torch.diag(torch.tensor([0.5, 0.25])))
    # two bins: 3 samples with mean confidence 0.6 and accuracy 1/3, 1 sample with confidence 0.9 and accuracy 1
    out = torch.tensor([0, 0, 0, 3, 1.8, 1, 1, 0.9, 1], dtype=torch.float64)
    assert abs(_ece_from(out, 4, 2) - (abs(0.6 - 1 / 3) * 0.75 + abs(0.9 - 1.0) * 0.25)) < 1e-12


def test_new_entry_points_fail_loudly_without_a_gpu():
    pkg = _pkg()
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    from preds_b200.models import SiMLP
    from preds_b200.predictives import linear_sampling_predictive, linear_regression_predictive
    from preds_b200.refine import laplace_refine
    from preds_b200.utils import nll_cls
    model = SiMLP(4, 3, 1, 5)
    X, y = torch.randn(6, 4), torch.randint(3, (6,))
    theta = port.get_theta(model)
    lh = pkg.CategoricalLh()
    for call in (lambda: linear_sampling_predictive(X, model, lh, theta, torch.ones(len(theta)), mc_samples=2),
                 lambda: linear_regression_predictive(X, model, lh, theta, torch.ones(len(theta))),
                 lambda: laplace_refine(model, X, y, lh, 1.0, n_epochs=1),
                 lambda: nll_cls(torch.full((6, 3), 1 / 3), y, lh)):
        with pytest.raises(pkg._cabi.BnnpError):
            call()
