"""The CPU oracle's restatement of the LaplaceGGN path (post_process,
get_diagonal_ggn, linear_sampling_predictive) against fixtures from the unmodified reference."""
import pytest
import torch

from oracle import port
from tests import cases

TOL = 1e-4


def _setup(name):
    g = cases.load_golden(name)
    model = port.make_simlp(**cases.GGN_CASES[name])
    port.set_theta(model, torch.from_numpy(g['theta']))
    lh = cases.make_lh(g['lh'])
    X, y, Xte = (torch.from_numpy(g[k]) for k in ('X', 'y', 'Xte'))
    half = len(X) // 2
    P = len(g['theta'])
    pp, pm = cases.ggn_prior(g, P)
    P0 = torch.diag(pp) if torch.is_tensor(pp) else pp * torch.eye(P)
    mu0 = pm if torch.is_tensor(pm) else torch.full((P,), pm)
    return g, model, lh, [(X[:half], y[:half]), (X[half:], y[half:])], Xte, P0, mu0


@pytest.mark.parametrize('name', sorted(cases.GGN_CASES))
def test_post_process(name):
    g, model, lh, batches, Xte, P0, mu0 = _setup(name)
    post = port.post_process(model, lh, batches, P0, mu0)
    assert cases.rel_err(post['precision'], g['precision']) < 1e-5
    assert cases.rel_err(post['mu'], g['mu']) < TOL
    Sref = torch.from_numpy(g['Sigma_chol'])
    assert cases.rel_err(post['Sigma_chol'] @ post['Sigma_chol'].t(), Sref @ Sref.t()) < TOL
    Sd, Scd = port.diagonal_ggn(post['precision'])
    assert cases.rel_err(Sd, g['Sigma_d']) < 1e-5
    assert cases.rel_err(Scd, g['Sigma_chol_d']) < 1e-5


@pytest.mark.parametrize('name', sorted(cases.GGN_CASES))
def test_linear_sampling_predictive(name):
    g, model, lh, batches, Xte, P0, mu0 = _setup(name)
    eps = torch.from_numpy(g['eps'])
    theta = port.get_theta(model)
    mu, chol, chold = (torch.from_numpy(g[k]) for k in ('mu', 'Sigma_chol', 'Sigma_chol_d'))
    for tag, m, c, nl in (('full', mu, chol, False), ('map', theta, chol, False), ('diag', theta, chold, False),
                          ('vec', mu, torch.diag(chold).clone(), False), ('nolink', mu, chol, True)):
        out = port.linear_sampling_predictive(model, lh, Xte, m, c, eps, no_link=nl)
        assert out.shape == g['lin_' + tag].shape
        assert cases.rel_err(out, g['lin_' + tag]) < TOL, tag
    post = {'mu': theta, 'Sigma_chol': chol, 'cov_type': 'full'}
    out = port.bnn_predictive(model, lh, Xte, post, eps)
    assert cases.rel_err(out.reshape(g['nn_full'].shape), g['nn_full']) < TOL


@pytest.mark.parametrize('name', sorted(cases.GGN_CASES))
def test_refinement(name):
    """preds/refine.py restated, against the reference's outputs for the same draws."""
    g, model, lh, batches, Xte, P0, mu0 = _setup(name)
    X, y = torch.from_numpy(g['X']), torch.from_numpy(g['y'])
    mu, chol, chold, losses = port.laplace_refine(model, X, y, lh, 0.7, n_epochs=6, lr=1e-2)
    assert cases.rel_err(losses, g['lapref_losses']) < 1e-5
    assert cases.rel_err(mu, g['lapref_mu']) < 1e-5
    assert cases.rel_err(chol @ chol.t(), g['lapref_chol'] @ g['lapref_chol'].T) < TOL
    assert cases.rel_err(chold, g['lapref_chold']) < 1e-5
    post = {'precision': torch.from_numpy(g['precision']), 'Sigma_chol': torch.from_numpy(g['Sigma_chol'])}
    eps = torch.from_numpy(g['eps_vi'])
    mu, chol, losses = port.vi_refine(model, post, P0, X, y, lh, eps, lr=0.05)
    assert cases.rel_err(losses, g['vi_losses']) < TOL
    assert cases.rel_err(mu, g['vi_mu']) < TOL
    assert cases.rel_err(chol @ chol.t(), g['vi_chol'] @ g['vi_chol'].T) < TOL
    mu, chold, losses = port.vi_diag_refine(model, post, P0, X, y, lh, eps, lr=0.05)
    assert cases.rel_err(losses, g['vid_losses']) < TOL
    assert cases.rel_err(mu, g['vid_mu']) < TOL
    assert cases.rel_err(chold, g['vid_chol']) < TOL
