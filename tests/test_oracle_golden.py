"""The CPU oracle (oracle/port.py) against fixtures produced by the unmodified reference.

This is what pins the oracle: every quantity of the hot path -- Jacobians, likelihood Hessian,
full GGN, Sigma_chol, Kronecker factors, GLM and BNN predictives -- for every golden case."""
import numpy as np
import pytest
import torch

from oracle import port
from tests import cases

TOL = 1e-4   # relative, north-star bound; the port typically agrees to ~1e-6


def _setup(name):
    g = cases.load_golden(name)
    model = cases.make_port_model(name)
    port.set_theta(model, torch.from_numpy(g['theta']))
    lh = cases.make_lh(g['lh'])
    batches, Xte = cases.case_data(g)
    return g, model, lh, batches, Xte, float(g['prior'])


@pytest.mark.parametrize('name', cases.FULL_CASES)
def test_jacobians_and_hessian(name):
    g, model, lh, batches, Xte, prior = _setup(name)
    Js, f = port.jacobians(model, batches[0][0])
    assert Js.shape == g['Js0'].shape
    assert cases.rel_err(Js, g['Js0']) < 1e-5
    assert cases.rel_err(f, g['f0']) < 1e-5
    _, Hess, rs, _ = port.ggn_bundle(model, lh, *batches[0])
    assert cases.rel_err(Hess, g['Hess0']) < 1e-5
    assert cases.rel_err(rs, g['rs0']) < 1e-5


@pytest.mark.parametrize('name', cases.FULL_CASES)
def test_full_ggn_and_posterior(name):
    g, model, lh, batches, Xte, prior = _setup(name)
    post = port.infer(model, prior, lh, batches, 'full')
    assert cases.rel_err(post['precision'], g['H_full']) < 1e-5
    # compare the covariance (chol of an ill-conditioned inverse is compared via Sigma)
    Sig = post['Sigma_chol'] @ post['Sigma_chol'].t()
    Sig_ref = torch.from_numpy(g['Sigma_chol_full'])
    assert cases.rel_err(Sig, Sig_ref @ Sig_ref.t()) < TOL
    fmu, fvar = port.glm_moments(model, Xte, post)
    assert cases.rel_err(fmu, g['glm_fmu_full']) < 1e-5
    assert cases.rel_err(fvar, g['glm_fvar_full']) < TOL
    if lh.name != 'gaussian':
        out = port.glm_predictive(model, lh, Xte, post, torch.from_numpy(g['eps_glm']))
        assert cases.rel_err(out, g['glm_full']) < TOL
    out = port.bnn_predictive(model, lh, Xte, post, torch.from_numpy(g['eps_bnn']))
    assert cases.rel_err(out, g['bnn_full']) < 5e-4


@pytest.mark.parametrize('name', [c for c in cases.FULL_CASES if c != 'mlp_bern'])
def test_diag(name):
    g, model, lh, batches, Xte, prior = _setup(name)
    post = port.infer(model, prior, lh, batches, 'diag')
    assert cases.rel_err(post['Sigma_chol'], g['Sigma_chol_diag']) < 1e-5
    # identity that pins DiagGGNExact: diag path == diag(full GGN) up to the 1e-7 clamp
    H = torch.from_numpy(g['H_full'])
    assert cases.rel_err(post['precision'], H.diagonal()) < 1e-4
    fmu, fvar = port.glm_moments(model, Xte, post)
    assert cases.rel_err(fvar, g['glm_fvar_diag']) < TOL
    if lh.name != 'gaussian':
        out = port.glm_predictive(model, lh, Xte, post, torch.from_numpy(g['eps_glm']))
        assert cases.rel_err(out, g['glm_diag']) < TOL
    out = port.bnn_predictive(model, lh, Xte, post, torch.from_numpy(g['eps_bnn']))
    assert cases.rel_err(out, g['bnn_diag']) < TOL


@pytest.mark.parametrize('name', [c for c in cases.FULL_CASES if c != 'mlp_bern'])
def test_kron(name):
    g, model, lh, batches, Xte, prior = _setup(name)
    post = port.infer(model, prior, lh, batches, 'kron')
    kron = post['kron']
    for pi, qh in enumerate(kron.qhs):
        for fi, m in enumerate(qh):
            assert cases.rel_err(m, g['kron_q_%d_%d' % (pi, fi)]) < 1e-5
    fmu, fvar = port.glm_moments(model, Xte, post)
    assert cases.rel_err(fvar, g['glm_fvar_kron']) < TOL
    if lh.name != 'gaussian':
        out = port.glm_predictive(model, lh, Xte, post, torch.from_numpy(g['eps_glm']))
        assert cases.rel_err(out, g['glm_kron']) < TOL
    eps = cases.kron_eps_from(g)
    out = port.bnn_predictive(model, lh, Xte, post, eps)
    assert cases.rel_err(out, g['bnn_kron']) < TOL
    kron.dampen = True
    _, fvar_d = port.glm_moments(model, Xte, post)
    assert cases.rel_err(fvar_d, g['glm_fvar_kron_dampen']) < TOL
    out = port.bnn_predictive(model, lh, Xte, post, eps)
    assert cases.rel_err(out, g['bnn_kron_dampen']) < TOL


def test_cifar10net_probe():
    """The reference tests' own CIFAR10Net fixture (seed 71 / seed 15), probed."""
    g = cases.load_golden('cifar10net_probe')
    torch.manual_seed(71)
    model = port.make_cifar10net(in_channels=2, n_out=3)
    theta = port.get_theta(model)
    idx = torch.from_numpy(g['theta_idx'])
    assert np.array_equal(theta[idx].numpy(), g['theta_at_idx'])   # same init stream
    torch.manual_seed(15)
    X = torch.randn(8, 2, 28, 28)
    y = torch.from_numpy(g['y'])
    gv = torch.Generator().manual_seed(4243)
    v = torch.randn(len(theta), generator=gv)
    Js, f = port.jacobians(model, X)
    assert cases.rel_err(f, g['f']) < 1e-5
    assert cases.rel_err(Js[:, idx, :], g['Js_at_idx']) < 1e-5
    assert cases.rel_err(torch.einsum('npk,p->nk', Js, v), g['Js_dot_v']) < 1e-4
    del Js
    lh = port.CategoricalLh()
    batches = [(X[:5], y[:5]), (X[5:], y[5:])]
    post = port.infer(model, 1.0, lh, batches, 'diag')
    assert cases.rel_err(post['precision'][idx], g['diag_prec_at_idx']) < 1e-5
    post = port.infer(model, 1.0, lh, batches, 'kron')
    for pi, qh in enumerate(post['kron'].qhs):
        for fi, m in enumerate(qh):
            pv = torch.randn(m.shape[0], generator=torch.Generator().manual_seed(5000 + 10 * pi + fi))
            assert cases.rel_err(m @ pv, g['kron_q_%d_%d_mv' % (pi, fi)]) < 1e-4
    Xte = torch.randn(3, 2, 28, 28, generator=torch.Generator().manual_seed(16))
    fmu, fvar = port.glm_moments(model, Xte, post)
    assert cases.rel_err(fmu, g['glm_fmu_kron']) < 1e-5
    assert cases.rel_err(fvar, g['glm_fvar_kron']) < TOL


def test_cifar100net_probe():
    """CIFAR100Net (preds/models.py:182-236, K = 100: stride-2 and 1x1 convs, AvgPool2d(6)) on two 32x32 examples,
    probed: the oracle's restatement of the model and of the BackPACK quantities at K = 100."""
    g = cases.load_golden('cifar100net_probe')
    torch.manual_seed(71)
    model = port.make_cifar100net(3, 100)
    theta = port.get_theta(model)
    idx = torch.from_numpy(g['theta_idx'])
    assert np.array_equal(theta[idx].numpy(), g['theta_at_idx'])   # same init stream
    X = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(15))
    y = torch.from_numpy(g['y'])
    v = torch.randn(len(theta), generator=torch.Generator().manual_seed(4243))
    Js, f = port.jacobians(model, X)
    assert cases.rel_err(f, g['f']) < 1e-5
    assert cases.rel_err(Js[:, idx, :], g['Js_at_idx']) < 1e-5
    assert cases.rel_err(torch.einsum('npk,p->nk', Js, v), g['Js_dot_v']) < 1e-4
    del Js
    lh = port.CategoricalLh()
    post = port.infer(model, 1.0, lh, [(X, y)], 'diag')
    assert cases.rel_err(post['precision'][idx], g['diag_prec_at_idx']) < 1e-5
    post = port.infer(model, 1.0, lh, [(X, y)], 'kron')
    for pi, qh in enumerate(post['kron'].qhs):
        for fi, m in enumerate(qh):
            pv = torch.randn(m.shape[0], generator=torch.Generator().manual_seed(5000 + 10 * pi + fi))
            assert cases.rel_err(m @ pv, g['kron_q_%d_%d_mv' % (pi, fi)]) < 1e-4
    Xte = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(16))
    fmu, fvar = port.glm_moments(model, Xte, post)
    assert cases.rel_err(fmu, g['glm_fmu_kron']) < 1e-5
    assert cases.rel_err(fvar, g['glm_fvar_kron']) < TOL
