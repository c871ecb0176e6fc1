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


def test_function_space_host_logic_and_loud_failure():
    """Prior-variance groups of gp_quantities (preds/gps.py:700-735: one kernel slot per optimised log-variance tensor, the
    rest summed into the last slot with its prior variance as a scale; contiguous parameter ranges merged), the
    constructor side effects of FunctionaLaplace (preds/laplace.py:140-160), and no CPU fallback."""
    pkg = _pkg()
    from preds_b200 import gps
    model = pkg.MLPS(3, [4], 2, 'tanh')
    params = list(model.parameters())                           # w1 (12), b1 (4), w2 (8), b2 (2)
    va, vb = torch.tensor(0.5), torch.tensor(-1.0)
    for i, p in enumerate(params):
        p.log_prior_var = va if i in (0, 1, 3) else vb
    all_vars, opt_vars, n_slots, slots, P = gps._groups(model, ())
    assert P == 26 and len(all_vars) == 2 and opt_vars == [] and n_slots == 1
    assert [(s, lo, hi) for s, _, lo, hi in slots] == [(0, 0, 16), (0, 16, 24), (0, 24, 26)]
    assert abs(slots[0][1] - float(va.exp())) < 1e-6 and abs(slots[1][1] - float(vb.exp())) < 1e-6
    all_vars, opt_vars, n_slots, slots, P = gps._groups(model, {va})
    assert n_slots == 2 and len(opt_vars) == 1 and opt_vars[0] is va
    assert [(s, sc is None, lo, hi) for s, sc, lo, hi in slots] == [(0, True, 0, 16), (1, False, 16, 24), (0, True, 24, 26)]
    all_vars, opt_vars, n_slots, slots, P = gps._groups(model, va)                 # a bare tensor is accepted too
    assert n_slots == 2
    all_vars, opt_vars, n_slots, slots, P = gps._groups(model, {va, vb})           # everything optimised: no rest slot
    assert n_slots == 2 and all(sc is None for _, sc, _, _ in slots)
    params[0].requires_grad_(False)
    with pytest.raises(pkg._cabi.BnnpError):
        gps._groups(model, ())
    params[0].requires_grad_(True)

    lh = pkg.CategoricalLh()
    lap = pkg.FunctionaLaplace(model, 0.25, lh)
    assert lap.P == 26 and not lap.inferred and model.likelihood is lh
    assert all(p.log_prior_var is params[0].log_prior_var for p in params)
    assert abs(float(params[0].log_prior_var) - float(-torch.log(torch.tensor(0.25)))) < 1e-7
    with pytest.raises(ValueError):
        lap.predictive_samples()
    with pytest.raises(AssertionError):
        pkg.FunctionaLaplace(model, torch.ones(26), lh)                          # np.isscalar(prior_prec)
    if not torch.cuda.is_available():
        with pytest.raises(pkg._cabi.BnnpError):
            lap.infer([(torch.randn(4, 3), None)], [(torch.randn(2, 3), None)])
        with pytest.raises(pkg._cabi.BnnpError):
            pkg.gp_quantities(model, torch.randn(4, 3))
    with pytest.raises(NotImplementedError):
        pkg.gp_quantities(model, torch.randn(4, 3), save_jacobians=True)


def test_factorisation_host_logic():
    """Tile ownership of the shared factorisation (128-column strips dealt out cyclically in GLOBAL coordinates) and the
    engine switch."""
    from preds_b200 import factor
    m = factor._own_cols_mask(1024, 300, 1, 4, 'cpu')
    cols = torch.arange(1024, 1324)
    assert torch.equal(m, (((cols // 128) % 4) == 1).float())
    total = sum(factor._own_cols_mask(2048, 1000, r, 3, 'cpu') for r in range(3))
    assert torch.equal(total, torch.ones(1000))                                  # every column has exactly one owner
    assert factor.use_tc(factor.MIN_P) and not factor.use_tc(factor.MIN_P - 1)
    old = factor.ENGINE
    factor.ENGINE = 'library'
    try:
        assert not factor.use_tc(100000)
    finally:
        factor.ENGINE = old
    assert factor._group(False) == (None, 0, 1) and factor._group(True) == (None, 0, 1)      # no process group here
    A = torch.eye(8) * 4
    assert torch.equal(factor.cholesky(A), torch.eye(8) * 2)                     # below MIN_P: the library call
