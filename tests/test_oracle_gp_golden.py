"""Oracle restatement of the function-space path against the outputs of the UNMODIFIED reference
(tests/golden/gp_*.npz; ``preds/gps.py:664-927`` gp_quantities, ``:384-447`` GP_predictive, ``:930-952`` mc_preds,
``preds/laplace.py:138-207`` FunctionaLaplace)."""
import pytest
import torch

from oracle import port
from tests import cases

TOL = 2e-5


def _setup(name):
    g = cases.load_golden(name)
    model = cases.make_gp_port_model(name)
    port.set_theta(model, torch.from_numpy(g['theta']))
    Xtr = torch.cat([torch.from_numpy(g['Xa']), torch.from_numpy(g['Xb'])])
    return g, model, Xtr, torch.from_numpy(g['Xte'])


@pytest.mark.parametrize('name', cases.GP_CASES)
def test_gp_quantities(name):
    g, model, Xtr, Xte = _setup(name)
    q = port.gp_quantities(model, Xtr, Xte)
    for k, v in q.items():
        assert tuple(v.shape) == g['q_' + k].shape, k
        assert cases.rel_err(v, g['q_' + k]) < TOL, k
    # prior variance inside the kernel, two outputs, full test block, groups not collapsed (one group: slot axis of 1)
    C = model.output_size
    prior = float(g['prior'])
    q2 = port.gp_quantities(model, Xtr, Xte, prior_var=1.0 / prior, outputs=[0, C - 1], only_pred_var=False)
    assert abs(float(g['q_prior_var']) - 1.0 / prior) < 1e-6
    for k, v in q2.items():
        want = torch.from_numpy(g['q2_' + k])
        if k.startswith('K_'):
            assert want.shape[1] == 1
            want = want[:, 0]
        assert tuple(v.shape) == tuple(want.shape), k
        assert cases.rel_err(v, want) < TOL, k


@pytest.mark.parametrize('name', cases.GP_CASES)
def test_gp_predictive(name):
    g, model, Xtr, Xte = _setup(name)
    q = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('q_') and v.ndim > 0}
    prior = float(g['prior'])
    mean, cov = port.gp_predictive_moments(q, prior, indep=False)
    assert cases.rel_err(cov, g['gp_Sigma']) < 1e-4
    assert cases.rel_err(mean, q['f_star_test'].t()) == 0.0
    for indep in (0, 1):
        mean, cov = port.gp_predictive_moments(q, prior, indep=bool(indep))
        p = port.gp_mc_mean(mean, cov, torch.from_numpy(g['eps_%d' % indep]))
        assert float((p - torch.from_numpy(g['pred_%d' % indep])).abs().max()) < 1e-5, indep
