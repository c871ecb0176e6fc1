"""Oracle restatements of the post-predictive steps against the reference's outputs."""
import numpy as np
import pytest
import torch

from oracle import port
from tests import cases


def test_metrics():
    g = cases.load_golden('extras')
    p, y = torch.from_numpy(g['cat_p']), torch.from_numpy(g['cat_y'])
    for bins in (10, 15):
        m = port.classification_metrics(p, y, 'categorical', bins)
        assert m['acc'] == float(g['cat_acc']) == float(g['cat_macc'])        # integer work: exact
        assert abs(m['nll'] - float(g['cat_nll'])) < 1e-5 * abs(float(g['cat_nll']))
        assert abs(m['ece'] - float(g['cat_ece%d' % bins])) < 1e-5
    m = port.classification_metrics(torch.from_numpy(g['bern_p']), torch.from_numpy(g['bern_y']), 'bernoulli', 10)
    assert m['acc'] == float(g['bern_acc'])
    assert abs(m['nll'] - float(g['bern_nll'])) < 1e-5 * abs(float(g['bern_nll']))
    assert abs(m['ece'] - float(g['bern_ece10'])) < 1e-5


@pytest.mark.parametrize('tag', ['reg', 'cls'])
def test_linear_regression_predictive(tag):
    g = cases.load_golden('extras')
    K, lh = (1, port.GaussianLh(0.6)) if tag == 'reg' else (3, port.CategoricalLh())
    model = port.make_mlps(4, [9], K, 'tanh')
    port.set_theta(model, torch.from_numpy(g[tag + '_theta']))
    X, mu, chol = (torch.from_numpy(g[tag + k]) for k in ('_X', '_mu', '_chol'))
    Sigma = chol @ chol.t()
    for ctag, c in (('full', chol), ('diag', torch.sqrt(torch.diag(Sigma)))):
        ms, vf, vn = port.linear_regression_predictive(model, lh, X, mu, c)
        for name, got in (('mu_star', ms), ('var_f', vf), ('var_noise', vn)):
            want = g['%s_%s_%s' % (tag, ctag, name)]
            assert tuple(got.shape) == want.shape
            assert cases.rel_err(got, want) < 1e-4, (ctag, name)
