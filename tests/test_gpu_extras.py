"""SURVEY.md 8f rank 3 on the GPU: metric reductions (exact integer work), the delta-method regression predictive
and the sample-mean fused into the GLM sampler -- against goldens from the unmodified reference, the CPU oracle
and size-independent properties."""
import numpy as np
import pytest
import torch

from oracle import port
from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-4


def test_metrics_against_reference_golden():
    from preds_b200 import CategoricalLh, BernoulliLh
    from preds_b200.utils import acc, macc, nll_cls, ece, classification_metrics
    g = cases.load_golden('extras')
    p, y = torch.from_numpy(g['cat_p']).cuda(), torch.from_numpy(g['cat_y']).cuda()
    lh = CategoricalLh()
    assert acc(p, y, lh) == float(g['cat_acc'])               # arg-max / compare / count: exact
    assert macc(p, y) == float(g['cat_macc'])
    assert abs(nll_cls(p, y, lh) - float(g['cat_nll'])) < 1e-5 * float(g['cat_nll'])
    for bins in (10, 15):
        assert abs(ece(p, y, lh, bins=bins) - float(g['cat_ece%d' % bins])) < 1e-5
    m = classification_metrics(p, y, lh, bins=10)
    assert m['acc'] == float(g['cat_acc']) and abs(m['ece'] - float(g['cat_ece10'])) < 1e-5
    pb, yb = torch.from_numpy(g['bern_p']).cuda(), torch.from_numpy(g['bern_y']).cuda()
    lh = BernoulliLh()
    assert acc(pb, yb, lh) == float(g['bern_acc'])
    assert abs(nll_cls(pb, yb, lh) - float(g['bern_nll'])) < 1e-5 * float(g['bern_nll'])
    assert abs(ece(pb, yb, lh, bins=10) - float(g['bern_ece10'])) < 1e-5
    with pytest.raises(ValueError):
        nll_cls(p, y, None)


def test_metrics_large_against_oracle():
    """10^6 predictions x 10 classes: counts are exact, the fp64-accumulated sums agree to 1e-6."""
    from preds_b200 import CategoricalLh
    from preds_b200.utils import classification_metrics
    g = torch.Generator().manual_seed(5)
    p = torch.softmax(2 * torch.randn(1_000_000, 10, generator=g), dim=1)
    y = torch.randint(10, (1_000_000,), generator=g)
    want = port.classification_metrics(p, y, 'categorical', 10)
    got = classification_metrics(p.cuda(), y.cuda(), CategoricalLh(), bins=10)
    assert got['acc'] == want['acc']
    assert abs(got['nll'] - want['nll']) < 1e-5 * want['nll']
    assert abs(got['ece'] - want['ece']) < 1e-6


@pytest.mark.parametrize('tag', ['reg', 'cls'])
def test_linear_regression_predictive(tag):
    import preds_b200
    from preds_b200.models import MLPS
    from preds_b200.predictives import linear_regression_predictive
    g = cases.load_golden('extras')
    K, lh = (1, preds_b200.GaussianLh(0.6)) if tag == 'reg' else (3, preds_b200.CategoricalLh())
    model = MLPS(4, [9], K, 'tanh')
    port.set_theta(model, torch.from_numpy(g[tag + '_theta']))
    model = model.cuda()
    X, mu, chol = (torch.from_numpy(g[tag + k]).cuda() for k in ('_X', '_mu', '_chol'))
    Sigma = chol @ chol.T
    for ctag, c in (('full', chol), ('diag', torch.sqrt(torch.diag(Sigma)).contiguous())):
        ms, vf, vn = linear_regression_predictive(X, model, lh, mu, c)
        for name, got in (('mu_star', ms), ('var_f', vf), ('var_noise', vn)):
            want = g['%s_%s_%s' % (tag, ctag, name)]
            assert tuple(got.shape) == want.shape
            assert cases.rel_err(got.cpu(), want) < TOL, (ctag, name)


@pytest.mark.parametrize('cov', ['full', 'diag', 'kron'])
def test_fused_sample_mean_equals_mean_of_samples(cov):
    """predictive_mean_glm == predictive_samples_glm(...).mean(0) for the same injected draws, and for in-kernel
    Philox draws with the same seed (the fused kernel walks the same counter space)."""
    from preds_b200 import MLPS, CategoricalLh, Laplace
    torch.manual_seed(71)
    model = MLPS(8, [16], 5, 'tanh').cuda()
    g = torch.Generator().manual_seed(15)
    X, y = torch.randn(200, 8, generator=g).cuda(), torch.randint(5, (200,), generator=g).cuda()
    Xte = torch.randn(333, 8, generator=g).cuda()
    lap = Laplace(model, 1.0, CategoricalLh())
    lap.infer([(X, y)], cov_type=cov)
    S = 200
    eps = torch.randn(S, 333, 5, generator=g).cuda()
    want = lap.predictive_samples_glm(Xte, n_samples=S, eps=eps).double().mean(0)
    got = lap.predictive_mean_glm(Xte, n_samples=S, eps=eps)
    assert tuple(got.shape) == (333, 5)
    assert cases.rel_err(got.cpu(), want.cpu()) < 1e-5
    want = lap.predictive_samples_glm(Xte, n_samples=S, seed=99).double().mean(0)
    got = lap.predictive_mean_glm(Xte, n_samples=S, seed=99)
    assert cases.rel_err(got.cpu(), want.cpu()) < 1e-5
    assert float((got.sum(-1) - 1).abs().max()) < 1e-5
