"""Function-space inference on the device (preds/gps.py:664-927, preds/laplace.py:138-207):
``gp_quantities`` / ``FunctionaLaplace`` against the goldens of the unmodified reference and against the oracle at sizes
where the P axis is walked in several tensor-core segments."""
import pytest
import torch

from oracle import port
from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _setup(name):
    import preds_b200
    from preds_b200 import models
    g = cases.load_golden(name)
    model = cases.make_gp_product_model(name, models)
    torch.nn.utils.vector_to_parameters(torch.from_numpy(g['theta']), model.parameters())
    return g, model.cuda(), preds_b200


@pytest.fixture(params=[True, False], ids=['structured', 'materialised'])
def ntk_route(request):
    """Linear-only networks have two routes to the kernels (Hadamard of short Grams / materialised one-column Jacobians);
    conv nets always take the second."""
    from preds_b200 import gps
    old = gps.STRUCTURED
    gps.STRUCTURED = request.param
    yield request.param
    gps.STRUCTURED = old


@pytest.mark.parametrize('name', cases.GP_CASES)
def test_functional_laplace_against_reference_golden(name, ntk_route):
    g, model, pb = _setup(name)
    Xa, Xb, Xte = (torch.from_numpy(g[k]) for k in ('Xa', 'Xb', 'Xte'))
    prior = float(g['prior'])
    lap = pb.FunctionaLaplace(model, prior, pb.CategoricalLh())
    with pytest.raises(ValueError):
        lap.predictive_samples()
    lap.infer([(Xa, None), (Xb, None)], [(Xte, None)], batch_size_per_gpu=5)
    q = lap.gp_quantities
    for k, v in q.items():
        if torch.is_tensor(v):
            assert v.is_cuda and tuple(v.shape) == g['q_' + k].shape, k
            assert cases.rel_err(v.cpu(), g['q_' + k]) < TOL, k
    assert abs(q['prior_var'] - float(g['q_prior_var'])) < 1e-6
    f_mu, f_var = lap.predictive_moments(indep=False)
    assert cases.rel_err(f_var.cpu(), g['gp_Sigma']) < 5e-4
    for indep in (0, 1):
        eps = torch.from_numpy(g['eps_%d' % indep])
        p = lap.predictive_samples(n_samples=eps.shape[0], batch_size=eps.shape[0], indep=bool(indep), eps=eps)
        assert tuple(p.shape) == g['pred_%d' % indep].shape
        assert float((p.cpu() - torch.from_numpy(g['pred_%d' % indep])).abs().max()) < 1e-4, indep
    # without injected draws: a valid probability table close to the injected-draw one (Monte-Carlo error only)
    p2 = lap.predictive_samples(n_samples=4000, batch_size=1000)
    assert float((p2.sum(1) - 1).abs().max()) < 1e-5
    assert float((p2 - p).abs().max()) < 0.1


@pytest.mark.parametrize('name', cases.GP_CASES)
def test_gp_quantities_option_paths(name):
    """The dictionary as ``preds.gps.gp_quantities`` returns it (host tensors): prior variance inside the kernel, groups
    not collapsed, the full test block, a subset of the outputs; argument checks of preds/gps.py:668-693."""
    g, model, pb = _setup(name)
    Xtr = torch.cat([torch.from_numpy(g['Xa']), torch.from_numpy(g['Xb'])])
    Xte = torch.from_numpy(g['Xte'])
    C = model.output_size
    log_var = -torch.log(torch.tensor(float(g['prior'])))
    for p in model.parameters():
        p.log_prior_var = log_var
    q2 = pb.gp_quantities(model, Xtr, Xte, outputs=[0, C - 1], only_pred_var=False, collapse_groups=False,
                          with_prior_prec=True, batch_size_per_gpu=4)
    for k in ('K_train', 'K_train_test', 'K_test', 'f_gp_train', 'f_star_train', 'f_gp_test', 'f_star_test'):
        assert not q2[k].is_cuda and tuple(q2[k].shape) == g['q2_' + k].shape, k
        assert cases.rel_err(q2[k], g['q2_' + k]) < TOL, k
    assert q2['prior_var'] is None
    with pytest.raises(ValueError):
        pb.gp_quantities(model, Xtr, Xte, outputs=C)
    with pytest.raises(RuntimeError):
        pb.gp_quantities(model, Xtr, Xte, max_ram_gb=1e-9)
    only_train = pb.gp_quantities(model, Xtr, with_prior_prec=False)
    assert 'K_test' not in only_train and cases.rel_err(only_train['K_train'][:, 0], g['q_K_train']) < TOL


def test_gp_kernels_over_several_k_segments_and_prior_groups(ntk_route):
    """P = 10 309 (three 4096-wide segments of the tensor-core product), 300 x 200 inputs, against the oracle; then two
    prior-variance groups of which one is 'optimised' (own kernel slot, unscaled) against a float64 evaluation from the
    oracle's Jacobians."""
    import preds_b200 as pb
    torch.manual_seed(5)
    model = pb.MLPS(60, [128, 20], 9, 'tanh')
    omodel = port.make_mlps(60, [128, 20], 9, 'tanh')
    port.set_theta(omodel, port.get_theta(model))
    g = torch.Generator().manual_seed(6)
    Xtr, Xte = torch.randn(300, 60, generator=g), torch.randn(200, 60, generator=g)
    model = model.cuda()
    lap = pb.FunctionaLaplace(model, 2.0, pb.CategoricalLh())
    lap.infer([(Xtr, None)], [(Xte, None)])
    oq = port.gp_quantities(omodel, Xtr, Xte)
    for k, v in oq.items():
        assert cases.rel_err(lap.gp_quantities[k].cpu(), v) < TOL, k
    for indep in (False, True):
        f_mu, f_var = lap.predictive_moments(indep=indep)
        om, ov = port.gp_predictive_moments(oq, 2.0, indep=indep)
        assert cases.rel_err(f_mu.cpu(), om) < TOL
        assert cases.rel_err(f_var.cpu(), ov) < 1e-3, indep
    # two groups: the first layer's parameters share an optimised log-variance, the rest a fixed one
    params = list(model.parameters())
    va, vb = torch.tensor(0.3), torch.tensor(-0.4)
    for i, p in enumerate(params):
        p.log_prior_var = va if i < 2 else vb
    q = pb.gp_quantities(model, Xtr[:40], Xte[:30], outputs=[2], optimize_params={va}, collapse_groups=False)
    J, _ = port.jacobians(omodel, Xtr[:40])
    Jt, _ = port.jacobians(omodel, Xte[:30])
    J, Jt = J[:, :, 2].double(), Jt[:, :, 2].double()
    n0 = params[0].numel() + params[1].numel()
    want0 = J[:, :n0] @ J[:, :n0].T
    want1 = float(vb.exp()) * (J[:, n0:] @ J[:, n0:].T)
    assert tuple(q['K_train'].shape) == (1, 2, 40, 40)
    assert cases.rel_err(q['K_train'][0, 0], want0) < TOL and cases.rel_err(q['K_train'][0, 1], want1) < TOL
    assert cases.rel_err(q['K_train_test'][0, 1], float(vb.exp()) * (J[:, n0:] @ Jt[:, n0:].T)) < TOL
    assert cases.rel_err(q['K_test'][0, 0], (Jt[:, :n0] ** 2).sum(1)) < TOL
    qc = pb.gp_quantities(model, Xtr[:40], Xte[:30], outputs=[2], optimize_params={va}, collapse_groups=True)
    assert cases.rel_err(qc['K_train'][0], float(va.exp()) * want0 + want1) < TOL


def test_gp_quantities_single_output_model(ntk_route):
    """K = 1 (regression-style network, MLPS(..., 1)): gp_quantities has one output slot; against the oracle, ragged sizes."""
    import preds_b200 as pb
    torch.manual_seed(9)
    model = pb.MLPS(7, [11, 5], 1, 'tanh')
    omodel = port.make_mlps(7, [11, 5], 1, 'tanh')
    port.set_theta(omodel, port.get_theta(model))
    g = torch.Generator().manual_seed(10)
    Xtr, Xte = torch.randn(37, 7, generator=g), torch.randn(13, 7, generator=g)
    model = model.cuda()
    log_var = torch.tensor(0.0)
    for p in model.parameters():
        p.log_prior_var = log_var
    q = pb.gp_quantities(model, Xtr, Xte, collapse_groups=True, with_prior_prec=False)
    oq = port.gp_quantities(omodel, Xtr, Xte)
    for k, v in oq.items():
        assert tuple(q[k].shape) == tuple(v.shape), k
        assert cases.rel_err(q[k], v) < TOL, k
    assert q['prior_var'] == 1.0
