"""Parity of the CUDA path (preds_b200 -> C ABI -> sm_100a kernels) with (a) the committed golden
fixtures produced by the unmodified reference and (b) the CPU oracle on the same inputs.

Tolerance: 1e-4 relative (max |diff| / max |ref|), the north-star bound for fp32, stated per
assert.  Integer / index work (max-pool arg-max routing) is covered through the CNN cases, whose
Jacobians would differ at O(1) if a single index were routed differently.
"""
import numpy as np
import pytest
import torch

from oracle import port
from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-4
# Point-wise BNN samples push perturbed WEIGHTS through the network: eigensolver-level (cuSOLVER vs
# LAPACK) differences of the posterior factor are amplified, so these carry a looser stated bound.
TOL_BNN = 5e-4


def _product(name):
    import preds_b200
    from preds_b200 import models
    g = cases.load_golden(name)
    model = cases.make_product_model(name, models)
    port.set_theta(model, torch.from_numpy(g['theta']))
    model = model.cuda()
    lh = cases.make_lh(g['lh'], lib=preds_b200)
    batches, Xte = cases.case_data(g)
    return g, model, lh, batches, Xte, float(g['prior'])


@pytest.mark.parametrize('name', cases.FULL_CASES)
def test_jacobians_and_ggn_bundle(name):
    from preds_b200 import Jacobians, GGN
    g, model, lh, batches, Xte, prior = _product(name)
    X0, y0 = batches[0]
    Js, f = Jacobians(model, X0)
    assert tuple(Js.shape) == g['Js0'].shape
    assert cases.rel_err(f.cpu(), g['f0']) < 1e-5
    assert cases.rel_err(Js.cpu(), g['Js0']) < 1e-5
    Jb, Hess, rs, f2 = GGN(model, lh, X0, y0, ret_f=True)
    assert cases.rel_err(Hess.cpu(), g['Hess0']) < 1e-5
    assert cases.rel_err(rs.cpu(), g['rs0']) < 1e-5
    assert Jb.dim() == 3 and Hess.dim() == 3


@pytest.mark.parametrize('engine', [1, 2])
@pytest.mark.parametrize('name', cases.MLP_CASES)
def test_infer_full(name, engine):
    from preds_b200 import Laplace
    g, model, lh, batches, Xte, prior = _product(name)
    lap = Laplace(model, prior, lh)
    lap.ggn_engine = engine
    lap.infer(batches, cov_type='full')
    assert lap.inferred and lap.Sigma_chol.shape == (lap.P, lap.P)
    assert cases.rel_err(lap.precision.cpu(), g['H_full']) < TOL
    Sref = torch.from_numpy(g['Sigma_chol_full'])
    assert cases.rel_err(lap.Sigma.cpu(), Sref @ Sref.t()) < TOL
    assert cases.rel_err((lap.Sigma_chol @ lap.Sigma_chol.T).cpu(), Sref @ Sref.t()) < TOL
    # GLM predictive with the reference's injected draws
    eps = torch.from_numpy(g['eps_glm'])
    out = lap.predictive_samples_glm(Xte, n_samples=eps.shape[0], eps=eps)
    if lh.code == 1:
        assert cases.rel_err(out[0].cpu(), g['glm_fmu_full']) < 1e-5
        assert cases.rel_err(out[1].cpu(), g['glm_fvar_full']) < TOL
    else:
        assert tuple(out.shape) == g['glm_full'].shape
        assert cases.rel_err(out.cpu(), g['glm_full']) < TOL
    eb = torch.from_numpy(g['eps_bnn'])
    out = lap.predictive_samples_bnn(Xte, n_samples=eb.shape[1], eps=eb)
    assert cases.rel_err(out.cpu(), g['bnn_full']) < TOL_BNN


def test_infer_full_rejects_cnn():
    from preds_b200 import Laplace
    from preds_b200._cabi import BnnpError
    g, model, lh, batches, Xte, prior = _product('cnn_tiny')
    with pytest.raises(BnnpError):
        Laplace(model, prior, lh).infer(batches, cov_type='full')


@pytest.mark.parametrize('name', [c for c in cases.FULL_CASES if c != 'mlp_bern'])
def test_infer_diag(name):
    from preds_b200 import Laplace
    g, model, lh, batches, Xte, prior = _product(name)
    lap = Laplace(model, prior, lh)
    lap.infer(batches, cov_type='diag')
    assert cases.rel_err(lap.Sigma_chol.cpu(), g['Sigma_chol_diag']) < 1e-5
    eps = torch.from_numpy(g['eps_glm'])
    out = lap.predictive_samples_glm(Xte, n_samples=eps.shape[0], eps=eps)
    if lh.code == 1:
        assert cases.rel_err(out[1].cpu(), g['glm_fvar_diag']) < TOL
    else:
        assert cases.rel_err(out.cpu(), g['glm_diag']) < TOL
    eb = torch.from_numpy(g['eps_bnn'])
    out = lap.predictive_samples_bnn(Xte, n_samples=eb.shape[1], eps=eb)
    assert cases.rel_err(out.cpu(), g['bnn_diag']) < TOL_BNN


@pytest.mark.parametrize('name', [c for c in cases.FULL_CASES if c != 'mlp_bern'])
def test_infer_kron(name):
    from preds_b200 import Laplace, Kron
    from preds_b200.predictives import glm_moments
    g, model, lh, batches, Xte, prior = _product(name)
    lap = Laplace(model, prior, lh)
    lap.infer(batches, cov_type='kron')
    kron = lap.Sigma_chol
    assert type(kron) is Kron and lap.Sigma is kron
    for pi, qh in enumerate(kron.qhs):
        for fi, m in enumerate(qh):
            assert cases.rel_err(m.cpu(), g['kron_q_%d_%d' % (pi, fi)]) < 1e-5, (pi, fi)
    fmu, fvar = glm_moments(Xte, model, lap.mu, kron)
    assert cases.rel_err(fmu.cpu(), g['glm_fmu_kron']) < 1e-5
    assert cases.rel_err(fvar.cpu(), g['glm_fvar_kron']) < TOL
    eps = torch.from_numpy(g['eps_glm'])
    if lh.code != 1:
        out = lap.predictive_samples_glm(Xte, n_samples=eps.shape[0], eps=eps)
        assert cases.rel_err(out.cpu(), g['glm_kron']) < TOL
    ek = cases.kron_eps_from(g)
    S = eps.shape[0]
    out = lap.predictive_samples_bnn(Xte, n_samples=S, eps=ek)
    assert cases.rel_err(out.cpu(), g['bnn_kron']) < TOL_BNN
    # dampening toggled after inference (experiments/imginference.py:160)
    kron.dampen = True
    _, fvar_d = glm_moments(Xte, model, lap.mu, kron)
    assert cases.rel_err(fvar_d.cpu(), g['glm_fvar_kron_dampen']) < TOL
    out = lap.predictive_samples_bnn(Xte, n_samples=S, eps=ek)
    assert cases.rel_err(out.cpu(), g['bnn_kron_dampen']) < TOL_BNN


def test_cifar10net_probe():
    """The reference tests' own CIFAR10Net fixture (seed 71, X seed 15): conv / max-pool /
    dense Jacobians, diagonal GGN, KFLR factors and the Kron GLM predictive."""
    from preds_b200 import CIFAR10Net, CategoricalLh, Jacobians, Laplace
    from preds_b200.predictives import glm_moments
    g = cases.load_golden('cifar10net_probe')
    torch.manual_seed(71)
    model = CIFAR10Net(in_channels=2, n_out=3).cuda()
    torch.manual_seed(15)
    X = torch.randn(8, 2, 28, 28)
    y = torch.from_numpy(g['y'])
    idx = torch.from_numpy(g['theta_idx'])
    P = sum(p.numel() for p in model.parameters())
    v = torch.randn(P, generator=torch.Generator().manual_seed(4243))
    Js, f = Jacobians(model, X)
    assert cases.rel_err(f.cpu(), g['f']) < 1e-5
    assert cases.rel_err(Js[:, idx.cuda(), :].cpu(), g['Js_at_idx']) < 1e-5
    assert cases.rel_err(torch.einsum('npk,p->nk', Js, v.cuda()).cpu(), g['Js_dot_v']) < 1e-4
    del Js
    lh = CategoricalLh()
    batches = [(X[:5], y[:5]), (X[5:], y[5:])]
    lap = Laplace(model, 1.0, lh)
    lap.infer(batches, cov_type='diag')
    assert cases.rel_err(lap.precision[idx.cuda()].cpu(), g['diag_prec_at_idx']) < 1e-5
    assert abs(float(lap.precision.double().sum()) - float(g['diag_prec_sum'])) < 1e-4 * float(g['diag_prec_sum'])
    lap = Laplace(model, 1.0, lh)
    lap.infer(batches, cov_type='kron')
    for pi, qh in enumerate(lap.Sigma_chol.qhs):
        for fi, m in enumerate(qh):
            pv = torch.randn(m.shape[0], generator=torch.Generator().manual_seed(5000 + 10 * pi + fi))
            assert cases.rel_err((m @ pv.cuda()).cpu(), g['kron_q_%d_%d_mv' % (pi, fi)]) < 1e-4, (pi, fi)
    Xte = torch.randn(3, 2, 28, 28, generator=torch.Generator().manual_seed(16))
    fmu, fvar = glm_moments(Xte, model, lap.mu, lap.Sigma_chol)
    assert cases.rel_err(fmu.cpu(), g['glm_fmu_kron']) < 1e-5
    assert cases.rel_err(fvar.cpu(), g['glm_fvar_kron']) < TOL
    eps = torch.randn(4, 3, 3, generator=torch.Generator().manual_seed(711))
    out = lap.predictive_samples_glm(Xte, n_samples=4, eps=eps)
    assert cases.rel_err(out.cpu(), g['glm_kron']) < TOL


def test_cifar100net_probe():
    """CIFAR100Net (preds/models.py:182-236: nine convs with stride-2 and 1x1 ones, AvgPool2d(6), K = 100) on two
    3x32x32 examples against the reference golden: Jacobians, diagonal GGN, KFLR factors, Kron GLM moments -- and the
    wide-output likelihood / sampler kernels (csrc/wide_k.cu) against the oracle on the same injected draws."""
    from preds_b200 import CIFAR100Net, CategoricalLh, Jacobians, Laplace, GGN
    from preds_b200.predictives import glm_moments
    from tests import routing
    g = cases.load_golden('cifar100net_probe')
    torch.manual_seed(71)
    model = CIFAR100Net(3, 100)
    omodel = port.make_cifar100net(3, 100)
    port.set_theta(omodel, port.get_theta(model))
    model = model.cuda()
    X = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(15))
    y = torch.from_numpy(g['y'])
    idx = torch.from_numpy(g['theta_idx'])
    P = sum(p.numel() for p in model.parameters())
    assert P == 1387108
    v = torch.randn(P, generator=torch.Generator().manual_seed(4243))
    Js, f = Jacobians(model, X)
    assert tuple(Js.shape) == (2, P, 100)
    assert cases.rel_err(f.cpu(), g['f']) < 2e-5
    assert cases.rel_err(Js[:, idx.cuda(), :].cpu(), g['Js_at_idx']) < 5e-5
    assert cases.rel_err(torch.einsum('npk,p->nk', Js, v.cuda()).cpu(), g['Js_dot_v']) < 1e-4
    del Js
    # likelihood terms at K = 100 (warp-per-example kernels) against the oracle
    lh, olh = CategoricalLh(), port.CategoricalLh()
    fo = f.cpu()
    assert cases.rel_err(lh.Hessian(f).cpu(), olh.hessian(fo)) < 1e-5
    assert cases.rel_err(lh.residual(y.cuda(), f).cpu(), olh.residual(y, fo)) < 1e-5
    assert cases.rel_err(lh.inv_link(f).cpu(), olh.inv_link(fo)) < 1e-5
    lap = Laplace(model, 1.0, lh)
    lap.infer([(X, y)], cov_type='diag')
    assert cases.rel_err(lap.precision[idx.cuda()].cpu(), g['diag_prec_at_idx']) < TOL
    assert abs(float(lap.precision.double().sum()) - float(g['diag_prec_sum'])) < 1e-4 * float(g['diag_prec_sum'])
    lap = Laplace(model, 1.0, lh)
    lap.infer([(X, y)], cov_type='kron')
    for pi, qh in enumerate(lap.Sigma_chol.qhs):
        for fi, m in enumerate(qh):
            pv = torch.randn(m.shape[0], generator=torch.Generator().manual_seed(5000 + 10 * pi + fi))
            assert cases.rel_err((m @ pv.cuda()).cpu(), g['kron_q_%d_%d_mv' % (pi, fi)]) < TOL, (pi, fi)
    Xte = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(16))
    fmu, fvar = glm_moments(Xte, model, lap.mu, lap.Sigma_chol)
    assert cases.rel_err(fmu.cpu(), g['glm_fmu_kron']) < 2e-5
    assert cases.rel_err(fvar.cpu(), g['glm_fvar_kron']) < TOL
    # the wide sampler (block per input, 100 x 100 Cholesky in shared memory) on injected draws, with and without the
    # fused mean, against torch on the SAME moments
    S = 6
    eps = torch.randn(S, 1, 100, generator=torch.Generator().manual_seed(711))
    out = lap.predictive_samples_glm(Xte, n_samples=S, eps=eps)
    Lc = torch.linalg.cholesky(fvar.double().cpu())
    ref = torch.softmax(fmu.double().cpu().unsqueeze(0) + torch.einsum('nkl,snl->snk', Lc, eps.double()), dim=-1)
    assert tuple(out.shape) == (S, 1, 100)
    assert cases.rel_err(out.cpu(), ref) < TOL
    mean = lap.predictive_mean_glm(Xte, n_samples=S, eps=eps)
    assert cases.rel_err(mean.cpu(), ref.mean(0)) < TOL


def test_oracle_same_inputs_midsize():
    """CUDA path vs the CPU oracle on a banana-shaped (config 1) problem at a size the oracle
    finishes in seconds: MLPS(2,[50,50],2), N=600, two loader batches, all three structures."""
    from preds_b200 import MLPS, CategoricalLh, Laplace
    torch.manual_seed(71)
    model = MLPS(2, [50, 50], 2, 'tanh')
    X = torch.randn(600, 2, generator=torch.Generator().manual_seed(15))
    y = torch.randint(2, (600,), generator=torch.Generator().manual_seed(15))
    Xte = torch.randn(64, 2, generator=torch.Generator().manual_seed(16))
    batches = [(X[:400], y[:400]), (X[400:], y[400:])]
    omodel = port.make_mlps(2, [50, 50], 2, 'tanh')
    port.set_theta(omodel, port.get_theta(model))
    olh = port.CategoricalLh()
    model = model.cuda()
    S = 16
    eps = torch.randn(S, 64, 2, generator=torch.Generator().manual_seed(711))
    for cov in ('full', 'diag', 'kron'):
        post = port.infer(omodel, 1.0, olh, batches, cov)
        lap = Laplace(model, 1.0, CategoricalLh())
        lap.infer(batches, cov_type=cov)
        if cov == 'full':
            assert cases.rel_err(lap.precision.cpu(), post['precision']) < TOL
        ref = port.glm_predictive(omodel, olh, Xte, post, eps)
        out = lap.predictive_samples_glm(Xte, n_samples=S, eps=eps)
        assert cases.rel_err(out.cpu(), ref) < TOL, cov


def test_glm_sampler_philox_statistics():
    """In-kernel Philox draws: sample mean/covariance of the logits match (f_mu, f_var)."""
    from preds_b200._cabi import lib, check, ptr, stream, LIK_GAUSSIAN
    N, K, S = 4, 3, 200000
    g = torch.Generator().manual_seed(5)
    A = torch.randn(N, K, K, generator=g)
    fvar = (A @ A.transpose(1, 2) + 0.1 * torch.eye(K)).cuda().contiguous()
    fmu = torch.randn(N, K, generator=g).cuda().contiguous()
    out = torch.empty(S, N, K, device='cuda')
    status = torch.empty(N, dtype=torch.int32, device='cuda')
    check(lib.bnnp_glm_sample(LIK_GAUSSIAN, ptr(fmu), ptr(fvar), None, 1234, S, N, K, ptr(out), ptr(status), stream()))
    assert int(status.sum()) == 0
    assert float((out.mean(0) - fmu).abs().max()) < 0.02
    for n in range(N):
        cov = torch.cov(out[:, n, :].T)
        assert float((cov - fvar[n]).abs().max()) < 0.05 * float(fvar[n].abs().max())
    # a non-PD covariance is flagged and the reference's diagonal fallback is used
    fvar[1] = -fvar[1]
    check(lib.bnnp_glm_sample(LIK_GAUSSIAN, ptr(fmu), ptr(fvar), None, 1234, 8, N, K, ptr(out), ptr(status), stream()))
    assert status.cpu().tolist() == [0, 1, 0, 0]


@pytest.mark.parametrize('engine', [1, 2, 3])
@pytest.mark.parametrize('arch', [(20, [30, 12], 5, 70), (64, [40], 10, 300), (3, [4], 1, 9)])
def test_glm_fvar_full_engines(arch, engine):
    """J Sigma J^T on all engines (SIMT as written / tcgen05 generated / tcgen05 TMA-fed) vs fp64 from materialised
    Jacobians: ragged tile tails, several layers (cross-layer unit pairs), K = 1."""
    from preds_b200 import MLPS, Jacobians
    from preds_b200.predictives import glm_moments
    din, hid, K, N = arch
    torch.manual_seed(3)
    model = MLPS(din, hid, K, 'tanh').cuda()
    P = sum(p.numel() for p in model.parameters())
    X = torch.randn(N, din, device='cuda')
    A = torch.randn(P, P, device='cuda') / P ** 0.5
    Sigma = (A @ A.T + 0.05 * torch.eye(P, device='cuda')).contiguous()
    mu = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    Js, f = Jacobians(model, X)
    Jt = (Js.transpose(1, 2) if Js.dim() == 3 else Js.unsqueeze(1)).double()
    ref = torch.einsum('nkp,pq,ncq->nkc', Jt, Sigma.double(), Jt)
    fmu, fvar = glm_moments(X, model, mu, Sigma, engine=engine)
    assert cases.rel_err(fmu.cpu(), f.cpu()) < 1e-6
    assert cases.rel_err(fvar.cpu(), ref.cpu()) < (1e-5 if engine == 1 else 3e-5)


@pytest.mark.parametrize('arch', [
    dict(din=7, hid=[5], K=3, N=33, bias=True, act='tanh'),            # everything smaller than one tile
    dict(din=130, hid=[17], K=4, N=200, bias=True, act='relu'),        # row-tile tails, relu masks
    dict(din=64, hid=[40, 24], K=5, N=700, bias=True, act='tanh'),     # cross-layer blocks, N % 64 != 0
    dict(din=300, hid=[9], K=2, N=97, bias=False, act='tanh'),         # no bias rows / columns, 2 column tiles
    dict(din=33, hid=[20, 11, 6], K=1, N=150, bias=True, act='tanh'),  # K = 1 (regression-shaped Lambda), 4 layers
    dict(din=20, hid=[16], K=3, N=9000, bias=True, act='tanh'),        # three 4096-example accumulation segments
])
def test_full_ggn_tensor_core_shape_sweep(arch):
    """gram_tc (engine 2) vs fp64 sum_n J^T Lambda J from materialised Jacobians, and vs the SIMT engine."""
    from preds_b200 import MLPS, CategoricalLh, GaussianLh, Laplace, Jacobians
    torch.manual_seed(11)
    model = MLPS(arch['din'], arch['hid'], arch['K'], arch['act'], bias=arch['bias']).cuda()
    P = sum(p.numel() for p in model.parameters())
    N, K = arch['N'], arch['K']
    g = torch.Generator().manual_seed(12)
    X = torch.randn(N, arch['din'], generator=g).cuda()
    lh = GaussianLh(0.8) if K == 1 else CategoricalLh()
    y = torch.randn(N, 1, generator=g).cuda() if K == 1 else torch.randint(K, (N,), generator=g).cuda()
    H64 = torch.zeros(P, P, dtype=torch.float64, device='cuda')
    for s in range(0, N, 1024):
        Js, f = Jacobians(model, X[s:s + 1024])
        Jd = (Js if Js.dim() == 3 else Js.unsqueeze(2)).double()
        H64 += torch.einsum('mpk,mkl,mql->pq', Jd, lh.Hessian(f).double().reshape(len(f), K, K), Jd)
    errs = {}
    for eng in (1, 2):
        lap = Laplace(model, 0.0, lh)
        lap.ggn_engine = eng
        lap.infer([(X[:N // 3], y[:N // 3]), (X[N // 3:], y[N // 3:])], cov_type='full', factorise=False)
        H = lap.precision.double()
        assert torch.equal(lap.precision, lap.precision.T)
        errs[eng] = float((H - H64).abs().max() / H64.abs().max())
    assert errs[1] < 1e-5, errs
    assert errs[2] < 3e-5, errs
