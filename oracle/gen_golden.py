"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (/root/reference/preds)
through ``oracle/refharness``.  Run in the build container only:

    python oracle/gen_golden.py

TEST INFRASTRUCTURE.  The fixtures travel to the GPU box; the reference does not.
Every case stores its inputs (X, y, theta, prior, injected normals) and the reference outputs:
Jacobians (preds/gradients.py:20-46), likelihood Hessian (preds/likelihoods.py), the full GGN
precision (preds/laplace.py:41-42), Sigma_chol for full/diag (preds/laplace.py:43-45,60), the
Kronecker factors (preds/laplace.py:69-78), and the GLM / BNN predictives
(preds/predictives.py:50-76, :12-28) for the injected draws.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refharness  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def randn(seed, *shape):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def flat(model):
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()])


def np_(t):
    return t.detach().cpu().numpy()


def kron_eps(model, S, seed):
    eps = []
    for i, p in enumerate(model.parameters()):
        if p.dim() == 1:
            eps.append(randn(seed + i, p.shape[0], S))
        else:
            eps.append(randn(seed + i, S, p.shape[0], p[0].numel()))
    return eps


def run_case(preds, name, model, lh, batches, Xte, prior, S, lh_desc, store_full=True):
    """Run every hot-path entry of the reference on one configuration."""
    from preds.laplace import Laplace
    from preds.gradients import Jacobians
    from preds.optimizers import GGN

    out = {'lh': np.array(lh_desc), 'prior': np.array(prior, dtype=np.float64),
           'theta': np_(flat(model)), 'Xte': np_(Xte), 'n_batches': np.array(len(batches))}
    for i, (X, y) in enumerate(batches):
        out['X%d' % i], out['y%d' % i] = np_(X), np_(y)
    P = len(out['theta'])
    K = model.output_size
    is_gauss = type(lh).__name__ == 'GaussianLh'
    can_backpack = type(lh).__name__ != 'BernoulliLh'

    # Jacobians + GGN bundle of the first batch
    X0, y0 = batches[0]
    Js, f = Jacobians(model, X0)
    out['Js0'], out['f0'] = np_(Js), np_(f)
    Jb, Hess, rs, _ = GGN(model, lh, X0, y0, ret_f=True)
    out['Hess0'], out['rs0'] = np_(Hess), np_(rs)

    # full GGN precision exactly as preds/laplace.py:38-42
    lap = Laplace(model, prior, lh)
    precision = lap.expand_prior_precision('full')
    for X, y in batches:
        Jb, Hess, rs, _ = GGN(model, lh, X, y, ret_f=True)
        precision += torch.einsum('mpk,mkl,mql->pq', Jb, Hess, Jb)
    out['H_full'] = np_(precision)

    eps_glm = randn(711, S, len(Xte), K)
    eps_bnn = randn(712, P, S)
    out['eps_glm'], out['eps_bnn'] = np_(eps_glm), np_(eps_bnn)

    cov_types = ['full'] + (['diag', 'kron'] if can_backpack else [])
    for cov in cov_types:
        lap = Laplace(model, prior, lh)
        lap.infer(batches, cov_type=cov)
        if cov == 'kron':
            kron = lap.Sigma_chol
            for pi, qh in enumerate(kron.qhs):
                for fi, m in enumerate(qh):
                    out['kron_q_%d_%d' % (pi, fi)] = np_(m)
        else:
            out['Sigma_chol_' + cov] = np_(lap.Sigma_chol)
        # GLM predictive
        with refharness.inject_normals([eps_glm]):
            res = lap.predictive_samples_glm(Xte, n_samples=S)
        if is_gauss:
            out['glm_fmu_' + cov], out['glm_fvar_' + cov] = np_(res[0]), np_(res[1])
        else:
            out['glm_' + cov] = np_(res)
            from preds.predictives import functional_sampling_predictive  # moments via Gaussian
            from preds.likelihoods import GaussianLh
            Sig = lap.Sigma if cov != 'kron' else lap.Sigma_chol
            fmu, fvar = functional_sampling_predictive(Xte, model, GaussianLh(), lap.mu, Sig, S)
            out['glm_fmu_' + cov], out['glm_fvar_' + cov] = np_(fmu), np_(fvar)
        # BNN predictive
        if cov == 'kron':
            eps_k = kron_eps(model, S, 900)
            for i, e in enumerate(eps_k):
                out['eps_kron_%d' % i] = np_(e)
            with refharness.inject_normals(eps_k):
                res = lap.predictive_samples_bnn(Xte, n_samples=S)
            # dampened variant toggled after inference (experiments/imginference.py:160)
            kron.dampen = True
            with refharness.inject_normals(eps_k):
                out['bnn_kron_dampen'] = np_(lap.predictive_samples_bnn(Xte, n_samples=S))
            from preds.predictives import functional_sampling_predictive
            from preds.likelihoods import GaussianLh
            fmu, fvar = functional_sampling_predictive(Xte, model, GaussianLh(), lap.mu, kron, S)
            out['glm_fvar_kron_dampen'] = np_(fvar)
            kron.dampen = False
        else:
            with refharness.inject_normals([eps_bnn]):
                res = lap.predictive_samples_bnn(Xte, n_samples=S)
        out['bnn_' + cov] = np_(res)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('%-12s P=%d K=%d  %.1f KB' % (name, P, K, os.path.getsize(os.path.join(OUT, name + '.npz')) / 1e3))


def ggn_case(preds, name, model, lh, lh_desc, X, y, Xte, prior_prec, prior_mu, S, n_steps=3):
    """the UCI classification driver's path (experiments/classification.py:20-112):
    a few LaplaceGGN.step()s, post_process, get_diagonal_ggn, and the linearised / BNN predictives with
    injected draws."""
    from preds.optimizers import LaplaceGGN, get_diagonal_ggn
    from preds.predictives import linear_sampling_predictive, nn_sampling_predictive
    out = {'lh': np.array(lh_desc), 'X': np_(X), 'y': np_(y), 'Xte': np_(Xte), 'theta0': np_(flat(model)),
           'prior_prec': np_(torch.as_tensor(prior_prec)), 'prior_mu': np_(torch.as_tensor(prior_mu)),
           'n_steps': np.array(n_steps)}
    opt = LaplaceGGN(model, lr=1e-2, prior_prec=prior_prec, prior_mu=prior_mu)
    losses = []
    for _ in range(n_steps):
        def closure():
            model.zero_grad()
            return lh.log_likelihood(y, model(X))
        losses.append(opt.step(closure))
    out['losses'], out['theta'] = np.array(losses), np_(flat(model))
    half = len(X) // 2
    opt.post_process(model, lh, [(X[:half], y[:half]), (X[half:], y[half:])])
    st = opt.state
    out['precision'], out['Sigma_chol'], out['mu'] = np_(st['precision']), np_(st['Sigma_chol']), np_(st['mu'])
    Sigmad, Sigma_chold = get_diagonal_ggn(opt)
    out['Sigma_d'], out['Sigma_chol_d'] = np_(Sigmad), np_(Sigma_chold)
    P = len(out['theta'])
    eps = randn(713, P, S)
    out['eps'] = np_(eps)
    theta_star = flat(model).detach()
    for tag, mu, chol in (('full', st['mu'], st['Sigma_chol']), ('map', theta_star, st['Sigma_chol']),
                          ('diag', theta_star, Sigma_chold), ('vec', st['mu'], torch.diag(Sigma_chold).clone())):
        with refharness.inject_normals([eps]):
            out['lin_' + tag] = np_(linear_sampling_predictive(Xte, model, lh, mu, chol, mc_samples=S))
    with refharness.inject_normals([eps]):
        out['lin_nolink'] = np_(linear_sampling_predictive(Xte, model, lh, st['mu'], st['Sigma_chol'],
                                                           mc_samples=S, no_link=True))
    with refharness.inject_normals([eps]):
        out['nn_full'] = np_(nn_sampling_predictive(Xte, model, lh, theta_star, st['Sigma_chol'], mc_samples=S))
    # GLM refinement on the same posterior (experiments/classification.py:137-174)
    from preds.refine import laplace_refine, vi_refine, vi_diag_refine
    m, S_chol, S_chold, losses = laplace_refine(model, X, y, lh, 0.7, n_epochs=6, lr=1e-2)
    out['lapref_mu'], out['lapref_chol'], out['lapref_chold'] = np_(m), np_(S_chol), np_(S_chold)
    out['lapref_losses'] = np.array(losses)
    eps_vi = [randn(800 + i, P) for i in range(4)]
    out['eps_vi'] = np_(torch.stack(eps_vi))
    with refharness.inject_normals(eps_vi):
        m, S_chol, losses = vi_refine(model, opt, X, y, lh, n_epochs=4, lr=0.05)
    out['vi_mu'], out['vi_chol'], out['vi_losses'] = np_(m), np_(S_chol), np.array(losses)
    with refharness.inject_normals(eps_vi):
        m, S_chold, losses = vi_diag_refine(model, opt, X, y, lh, n_epochs=4, lr=0.05)
    out['vid_mu'], out['vid_chol'], out['vid_losses'] = np_(m), np_(S_chold), np.array(losses)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('%-12s P=%d  %.1f KB' % (name, P, os.path.getsize(os.path.join(OUT, name + '.npz')) / 1e3))


def extras_case(preds, name):
    """metric reductions (preds/utils.py:6-52) on seeded probabilities -- including exact
    ties, confidences on bin boundaries and near-0/1 probabilities -- and linear_regression_predictive
    (preds/predictives.py:79-94) for a Gaussian and a categorical model."""
    from preds.utils import acc, macc, nll_cls, ece
    from preds.likelihoods import CategoricalLh, BernoulliLh, GaussianLh
    from preds.models import MLPS
    from preds.laplace import Laplace
    from preds.predictives import linear_regression_predictive
    out = {}
    g = torch.Generator().manual_seed(50)
    p = torch.softmax(3 * torch.randn(300, 5, generator=g), dim=1)
    p[0] = torch.tensor([0.5, 0.5, 0, 0, 0])               # tie: first index wins
    p[1] = torch.tensor([0.2, 0.2, 0.2, 0.2, 0.2])
    p[2] = torch.tensor([0.0, 0.0, 1.0, 0.0, 0.0])         # clamped log(0)
    p[3] = torch.tensor([0.3, 0.7, 0, 0, 0])               # confidence exactly on a bin edge (0.7 in fp32)
    p[4] = torch.tensor([0.1, 0.9, 0, 0, 0])
    y = torch.randint(5, (300,), generator=g)
    y[2] = 0
    out['cat_p'], out['cat_y'] = np_(p), np_(y)
    lh = CategoricalLh()
    out['cat_nll'], out['cat_acc'], out['cat_macc'] = (np.array(v) for v in (nll_cls(p, y, lh), acc(p, y, lh), macc(p, y)))
    for b in (10, 15):
        out['cat_ece%d' % b] = np.array(ece(p, y, lh, bins=b))
    pb = torch.sigmoid(2.5 * torch.randn(257, generator=g))
    pb[0], pb[1], pb[2], pb[3] = 0.5, 0.0, 1.0, 0.8
    yb = torch.randint(2, (257,), generator=g).float()
    out['bern_p'], out['bern_y'] = np_(pb), np_(yb)
    lh = BernoulliLh()
    out['bern_nll'], out['bern_acc'], out['bern_ece10'] = (np.array(v) for v in (nll_cls(pb, yb, lh), acc(pb, yb, lh),
                                                                                   ece(pb, yb, lh, bins=10)))
    # linear_regression_predictive
    for tag, K, lh, sig in (('reg', 1, GaussianLh(0.6), 0.6), ('cls', 3, CategoricalLh(), None)):
        torch.manual_seed(79)
        model = MLPS(4, [9], K, 'tanh')
        X = randn(51, 30, 4)
        y = randn(52, 30, 1) if K == 1 else torch.randint(3, (30,), generator=torch.Generator().manual_seed(52))
        lap = Laplace(model, 0.9, lh)
        lap.infer([(X, y)], cov_type='full')
        Xte = randn(53, 7, 4)
        mu = lap.mu + 0.05 * randn(54, len(lap.mu))
        out[tag + '_theta'], out[tag + '_X'], out[tag + '_mu'] = np_(flat(model)), np_(Xte), np_(mu)
        out[tag + '_chol'] = np_(lap.Sigma_chol)
        for ctag, chol in (('full', lap.Sigma_chol), ('diag', torch.sqrt(torch.diag(lap.Sigma)).clone())):
            ms, vf, vn = linear_regression_predictive(Xte, model, lh, mu, chol)
            out['%s_%s_mu_star' % (tag, ctag)], out['%s_%s_var_f' % (tag, ctag)] = np_(ms), np_(vf)
            out['%s_%s_var_noise' % (tag, ctag)] = np_(vn)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('%-12s %.1f KB' % (name, os.path.getsize(os.path.join(OUT, name + '.npz')) / 1e3))


def probe_case(preds, name):
    """Full-size CIFAR10Net (reference tests' own fixture: seed 71, in_channels=2, n_out=3,
    X seed 15 -- tests/test_jacobians.py:9-30): tensors too large to store, so the fixture
    holds seeded random probes of them."""
    from preds.models import CIFAR10Net
    from preds.likelihoods import CategoricalLh
    from preds.laplace import Laplace
    from preds.gradients import Jacobians
    torch.manual_seed(71)
    model = CIFAR10Net(in_channels=2, n_out=3)
    torch.manual_seed(15)
    X = torch.randn(8, 2, 28, 28)
    y = torch.softmax(randn(15, 8, 3), dim=1).argmax(dim=1)
    theta = flat(model)
    P = len(theta)
    g = torch.Generator().manual_seed(4242)
    idx = torch.randperm(P, generator=g)[:4096]
    v = randn(4243, P)
    out = {'theta_idx': np_(idx), 'theta_at_idx': np_(theta[idx]), 'theta_dot_v': np_(theta @ v),
           'y': np_(y)}
    Js, f = Jacobians(model, X)
    out['f'] = np_(f)
    out['Js_at_idx'] = np_(Js[:, idx, :])
    out['Js_dot_v'] = np_(torch.einsum('npk,p->nk', Js, v))
    lh = CategoricalLh()
    lap = Laplace(model, 1.0, lh)
    lap.infer([(X[:5], y[:5]), (X[5:], y[5:])], cov_type='diag')
    prec = 1 / lap.Sigma_chol ** 2
    out['diag_prec_at_idx'] = np_(prec[idx])
    out['diag_prec_sum'] = np_(prec.double().sum())
    lap = Laplace(model, 1.0, lh)
    lap.infer([(X[:5], y[:5]), (X[5:], y[5:])], cov_type='kron')
    for pi, qh in enumerate(lap.Sigma_chol.qhs):
        for fi, m in enumerate(qh):
            pv = randn(5000 + 10 * pi + fi, m.shape[0])
            out['kron_q_%d_%d_mv' % (pi, fi)] = np_(m @ pv)
            out['kron_q_%d_%d_tr' % (pi, fi)] = np_(m.diagonal().double().sum())
    Xte = randn(16, 3, 2, 28, 28)
    S = 4
    eps_glm = randn(711, S, 3, 3)
    with refharness.inject_normals([eps_glm]):
        out['glm_kron'] = np_(lap.predictive_samples_glm(Xte, n_samples=S))
    from preds.predictives import functional_sampling_predictive
    from preds.likelihoods import GaussianLh
    fmu, fvar = functional_sampling_predictive(Xte, model, GaussianLh(), lap.mu, lap.Sigma_chol, S)
    out['glm_fmu_kron'], out['glm_fvar_kron'] = np_(fmu), np_(fvar)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('%-12s P=%d  %.1f KB' % (name, P, os.path.getsize(os.path.join(OUT, name + '.npz')) / 1e3))


def probe100_case(preds, name):
    """CIFAR100Net (preds/models.py:182-236: nine convs incl. stride-2 and 1x1 ones, AvgPool2d(6), K = 100) on two
    3x32x32 examples: probes of the Jacobians, the diagonal GGN, the KFLR factors and the Kron GLM moments."""
    from preds.models import CIFAR100Net
    from preds.likelihoods import CategoricalLh, GaussianLh
    from preds.laplace import Laplace
    from preds.gradients import Jacobians
    from preds.predictives import functional_sampling_predictive
    torch.manual_seed(71)
    model = CIFAR100Net(in_channels=3, n_out=100)
    X = randn(15, 2, 3, 32, 32)
    y = torch.tensor([17, 93])
    theta = flat(model)
    P = len(theta)
    idx = torch.randperm(P, generator=torch.Generator().manual_seed(4242))[:512]
    v = randn(4243, P)
    out = {'theta_idx': np_(idx), 'theta_at_idx': np_(theta[idx]), 'theta_dot_v': np_(theta @ v), 'y': np_(y)}
    Js, f = Jacobians(model, X)
    out['f'] = np_(f)
    out['Js_at_idx'] = np_(Js[:, idx, :])
    out['Js_dot_v'] = np_(torch.einsum('npk,p->nk', Js, v))
    del Js
    lh = CategoricalLh()
    lap = Laplace(model, 1.0, lh)
    lap.infer([(X, y)], cov_type='diag')
    prec = 1 / lap.Sigma_chol ** 2
    out['diag_prec_at_idx'] = np_(prec[idx])
    out['diag_prec_sum'] = np_(prec.double().sum())
    lap = Laplace(model, 1.0, lh)
    lap.infer([(X, y)], cov_type='kron')
    for pi, qh in enumerate(lap.Sigma_chol.qhs):
        for fi, m in enumerate(qh):
            pv = randn(5000 + 10 * pi + fi, m.shape[0])
            out['kron_q_%d_%d_mv' % (pi, fi)] = np_(m @ pv)
            out['kron_q_%d_%d_tr' % (pi, fi)] = np_(m.diagonal().double().sum())
    Xte = randn(16, 1, 3, 32, 32)
    fmu, fvar = functional_sampling_predictive(Xte, model, GaussianLh(), lap.mu, lap.Sigma_chol, 2)
    out['glm_fmu_kron'], out['glm_fvar_kron'] = np_(fmu), np_(fvar)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('%-12s P=%d  %.1f KB' % (name, P, os.path.getsize(os.path.join(OUT, name + '.npz')) / 1e3))


def gp_case(preds, name, model, Xa, Xb, Xte, prior, S):
    """Function-space inference (preds/laplace.py:138-207, preds/gps.py:664-927) of the UNMODIFIED reference, run on the
    CPU through ``refharness.cuda_standins`` (the reference hard-codes CUDA devices; the stand-ins carry no arithmetic)."""
    from preds.laplace import FunctionaLaplace
    from preds.likelihoods import CategoricalLh
    from preds.gps import gp_quantities, GP_predictive
    from oracle.refharness.cuda_standins import cpu_as_cuda
    out = {'theta': np_(flat(model)), 'Xa': np_(Xa), 'Xb': np_(Xb), 'Xte': np_(Xte), 'prior': np.array(prior)}
    lh = CategoricalLh()
    lap = FunctionaLaplace(model, prior, lh)
    C = model.output_size
    with cpu_as_cuda():
        lap.infer([(Xa, None), (Xb, None)], [(Xte, None)], batch_size_per_gpu=5)
        q = lap.gp_quantities
        for k, v in q.items():
            if torch.is_tensor(v):
                out['q_' + k] = np_(v)
        out['q_prior_var'] = np.array(q['prior_var'])
        for indep in (0, 1):
            eps = randn(900 + indep, S, len(Xte), C)
            out['eps_%d' % indep] = np_(eps)
            with refharness.inject_normals([eps]):
                out['pred_%d' % indep] = np_(lap.predictive_samples(n_samples=S, batch_size=S, indep=bool(indep)))
        # the distribution behind indep=False before its mean is replaced (preds/laplace.py:200-205)
        fd = GP_predictive(lh, q['f_star_train'] - q['f_gp_train'], q['f_star_test'] - q['f_gp_test'],
                           q['K_train'] / prior, q['K_train_test'] / prior, q['K_test'] / prior, q['f_gp_train'],
                           out_device_id=-1, diag_only=True, batch_size_C=1, batch_size_M=len(Xte))
        out['gp_mu'], out['gp_Sigma'] = np_(fd.mean), np_(fd.covariance_matrix)
        # the other option paths: prior variance inside the kernel, groups not collapsed, full test block, two outputs
        Xtr = torch.cat([Xa, Xb])
        q2 = gp_quantities(model, Xtr, Xte, outputs=[0, C - 1], only_pred_var=False, collapse_groups=False,
                           with_prior_prec=True, batch_size_per_gpu=4)
        for k in ('K_train', 'K_train_test', 'K_test', 'f_gp_train', 'f_star_train', 'f_gp_test', 'f_star_test'):
            out['q2_' + k] = np_(q2[k])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('%-12s P=%d  %.1f KB' % (name, len(out['theta']), os.path.getsize(os.path.join(OUT, name + '.npz')) / 1e3))


def gp_cases(preds):
    from preds.models import MLPS
    from deepobs.pytorch.testproblems.testproblems_utils import tfconv2d, tfmaxpool2d
    torch.manual_seed(81)
    model = MLPS(5, [8, 6], 3, 'tanh')
    X = randn(40, 20, 5)
    gp_case(preds, 'gp_mlp', model, X[:12], X[12:], randn(41, 7, 5), 0.7, 64)
    torch.manual_seed(82)
    model = torch.nn.Sequential()
    model.add_module('conv1', tfconv2d(2, 4, 3, tf_padding_type='same'))
    model.add_module('relu1', torch.nn.ReLU())
    model.add_module('pool1', tfmaxpool2d(3, stride=2, tf_padding_type='same'))
    model.add_module('conv2', tfconv2d(4, 5, 3))
    model.add_module('tanh2', torch.nn.Tanh())
    model.add_module('flatten', torch.nn.Flatten())
    model.add_module('dense', torch.nn.Linear(45, 4))
    model.output_size = 4
    X = randn(42, 14, 2, 9, 9)
    gp_case(preds, 'gp_cnn', model, X[:8], X[8:], randn(43, 6, 2, 9, 9), 1.5, 48)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'gp':                     # the function-space cases only
        gp_cases(refharness.load())
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'cifar100net_probe':      # one case only (the others are unchanged)
        probe100_case(refharness.load(), 'cifar100net_probe')
        return
    preds = refharness.load()
    from preds.models import MLPS
    from preds.likelihoods import CategoricalLh, GaussianLh, BernoulliLh
    from deepobs.pytorch.testproblems.testproblems_utils import tfconv2d, tfmaxpool2d
    os.makedirs(OUT, exist_ok=True)

    # (A) banana-shaped classifier (config 1 shape, narrower), two loader batches
    torch.manual_seed(71)
    model = MLPS(2, [10, 10], 2, 'tanh')
    X = randn(15, 48, 2)
    y = torch.randint(2, (48,), generator=torch.Generator().manual_seed(15))
    run_case(preds, 'mlp_cls', model, CategoricalLh(), [(X[:32], y[:32]), (X[32:], y[32:])],
             randn(16, 6, 2), 0.7, 5, 'categorical')

    # (B) relu, K=3, single batch, confident logits (x6) to exercise the softmax clamp
    torch.manual_seed(72)
    model = MLPS(5, [12], 3, 'relu')
    with torch.no_grad():
        list(model.parameters())[-2].mul_(6.0)
    X = randn(17, 40, 5) * 3
    y = torch.randint(3, (40,), generator=torch.Generator().manual_seed(18))
    run_case(preds, 'mlp_relu3', model, CategoricalLh(), [(X, y)], randn(19, 7, 5), 1.3, 6,
             'categorical')

    # (C) regression K=1 (GaussianLh): predictive returns moments
    torch.manual_seed(73)
    model = MLPS(3, [8], 1, 'tanh')
    X = randn(20, 30, 3)
    y = randn(21, 30, 1)
    run_case(preds, 'mlp_reg', model, GaussianLh(0.6), [(X[:20], y[:20]), (X[20:], y[20:])],
             randn(22, 5, 3), 2.0, 4, 'gaussian:0.6')

    # (D) flatten + no bias
    torch.manual_seed(74)
    model = MLPS(16, [8, 6], 4, 'tanh', flatten=True, bias=False)
    X = randn(23, 24, 1, 4, 4)
    y = torch.randint(4, (24,), generator=torch.Generator().manual_seed(24))
    run_case(preds, 'mlp_flat', model, CategoricalLh(), [(X, y)], randn(25, 5, 1, 4, 4), 0.5, 5,
             'categorical')

    # (E) Bernoulli K=1: full covariance only (preds/likelihoods.py:74-75)
    torch.manual_seed(75)
    model = MLPS(4, [7], 1, 'tanh')
    X = randn(26, 36, 4)
    y = torch.randint(2, (36, 1), generator=torch.Generator().manual_seed(27)).float()
    run_case(preds, 'mlp_bern', model, BernoulliLh(), [(X, y)], randn(28, 5, 4), 1.0, 4,
             'bernoulli')

    # (F) tiny CNN from the same building blocks as preds/models.py:239-285
    torch.manual_seed(76)
    model = torch.nn.Sequential()
    model.add_module('conv1', tfconv2d(2, 4, 3, tf_padding_type='same'))
    model.add_module('relu1', torch.nn.ReLU())
    model.add_module('pool1', tfmaxpool2d(3, stride=2, tf_padding_type='same'))
    model.add_module('conv2', tfconv2d(4, 5, 3))
    model.add_module('tanh2', torch.nn.Tanh())
    model.add_module('flatten', torch.nn.Flatten())
    model.add_module('dense', torch.nn.Linear(45, 3))
    model.output_size = 3
    X = randn(29, 20, 2, 9, 9)
    y = torch.randint(3, (20,), generator=torch.Generator().manual_seed(30))
    run_case(preds, 'cnn_tiny', model, CategoricalLh(), [(X[:12], y[:12]), (X[12:], y[12:])],
             randn(31, 4, 2, 9, 9), 0.9, 5, 'categorical')

    # (H) LaplaceGGN path on the driver's own model class (functional activations), K=3, vector prior
    from preds.models import SiMLP
    torch.manual_seed(77)
    model = SiMLP(6, 3, 2, 8, activation='tanh')
    X = randn(32, 44, 6)
    y = torch.randint(3, (44,), generator=torch.Generator().manual_seed(33))
    P = sum(p.numel() for p in model.parameters())
    ggn_case(preds, 'ggn_simlp_cls', model, CategoricalLh(), 'categorical', X, y, randn(34, 9, 6),
             0.5 + torch.rand(P, generator=torch.Generator().manual_seed(35)), 0.1 * randn(36, P), 6)

    # (I) binary task: BernoulliLh with a single flattened output (experiments/classification.py:66-68)
    torch.manual_seed(78)
    model = SiMLP(4, 1, 1, 7, activation='tanh')
    X = randn(37, 40, 4)
    y = torch.randint(2, (40,), generator=torch.Generator().manual_seed(38)).float()
    ggn_case(preds, 'ggn_simlp_bern', model, BernoulliLh(), 'bernoulli', X, y, randn(39, 8, 4), 0.8, 0.0, 5)

    # (J) metric reductions and the delta-method regression predictive
    extras_case(preds, 'extras')

    # (G) the reference tests' own CIFAR10Net fixture, probed
    probe_case(preds, 'cifar10net_probe')

    # (K) CIFAR100Net, K = 100
    probe100_case(preds, 'cifar100net_probe')

    # (L) function-space inference
    gp_cases(preds)


if __name__ == '__main__':
    main()
