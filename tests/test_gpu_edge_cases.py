"""Edge cases of the drop-in API on the GPU, each checked against the CPU oracle on the same inputs:
single-example and ragged loader batches, vector / matrix priors, unknown cov_type, n_samples = 1,
host-resident inputs, shifted posterior mean (mu != theta*), Sigmoid / AvgPool / stride-2 layers."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import port
from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _pair(din=5, hid=(7,), K=3, act='tanh', bias=True, seed=71):
    from preds_b200 import MLPS
    torch.manual_seed(seed)
    model = MLPS(din, list(hid), K, act, bias=bias)
    omodel = port.make_mlps(din, list(hid), K, act, bias=bias)
    port.set_theta(omodel, port.get_theta(model))
    return model.cuda(), omodel


def _data(N, din, K, seed=15):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(N, din, generator=g), torch.randint(K, (N,), generator=g)


@pytest.mark.parametrize('cov', ['full', 'kron', 'diag'])
def test_ragged_and_single_example_batches(cov):
    from preds_b200 import Laplace, CategoricalLh
    model, omodel = _pair()
    X, y = _data(23, 5, 3)
    batches = [(X[:1], y[:1]), (X[1:9], y[1:9]), (X[9:22], y[9:22]), (X[22:], y[22:])]   # sizes 1, 8, 13, 1
    post = port.infer(omodel, 0.5, port.CategoricalLh(), batches, cov)
    lap = Laplace(model, 0.5, CategoricalLh())
    lap.infer(batches, cov_type=cov)
    Xte, _ = _data(1, 5, 3, seed=16)                                                       # a single test input
    eps = torch.randn(1, 1, 3, generator=torch.Generator().manual_seed(1))                # a single sample
    ref = port.glm_predictive(omodel, port.CategoricalLh(), Xte, post, eps)
    out = lap.predictive_samples_glm(Xte, n_samples=1, eps=eps)
    assert tuple(out.shape) == (1, 1, 3)
    assert cases.rel_err(out.cpu(), ref) < TOL


def test_vector_and_matrix_priors_full():
    from preds_b200 import Laplace, CategoricalLh
    model, omodel = _pair()
    X, y = _data(30, 5, 3)
    P = port.n_params(omodel)
    g = torch.Generator().manual_seed(3)
    vec = torch.rand(P, generator=g) + 0.5
    A = torch.randn(P, P, generator=g) * 0.05
    mat = A @ A.T + torch.diag(vec)
    for prior in (vec, mat):
        post = port.infer(omodel, prior, port.CategoricalLh(), [(X, y)], 'full')
        keep = prior.clone()
        lap = Laplace(model, prior.cuda(), CategoricalLh())
        lap.infer([(X, y)], cov_type='full')
        assert cases.rel_err(lap.precision.cpu(), post['precision']) < TOL
        assert torch.equal(prior, keep)                          # the caller's prior is never mutated
    post = port.infer(omodel, vec, port.CategoricalLh(), [(X, y)], 'diag')
    lap = Laplace(model, vec.cuda(), CategoricalLh())
    lap.infer([(X, y)], cov_type='diag')
    assert cases.rel_err(lap.Sigma_chol.cpu(), post['Sigma_chol']) < 1e-5
    lap = Laplace(model, mat.cuda(), CategoricalLh())
    with pytest.raises(ValueError, match='Will not diagonalize'):
        lap.infer([(X, y)], cov_type='diag')
    with pytest.raises(AssertionError):                         # kron needs a scalar prior (preds/laplace.py:66)
        Laplace(model, vec.cuda(), CategoricalLh()).infer([(X, y)], cov_type='kron')


def test_unknown_cov_type_and_guards():
    """An unknown cov_type silently does nothing but marks the object inferred (preds/laplace.py:37-82)."""
    from preds_b200 import Laplace, CategoricalLh, BernoulliLh
    model, _ = _pair()
    X, y = _data(8, 5, 3)
    lap = Laplace(model, 1.0, CategoricalLh())
    with pytest.raises(ValueError, match='Need to infer first'):
        lap.predictive_samples_bnn(X)
    lap.infer([(X, y)], cov_type='lowrank')
    assert lap.inferred and lap.Sigma_chol is None
    with pytest.raises(ValueError):                             # Bernoulli has no nn_loss -> full only (:74-75)
        Laplace(model, 1.0, BernoulliLh()).infer([(X, y)], cov_type='diag')


def test_host_inputs_and_dataloader():
    """Inputs may live on the host (the reference moves them, preds/laplace.py:40) and the loader may be
    a DataLoader (N from len(loader.dataset), :68)."""
    from preds_b200 import Laplace, CategoricalLh
    from torch.utils.data import DataLoader, TensorDataset
    model, omodel = _pair()
    X, y = _data(50, 5, 3)
    loader = DataLoader(TensorDataset(X, y), batch_size=16, shuffle=False)
    post = port.infer(omodel, 1.0, port.CategoricalLh(), list(loader), 'kron')
    lap = Laplace(model, 1.0, CategoricalLh())
    lap.infer(loader, cov_type='kron')
    for qa, qb in zip(lap.Sigma_chol.qhs, post['kron'].qhs):
        for a, b in zip(qa, qb):
            assert cases.rel_err(a.cpu(), b) < 1e-5
    Xte, _ = _data(4, 5, 3, seed=16)
    out_host = lap.predictive_samples_glm(Xte, n_samples=3, seed=7)
    out_dev = lap.predictive_samples_glm(Xte.cuda(), n_samples=3, seed=7)
    assert torch.equal(out_host, out_dev)                       # same Philox stream, same inputs


def test_shifted_posterior_mean():
    """f_mu = f + J (mu - theta*) when the posterior mean was moved (preds/predictives.py:58)."""
    from preds_b200 import Laplace, CategoricalLh
    from preds_b200.predictives import glm_moments
    model, omodel = _pair()
    X, y = _data(30, 5, 3)
    post = port.infer(omodel, 1.0, port.CategoricalLh(), [(X, y)], 'full')
    lap = Laplace(model, 1.0, CategoricalLh())
    lap.infer([(X, y)], cov_type='full')
    shift = 0.05 * torch.randn(lap.P, generator=torch.Generator().manual_seed(2))
    post['mu'] = post['mu'] + shift
    lap.mu = lap.mu + shift.cuda()
    Xte, _ = _data(6, 5, 3, seed=16)
    fmu_o, fvar_o = port.glm_moments(omodel, Xte, post)
    fmu, fvar = glm_moments(Xte, model, lap.mu, lap.Sigma)
    assert cases.rel_err(fmu.cpu(), fmu_o) < 1e-5
    assert cases.rel_err(fvar.cpu(), fvar_o) < TOL


def test_other_layer_types():
    """Sigmoid, AvgPool2d, a stride-2 'same' convolution and ReLU without bias: Jacobians vs autograd."""
    from preds_b200 import Jacobians
    from preds_b200.models import tfconv2d
    torch.manual_seed(5)
    net = nn.Sequential()
    net.add_module('conv1', tfconv2d(2, 4, 3, stride=(2, 2), tf_padding_type='same', bias=False))
    net.add_module('act1', nn.Sigmoid())
    net.add_module('avg', nn.AvgPool2d(kernel_size=(2, 2)))
    net.add_module('flatten', nn.Flatten())
    net.add_module('dense1', nn.Linear(4 * 2 * 2, 6, bias=False))
    net.add_module('relu', nn.ReLU())
    net.add_module('dense2', nn.Linear(6, 2))
    net.output_size = 2
    X = torch.randn(5, 2, 7, 7)
    out = net(X)
    P = sum(p.numel() for p in net.parameters())
    naive = torch.empty(5, P, 2)
    for i in range(5):
        for k in range(2):
            g = torch.autograd.grad(out[i, k], list(net.parameters()), retain_graph=True)
            naive[i, :, k] = torch.cat([t.reshape(-1) for t in g])
    Js, f = Jacobians(net.cuda(), X)
    assert torch.allclose(f.cpu(), out.detach(), atol=1e-5)
    assert float((Js.cpu() - naive).abs().max()) < 1e-5


def test_unsupported_layers_fail_loudly():
    from preds_b200 import Jacobians
    from preds_b200._cabi import BnnpError
    net = nn.Sequential(nn.Linear(4, 4), nn.GELU(), nn.Linear(4, 2)).cuda()
    net.output_size = 2
    with pytest.raises(BnnpError, match='outside the supported zoo'):
        Jacobians(net, torch.randn(3, 4))


def test_full_covariance_too_large_fails_with_a_clear_message():
    """[P,P] for the cfg-3 MLP (P = 1.49M) is 8.9 TB: refuse up front instead of dying inside an allocation."""
    from preds_b200 import MLPS, CategoricalLh, Laplace, _cabi as cabi
    model = MLPS(784, [1024, 512, 256, 128], 10, flatten=True).cuda()
    lap = Laplace(model, 1.0, CategoricalLh())
    X, y = torch.randn(4, 1, 28, 28).cuda(), torch.randint(10, (4,)).cuda()
    with pytest.raises(cabi.BnnpError, match="needs about"):
        lap.infer([(X, y)], cov_type='full')


def test_back_to_back_posteriors_do_not_share_the_sigma_split():
    """Two posteriors of equal P >= 1024 inferred one after the other (a prior-precision sweep, as the reference
    experiments do): the bf16 hi/lo copies of Sigma that the TMA-fed GLM kernel keeps live ON the covariance tensor, so
    a new Sigma that the caching allocator places at the freed one's address can never be served the old split."""
    from preds_b200 import Laplace, CategoricalLh
    from preds_b200.predictives import glm_moments, FULL_TMA_MIN_P
    model, omodel = _pair(din=40, hid=(24,), K=3)                  # P = 40*24+24 + 24*3+3 = 1059
    assert sum(p.numel() for p in model.parameters()) >= FULL_TMA_MIN_P
    X, y = _data(300, 40, 3)
    Xte, _ = _data(130, 40, 3, seed=16)
    addrs, outs = [], []
    for prior in (0.3, 30.0):
        post = port.infer(omodel, prior, port.CategoricalLh(), [(X, y)], 'full')
        _, fvar_ref = port.glm_moments(omodel, Xte, post)
        lap = Laplace(model, prior, CategoricalLh())
        lap.infer([(X, y)], cov_type='full')
        addrs.append(lap.Sigma.data_ptr())
        _, fvar = glm_moments(Xte, model, lap.mu, lap.Sigma)
        assert cases.rel_err(fvar.cpu(), fvar_ref) < TOL, prior
        outs.append(fvar.cpu())
        del lap                                                     # frees Sigma: the next one may land on its address
    assert cases.rel_err(outs[0], outs[1]) > 0.5                    # the two posteriors really differ
    # an in-place update of Sigma invalidates the split as well
    lap = Laplace(model, 0.3, CategoricalLh())
    lap.infer([(X, y)], cov_type='full')
    _, a = glm_moments(Xte, model, lap.mu, lap.Sigma)
    lap.Sigma.mul_(4.0)
    _, b = glm_moments(Xte, model, lap.mu, lap.Sigma)
    assert cases.rel_err(b.cpu(), 4.0 * a.cpu()) < 1e-5


def test_scalar_likelihoods_are_elementwise_on_any_shape():
    """BernoulliLh / GaussianLh act on one output: inv_link / Hessian / residual are elementwise on [S, N] samples of a
    model that flattens its output (preds/likelihoods.py:61-75 are plain sigmoid / p(1-p))."""
    from preds_b200 import BernoulliLh, GaussianLh
    g = torch.Generator().manual_seed(2)
    fs = torch.randn(7, 33, generator=g).cuda()                     # [S, N]
    ys = (torch.rand(7, 33, generator=g) > 0.5).float().cuda()
    lh = BernoulliLh()
    p = torch.sigmoid(fs)
    assert torch.allclose(lh.inv_link(fs), p, atol=1e-6) and lh.inv_link(fs).shape == fs.shape
    assert torch.allclose(lh.Hessian(fs), (p * (1 - p)).clamp(1e-7, 1 - 1e-7), atol=1e-6)
    assert torch.allclose(lh.residual(ys, fs), ys - p, atol=1e-6)
    gl = GaussianLh(0.5)
    assert torch.equal(gl.inv_link(fs), fs)
    assert torch.allclose(gl.residual(ys, fs), (ys - fs) / 0.25, atol=1e-5)


def test_descriptor_verification_catches_small_functional_changes():
    """A model that rescales between its leaf modules by 0.1 % computes something the layer trace cannot represent; the
    build-time probe (now at fp32-forward accuracy, 2e-4 of the output scale, eight probe inputs) must refuse it."""
    from preds_b200 import Jacobians
    from preds_b200._cabi import BnnpError

    class Scaled(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.act, self.b = nn.Linear(6, 9), nn.Tanh(), nn.Linear(9, 3)
            self.output_size = 3

        def forward(self, x):
            return self.b(self.act(self.a(x)) * 1.001)

    torch.manual_seed(0)
    with pytest.raises(BnnpError):
        Jacobians(Scaled().cuda(), torch.randn(4, 6).cuda())


def test_smooth_conv_net_on_the_tensor_core_branches_against_the_oracle():
    """A conv net WITHOUT discrete decisions (tanh, no pooling), so that diag / KFLR agree with the oracle to 1e-4 with no
    routing injected, sized to take the tensor-core conv branches with shapes CIFAR10Net does not have: C_out = 32 and 20
    (three rows r of an example per 128-row tile of the diag product since V = 3), E = 27 / 288 (ragged 128-wide e blocks),
    L = 100 / 64, 12 x 12 inputs (hoisted-index unfold passes on odd sizes), K = 3."""
    import copy
    from preds_b200 import Laplace, CategoricalLh
    torch.manual_seed(21)
    net = nn.Sequential(nn.Conv2d(3, 32, 3), nn.Tanh(), nn.Conv2d(32, 20, 3), nn.Tanh(), nn.Flatten(), nn.Linear(20 * 8 * 8, 3))
    net.output_size = 3
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(0.5)
    onet = copy.deepcopy(net)
    g = torch.Generator().manual_seed(22)
    X, y = torch.randn(192, 3, 12, 12, generator=g), torch.randint(3, (192,), generator=g)
    batches = [(X[:100], y[:100]), (X[100:], y[100:])]
    net = net.cuda()
    for cov in ('diag', 'kron'):
        post = port.infer(onet, 0.9, port.CategoricalLh(), batches, cov)
        lap = Laplace(net, 0.9, CategoricalLh())
        lap.infer(batches, cov_type=cov)
        if cov == 'diag':
            got, want = lap.precision.cpu(), post['precision']
            off = 0
            for p in onet.parameters():            # per parameter tensor, relative to its own scale
                a, b = got[off:off + p.numel()], want[off:off + p.numel()]
                assert cases.rel_err(a, b) < TOL, tuple(p.shape)
                off += p.numel()
        else:
            for qa, qb in zip(lap.Sigma_chol.qhs, post['kron'].qhs):
                for a, b in zip(qa, qb):
                    assert cases.rel_err(a.cpu(), b) < TOL, tuple(b.shape)


def test_implicit_conv_backward_on_ragged_shapes_against_autograd():
    """Backward-data of the stride-1 convs on the implicit-GEMM kernel (csrc/conv_bwd_tc.cu) with shapes that leave every
    tile ragged: 9 x 11 maps (boxes of 16 x 8 pixels), C_in = 5 / 7 (MMA N padded to 16), C_out = 7 / 4 (< one 32-channel
    chunk), a 3 x 2 filter, asymmetric ZeroPad2d padding folded into the conv, a 1 x 1 conv; per-output autograd in fp32 on
    the CPU is the reference (tests/test_jacobians.py:60-66 of the reference), and the product + col2im route
    (option conv_bwd_explicit) must give the same Jacobians."""
    from preds_b200 import Jacobians
    from preds_b200._cabi import lib
    torch.manual_seed(33)
    net = nn.Sequential(nn.Conv2d(3, 5, 3, padding=1), nn.Tanh(), nn.ZeroPad2d((2, 0, 1, 3)), nn.Conv2d(5, 7, (3, 2)), nn.Tanh(),
                        nn.Conv2d(7, 4, 1), nn.Tanh(), nn.Flatten(), nn.Linear(4 * 11 * 12, 3))
    net.output_size = 3
    X = torch.randn(5, 3, 9, 11)
    out = net(X)
    params = list(net.parameters())
    naive = torch.empty(5, sum(p.numel() for p in params), 3)
    for i in range(5):
        for k in range(3):
            g = torch.autograd.grad(out[i, k], params, retain_graph=True)
            naive[i, :, k] = torch.cat([t.reshape(-1) for t in g])
    net = net.cuda()
    Js, f = Jacobians(net, X)
    assert torch.allclose(out.detach(), f.cpu(), atol=1e-5)
    assert torch.abs(Js.cpu() - naive).max() < 5e-6 * max(1.0, float(naive.abs().max()))
    assert lib.bnnp_set_option(b'conv_bwd_explicit', 1) == 0
    try:
        Js2, _ = Jacobians(net, X)
    finally:
        lib.bnnp_set_option(b'conv_bwd_explicit', 0)
    assert torch.abs(Js2 - Js).max() < 5e-6 * max(1.0, float(naive.abs().max()))


def test_implicit_conv_forward_on_ragged_shapes_against_torch():
    """The forward convolution on the implicit-GEMM kernel (conv_bwd_tc.cu with the channel roles swapped and the taps
    mirrored): same ragged network as above (asymmetric folded padding, 3 x 2 and 1 x 1 filters, 17 / 21 input channels on
    the implicit route (>= 16; the 3-channel first conv stays on the unfolded-input GEMM), biases), at N = 4096 so that every conv is
    above the tensor-core threshold; torch's fp32 CPU forward is the reference,
    and the unfolded-input GEMM route (option conv_bwd_explicit) must agree."""
    from preds_b200._engine import net_for
    from preds_b200._cabi import lib
    torch.manual_seed(34)
    net = nn.Sequential(nn.Conv2d(3, 17, 3, padding=1), nn.Tanh(), nn.ZeroPad2d((2, 0, 1, 3)), nn.Conv2d(17, 21, (3, 2)), nn.Tanh(),
                        nn.Conv2d(21, 4, 1), nn.Tanh(), nn.Flatten(), nn.Linear(4 * 11 * 12, 3))
    net.output_size = 3
    X = torch.randn(4096, 3, 9, 11)
    with torch.no_grad():
        want = net(X)
    net = net.cuda()
    eng = net_for(net)
    got = {}
    for opt in (0, 1):
        assert lib.bnnp_set_option(b'conv_bwd_explicit', opt) == 0
        try:
            _, _, f = eng.forward(eng.prepare(X), 3)
            got[opt] = f.cpu()
        finally:
            lib.bnnp_set_option(b'conv_bwd_explicit', 0)
    scale = float(want.abs().max())
    assert float((got[0] - want).abs().max()) < 1e-5 * scale
    assert float((got[1] - want).abs().max()) < 1e-5 * scale


def test_first_conv_gradient_planes_match_the_fp32_route():
    """Diagonal GGN of CIFAR10Net (conv -> relu -> stride-2 max-pool first): by default the pool backward leaves the first conv's
    output gradient as the bf16 hi / lo planes of the conv-weight diagonal kernel and the bias position sums come from the
    pooled gradient (BNNP_BWD_FIRST_CONV_PLANES); with option conv_bwd_explicit the fp32 gradient + split pass + product /
    col2im conv route of the round-2 start runs.  Same diagonal per parameter tensor (2e-4 of its own scale; the two routes
    differ in fp32 summation order only), and the flag is really taken."""
    from preds_b200 import Laplace, CategoricalLh, CIFAR10Net
    from preds_b200._cabi import lib
    from preds_b200._engine import net_for
    torch.manual_seed(5)
    model = CIFAR10Net(3, 10, use_tanh=True).cuda()
    g = torch.Generator().manual_seed(6)
    X, y = torch.randn(64, 3, 32, 32, generator=g).cuda(), torch.randint(10, (64,), generator=g).cuda()
    eng = net_for(model)
    eng.prepare(X)
    assert lib.bnnp_net_first_conv_planes_ok(eng.ops, eng.n_ops, 64, 10) == 1
    got = {}
    for opt in (0, 1):
        assert lib.bnnp_set_option(b'conv_bwd_explicit', opt) == 0
        try:
            assert lib.bnnp_net_first_conv_planes_ok(eng.ops, eng.n_ops, 64, 10) == 1 - opt
            lap = Laplace(model, 1.0, CategoricalLh())
            lap.infer([(X, y)], cov_type='diag')
            got[opt] = lap.precision.clone()
        finally:
            lib.bnnp_set_option(b'conv_bwd_explicit', 0)
    off = 0
    for p in model.parameters():
        a, b = got[0][off:off + p.numel()], got[1][off:off + p.numel()]
        assert float((a - b).abs().max()) <= 2e-4 * float(b.abs().max()), tuple(p.shape)
        off += p.numel()
