"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/bnnp_b200.h declares, validates arguments, and the host-side workspace planner and the
Python glue behave like the reference's API (no GPU compute is launched here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'bnnp_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(bnnp_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from preds_b200 import _cabi
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(_cabi.lib, name), 'missing export: ' + name
    # and the Python binding knows the prototype of each of them
    assert sorted(_cabi.EXPORTED) == declared
    assert _cabi.lib.bnnp_version() >= 100
    assert _cabi.lib.bnnp_strerror(-1) == b'invalid argument'


def _mlp_ops(sizes, bias=True):
    from preds_b200 import _cabi
    ops, off = [], 0
    for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:])):
        op = _cabi.Op()
        op.kind, op.in_c, op.in_h, op.in_w = _cabi.OP_LINEAR, a, 1, 1
        op.out_c, op.out_h, op.out_w = b, 1, 1
        op.kh = op.kw = op.stride_h = op.stride_w = 1
        op.has_bias = int(bias)
        op.weight = 1 << 20       # never dereferenced by the planner
        op.theta_off_w = off
        off += a * b
        op.theta_off_b = off if bias else -1
        off += b if bias else 0
        ops.append(op)
        if i < len(sizes) - 2:
            t = _cabi.Op()
            t.kind, t.in_c, t.in_h, t.in_w, t.out_c, t.out_h, t.out_w = _cabi.OP_TANH, b, 1, 1, b, 1, 1
            ops.append(t)
    return (_cabi.Op * len(ops))(*ops), off


def test_net_plan_layout():
    from preds_b200 import _cabi
    ops, P = _mlp_ops([784, 75, 10])
    plan = _cabi.Plan()
    assert _cabi.lib.bnnp_net_plan(ops, len(ops), 64, 10, C.byref(plan)) == 0
    assert plan.P == P == 59635 and plan.K == 10
    assert plan.Dtot == 788 + 76 and plan.Mtot == 76 + 12 and plan.n_linear == 2   # slices padded to 4
    # layer 1 writes straight into layer 2's slice of A_cat; tanh is in place; logits are standalone
    assert plan.op[0].act_off == plan.acat_off + plan.op[2].lin_col and plan.op[0].act_ld == plan.Dtot
    assert plan.op[1].act_off == plan.op[0].act_off
    assert plan.op[2].act_ld == 10
    assert plan.op[0].grad_off == plan.gcat_off and plan.op[2].grad_off == plan.gcat_off + 76
    assert plan.total_floats >= 64 * plan.Dtot + 640 * plan.Mtot + 640


def test_net_plan_rejects_bad_arguments():
    from preds_b200 import _cabi
    ops, _ = _mlp_ops([4, 3, 2])
    plan = _cabi.Plan()
    assert _cabi.lib.bnnp_net_plan(ops, len(ops), 0, 2, C.byref(plan)) == _cabi.ERR_INVALID
    ops[2].in_c = 5                       # shape mismatch between consecutive ops
    assert _cabi.lib.bnnp_net_plan(ops, len(ops), 8, 2, C.byref(plan)) == _cabi.ERR_INVALID
    with pytest.raises(ValueError):
        _cabi.check(_cabi.ERR_INVALID, 'x')
    with pytest.raises(_cabi.BnnpError):
        _cabi.check(-2, 'x')


def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device."""
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from preds_b200 import Laplace, CategoricalLh, MLPS, Jacobians
    from preds_b200._cabi import BnnpError
    model = MLPS(2, [4], 2)
    lap = Laplace(model, 1.0, CategoricalLh())
    with pytest.raises(BnnpError):
        lap.infer([(torch.randn(4, 2), torch.zeros(4).long())], cov_type='full')
    with pytest.raises(BnnpError):
        Jacobians(model, torch.randn(4, 2))


def test_laplace_host_logic():
    """Argument handling mirrors preds/laplace.py:84-86,95-96,115-135 and preds/kron.py:45-56."""
    from preds_b200 import Laplace, CategoricalLh, GaussianLh, BernoulliLh, MLPS, Kron
    model = MLPS(3, [4], 2)
    lap = Laplace(model, 0.5, CategoricalLh())
    assert lap.P == 3 * 4 + 4 + 4 * 2 + 2 and not lap.inferred
    with pytest.raises(ValueError, match='Need to infer first'):
        lap.predictive_samples_glm(torch.randn(2, 3))
    with pytest.raises(ValueError, match='Need to infer first'):
        lap.predictive_samples_bnn(torch.randn(2, 3))
    assert torch.equal(lap.expand_prior_precision('diag'), torch.full((lap.P,), 0.5))
    assert torch.equal(lap.expand_prior_precision('full'), torch.diag(torch.full((lap.P,), 0.5)))
    vec = torch.rand(lap.P)
    lap.prior_prec = vec
    assert torch.equal(lap.expand_prior_precision('full'), torch.diag(vec))
    lap.prior_prec = torch.eye(lap.P)
    with pytest.raises(ValueError, match='Will not diagonalize'):
        lap.expand_prior_precision('diag')
    lap.prior_prec = torch.zeros(2, 2, 2)
    with pytest.raises(ValueError, match='Invalid shape'):
        lap.expand_prior_precision('full')
    # Kron bookkeeping on the CPU (factor shapes as p.kflr; delta broadcasting)
    kron = Kron(model, 0.3)
    assert [[tuple(m.shape) for m in qh] for qh in kron.qhs] == [[(4, 4), (3, 3)], [(4, 4)], [(2, 2), (4, 4)], [(2, 2)]]
    assert len(kron.deltas) == 4 and float(kron.deltas[2]) == pytest.approx(0.3)
    with pytest.raises(ValueError, match='Deltas do not match'):
        kron.change_delta(torch.ones(3))
    with pytest.raises(ValueError, match='Invalid dimensionality'):
        kron.change_delta(torch.ones(2, 2))
    with pytest.raises(ValueError):
        kron.update([[torch.zeros(4, 4)] * 3] * 4)
    kron.update([[torch.ones_like(m) for m in qh] for qh in kron.qhs], factor=2.0, batch_factor=0.5)
    assert float(kron.qhs[0][0][0, 0]) == 2.0 and float(kron.qhs[0][1][0, 0]) == 0.5
    # likelihood conventions (tests/test_likelihood.py:62-75,99-110)
    assert GaussianLh(2.0).nn_loss()[1] == pytest.approx(1 / 8)
    assert CategoricalLh().nn_loss()[1] == 1
    with pytest.raises(ValueError):
        BernoulliLh().nn_loss()
    with pytest.raises(ValueError, match='invalid activation'):
        MLPS(2, [3], 2, activation='gelu')


def test_model_zoo_matches_reference_shapes():
    from preds_b200 import MLPS, CIFAR10Net, CIFAR100Net
    assert sum(p.numel() for p in MLPS(2, [50, 50], 2).parameters()) == 2802
    assert sum(p.numel() for p in MLPS(784, [75], 10, flatten=True).parameters()) == 59635
    net = CIFAR10Net(3, 10, use_tanh=True)
    assert sum(p.numel() for p in net.parameters()) == 895210
    assert net(torch.zeros(2, 3, 32, 32)).shape == (2, 10)
    assert CIFAR10Net(1, 10)(torch.zeros(1, 1, 28, 28)).shape == (1, 10)
    assert CIFAR100Net(3, 100)(torch.zeros(1, 3, 32, 32)).shape == (1, 100)
    # same init stream as the reference constructors (seed 71 fixture of tests/test_jacobians.py)
    from tests import cases
    g = cases.load_golden('cifar10net_probe')
    torch.manual_seed(71)
    net = CIFAR10Net(in_channels=2, n_out=3)
    theta = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    assert np.array_equal(theta[torch.from_numpy(g['theta_idx'])].numpy(), g['theta_at_idx'])


def test_shard_bounds():
    from preds_b200.parallel import shard_bounds
    for n, w in [(10, 3), (7, 8), (50000, 8), (1, 1)]:
        blocks = [shard_bounds(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _conv_pool_ops(in_c=3, hw=32, out_c=64, k=5, pool_k=3, pool_s=2, relu=True):
    """conv (valid) [-> relu] -> max-pool ('same' zero padding) -> flatten -> linear, as CIFAR10Net's first stage."""
    from preds_b200 import _cabi
    oh = hw - k + 1
    ph = -(-oh // pool_s)
    pad = max((ph - 1) * pool_s + pool_k - oh, 0)
    ops = []
    c = _cabi.Op()
    c.kind, c.in_c, c.in_h, c.in_w, c.out_c, c.out_h, c.out_w = _cabi.OP_CONV2D, in_c, hw, hw, out_c, oh, oh
    c.kh = c.kw = k
    c.stride_h = c.stride_w = 1
    c.has_bias, c.weight, c.bias = 1, 1 << 20, 1 << 21
    c.theta_off_w, c.theta_off_b = 0, out_c * in_c * k * k
    ops.append(c)
    if relu:
        r = _cabi.Op()
        r.kind, r.in_c, r.in_h, r.in_w, r.out_c, r.out_h, r.out_w = _cabi.OP_RELU, out_c, oh, oh, out_c, oh, oh
        ops.append(r)
    p = _cabi.Op()
    p.kind, p.in_c, p.in_h, p.in_w, p.out_c, p.out_h, p.out_w = _cabi.OP_MAXPOOL2D, out_c, oh, oh, out_c, ph, ph
    p.kh = p.kw = pool_k
    p.stride_h = p.stride_w = pool_s
    p.pad_l, p.pad_r = pad // 2, pad - pad // 2
    p.pad_t, p.pad_b = pad // 2, pad - pad // 2
    ops.append(p)
    f = _cabi.Op()
    f.kind, f.in_c, f.in_h, f.in_w, f.out_c, f.out_h, f.out_w = _cabi.OP_FLATTEN, out_c, ph, ph, out_c, ph, ph
    ops.append(f)
    return (_cabi.Op * len(ops))(*ops)


def test_first_conv_shortcuts_are_offered_only_where_they_apply():
    """Host logic of the two first-conv shortcuts of the factor passes (no device work): BNNP_BWD_FIRST_CONV_POSSUM (KFAC) needs
    conv [-> activation] -> max-pool with windows <= 3 x 3; BNNP_BWD_FIRST_CONV_PLANES (diag) additionally a stride >= 2 pool
    on maps whose width is a multiple of 4 and whose size is a multiple of 8, a batch large enough for the TMA-fed diagonal
    kernel, and none of the routing options that move the conv paths elsewhere."""
    from preds_b200 import _cabi
    lib = _cabi.lib
    ops = _conv_pool_ops()                                   # CIFAR10Net's first stage: 28 x 28 maps, 3 x 3 / stride-2 pool
    assert lib.bnnp_net_first_conv_possum_ok(ops, len(ops)) == 1
    assert lib.bnnp_net_first_conv_planes_ok(ops, len(ops), 256, 10) == 1
    assert lib.bnnp_net_first_conv_planes_ok(ops, len(ops), 1, 1) == 0          # below the tensor-core threshold
    assert lib.bnnp_net_first_conv_planes_ok(ops, len(ops), 0, 10) == 0
    for name in (b'conv_bwd_explicit', b'net_simt', b'gemm_generated'):
        assert lib.bnnp_set_option(name, 1) == 0
        try:
            assert lib.bnnp_net_first_conv_planes_ok(ops, len(ops), 256, 10) == 0, name
        finally:
            lib.bnnp_set_option(name, 0)
    assert lib.bnnp_net_first_conv_planes_ok(ops, len(ops), 256, 10) == 1
    ops = _conv_pool_ops(hw=30)                              # 26 x 26 maps: width not a multiple of 4
    assert lib.bnnp_net_first_conv_possum_ok(ops, len(ops)) == 1
    assert lib.bnnp_net_first_conv_planes_ok(ops, len(ops), 256, 10) == 0
    ops = _conv_pool_ops(pool_s=1)                           # stride-1 pool: four-wide pool backward does not apply
    assert lib.bnnp_net_first_conv_planes_ok(ops, len(ops), 256, 10) == 0
    ops = _conv_pool_ops(pool_k=5)                           # 5 x 5 window: neither shortcut
    assert lib.bnnp_net_first_conv_possum_ok(ops, len(ops)) == 0
    assert lib.bnnp_net_first_conv_planes_ok(ops, len(ops), 256, 10) == 0
    ops = _conv_pool_ops(relu=False)
    assert lib.bnnp_net_first_conv_possum_ok(ops, len(ops)) == 1
    assert lib.bnnp_net_first_conv_planes_ok(ops, len(ops), 256, 10) == 1
    mlp, _ = _mlp_ops([784, 75, 10])
    assert lib.bnnp_net_first_conv_possum_ok(mlp, len(mlp)) == 0
    assert lib.bnnp_net_first_conv_planes_ok(mlp, len(mlp), 4096, 10) == 0
