"""Host-side plumbing between ``nn.Module`` models and the C ABI: builds the network descriptor
(``bnnp_op_t[]``), owns the torch-allocated workspace, and wraps each kernel entry point.

No arithmetic of the hot path happens here; torch provides device memory and the stream only.
"""
import ctypes as C
import weakref

import torch
import torch.nn as nn

from . import _cabi as cabi
from ._cabi import lib, check, ptr, stream

_ACT = {nn.ReLU: cabi.OP_RELU, nn.Tanh: cabi.OP_TANH, nn.Sigmoid: cabi.OP_SIGMOID}
_IGNORED = (nn.Identity, nn.Dropout)        # model.eval() semantics (preds/laplace.py:35)


def _pair(x):
    return (x, x) if isinstance(x, int) else tuple(x)


def _leaf_trace(model, sample):
    """Run one example through the model recording (leaf module, in shape, out shape, padding)
    in execution order.  deepobs-style ``ZeroPad2d`` paddings are set by forward-pre-hooks
    (preds/models.py:250-266), so they are only known after a dry run."""
    trace, handles = [], []

    def hook(mod, inp, out):
        pad = tuple(_pair4(mod.padding)) if isinstance(mod, nn.ZeroPad2d) else None
        trace.append((mod, tuple(inp[0].shape[1:]), tuple(out.shape[1:]), pad))

    for m in model.modules():
        if len(list(m.children())) == 0:
            handles.append(m.register_forward_hook(hook))
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            model(sample)
    finally:
        for h in handles:
            h.remove()
        model.train(was_training)
    return trace


_FUNCTIONAL_ACTS = {torch.relu: nn.ReLU, torch.tanh: nn.Tanh, torch.sigmoid: nn.Sigmoid}


def _functional_mlp(model):
    """``preds.models.MLP`` / ``SiMLP`` (preds/models.py:124-176) apply their activation as a plain function,
    invisible to a leaf-module trace.  Recognise that layout by its attributes (``hidden_layers``,
    ``output_layer``, ``act``) and synthesise the trace: Linear, act, ..., Linear."""
    if not (hasattr(model, 'hidden_layers') and hasattr(model, 'output_layer') and hasattr(model, 'act')):
        return None
    layers = list(model.hidden_layers) + [model.output_layer]
    if not all(isinstance(l, nn.Linear) for l in layers):
        return None
    if model.act not in _FUNCTIONAL_ACTS:
        raise cabi.BnnpError('functional activation %r is outside the supported zoo (relu, tanh, sigmoid)'
                             % (getattr(model.act, '__name__', model.act),))
    act = _FUNCTIONAL_ACTS[model.act]
    trace = []
    for l in layers[:-1]:
        trace.append((l, (l.in_features,), (l.out_features,), None))
        trace.append((act(), (l.out_features,), (l.out_features,), None))
    trace.append((layers[-1], (layers[-1].in_features,), (layers[-1].out_features,), None))
    return trace


def _pair4(p):
    return (p, p, p, p) if isinstance(p, int) else tuple(p)


def _shape3(shape):
    if len(shape) == 3:
        return shape
    if len(shape) == 1:
        return (shape[0], 1, 1)
    raise cabi.BnnpError('unsupported activation shape %s' % (shape,))


class Net:
    """Descriptor + workspace for one model on one device."""

    def __init__(self, model):
        self._model = weakref.ref(model)    # the cache must not keep the model alive
        self.params = list(model.parameters())
        if not self.params:
            raise ValueError('model has no parameters')
        self.device = self.params[0].device
        cabi.require_device(self.params[0])
        self.P = sum(p.numel() for p in self.params)
        self._sig = None
        self._verified = False
        self.flat_output = False
        self._plans = {}
        self._ws = None
        self._scratch = None
        # descriptor: built by the first prepare(); a rank that owns no batches never gets there
        self.ops, self.mods, self.n_ops = None, [], 0

    # ------------------------------------------------------------------ descriptor
    def _build(self, in_shape):
        model = self._model()
        sample = torch.zeros((1,) + tuple(in_shape), device=self.device)
        trace = _functional_mlp(model) or _leaf_trace(model, sample)
        offs, off = {}, 0
        for p in self.params:
            offs[id(p)] = off
            off += p.numel()
        ops, mods = [], []
        pend_pad = (0, 0, 0, 0)
        for mod, ishape, oshape, pad in trace:
            if isinstance(mod, _IGNORED):
                continue
            if isinstance(mod, nn.ZeroPad2d):
                pend_pad = tuple(a + b for a, b in zip(pend_pad, pad))
                continue
            op = cabi.Op()
            if isinstance(mod, nn.ZeroPad2d):
                continue
            # input shape as seen by the op = shape before any folded padding
            if any(pend_pad):
                l, r, t, b = pend_pad
                ishape = (ishape[0], ishape[1] - t - b, ishape[2] - l - r)
            if isinstance(mod, nn.Linear):
                if any(pend_pad):
                    raise cabi.BnnpError('ZeroPad2d in front of Linear is unsupported')
                if len(ishape) != 1:
                    raise cabi.BnnpError('Linear expects a flattened input (got %s)' % (ishape,))
                op.kind = cabi.OP_LINEAR
                op.in_c, op.in_h, op.in_w = mod.in_features, 1, 1
                op.out_c, op.out_h, op.out_w = mod.out_features, 1, 1
                op.kh = op.kw = op.stride_h = op.stride_w = 1
            elif isinstance(mod, nn.Conv2d):
                if _pair(mod.dilation) != (1, 1) or mod.groups != 1 or mod.padding_mode != 'zeros':
                    raise cabi.BnnpError('Conv2d with dilation/groups/non-zero padding mode is unsupported')
                ph, pw = _pair(mod.padding)
                op.kind = cabi.OP_CONV2D
                op.in_c, op.in_h, op.in_w = ishape
                op.out_c, op.out_h, op.out_w = oshape
                op.kh, op.kw = _pair(mod.kernel_size)
                op.stride_h, op.stride_w = _pair(mod.stride)
                pend_pad = (pend_pad[0] + pw, pend_pad[1] + pw, pend_pad[2] + ph, pend_pad[3] + ph)
            elif isinstance(mod, (nn.MaxPool2d, nn.AvgPool2d)):
                is_max = isinstance(mod, nn.MaxPool2d)
                if is_max and (_pair(mod.dilation) != (1, 1) or mod.ceil_mode):
                    raise cabi.BnnpError('MaxPool2d with dilation/ceil_mode is unsupported')
                if not is_max and (mod.ceil_mode or _pair(mod.padding) != (0, 0)):
                    raise cabi.BnnpError('AvgPool2d with padding/ceil_mode is unsupported')
                if is_max and _pair(mod.padding) != (0, 0):
                    # torch pads max-pool with -inf, not zeros: not expressible as zero padding
                    raise cabi.BnnpError('MaxPool2d(padding>0) is unsupported; use ZeroPad2d + padding=0')
                op.kind = cabi.OP_MAXPOOL2D if is_max else cabi.OP_AVGPOOL2D
                op.in_c, op.in_h, op.in_w = ishape
                op.out_c, op.out_h, op.out_w = oshape
                op.kh, op.kw = _pair(mod.kernel_size)
                st = mod.stride if mod.stride is not None else mod.kernel_size
                op.stride_h, op.stride_w = _pair(st)
            elif type(mod) in _ACT:
                if any(pend_pad):
                    raise cabi.BnnpError('ZeroPad2d in front of an activation is unsupported')
                op.kind = _ACT[type(mod)]
                op.in_c, op.in_h, op.in_w = _shape3(ishape)
                op.out_c, op.out_h, op.out_w = _shape3(oshape)
            elif isinstance(mod, nn.Flatten):
                op.kind = cabi.OP_FLATTEN
                op.in_c, op.in_h, op.in_w = _shape3(ishape)
                n = 1
                for s in oshape:
                    n *= s
                op.out_c, op.out_h, op.out_w = n, 1, 1
            else:
                raise cabi.BnnpError('layer type %s is outside the supported zoo (Linear, Conv2d, ReLU, '
                                     'Tanh, Sigmoid, MaxPool2d, AvgPool2d, ZeroPad2d, Flatten)'
                                     % type(mod).__name__)
            if op.kind in (cabi.OP_CONV2D, cabi.OP_MAXPOOL2D, cabi.OP_AVGPOOL2D):
                op.pad_l, op.pad_r, op.pad_t, op.pad_b = pend_pad
            pend_pad = (0, 0, 0, 0)
            if op.kind in (cabi.OP_LINEAR, cabi.OP_CONV2D):
                op.has_bias = int(mod.bias is not None)
                op.theta_off_w = offs[id(mod.weight)]
                op.theta_off_b = offs[id(mod.bias)] if mod.bias is not None else -1
            ops.append(op)
            mods.append(mod)
        if len(ops) > cabi.MAX_OPS:
            raise cabi.BnnpError('too many layers (%d > %d)' % (len(ops), cabi.MAX_OPS))
        if not ops or ops[0].kind == cabi.OP_FLATTEN:
            # a leading Flatten only reshapes the caller's tensor
            while ops and ops[0].kind == cabi.OP_FLATTEN:
                ops.pop(0)
                mods.pop(0)
        covered = sum((m.weight.numel() + (m.bias.numel() if m.bias is not None else 0))
                      for m in mods if isinstance(m, (nn.Linear, nn.Conv2d)))
        if covered != self.P:
            raise cabi.BnnpError('model has parameters outside Linear/Conv2d layers (%d of %d covered)'
                                 % (covered, self.P))
        self.ops = (cabi.Op * len(ops))(*ops)
        self.mods = mods
        self.n_ops = len(ops)
        self.in_size = ops[0].in_c * ops[0].in_h * ops[0].in_w
        self.K = ops[-1].out_c * ops[-1].out_h * ops[-1].out_w
        self.linear_only = all(o.kind not in (cabi.OP_CONV2D, cabi.OP_MAXPOOL2D, cabi.OP_AVGPOOL2D)
                               for o in ops)
        self._plans = {}
        self._sig = tuple(in_shape)
        self._verified = False

    def _verify(self, in_shape):
        """The descriptor must compute what ``model(x)`` computes: anything applied outside the traced
        modules (functional activations, residual adds, reshapes) would otherwise give silently wrong
        Jacobians.  Eight probe inputs through both, compared at the accuracy of the fp32 forward pass; also records whether the model flattens a single output
        ([N] instead of [N,1], preds/models.py:165), which ``Jacobians`` / the predictives mirror."""
        model = self._model()
        gen = torch.Generator(device='cpu').manual_seed(1234)
        n_probe = 8
        x = (2.0 * torch.randn((n_probe,) + tuple(in_shape), generator=gen)).to(self.device)
        was_training = model.training
        model.eval()
        tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
        try:
            with torch.no_grad():
                want = model(x)
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
            model.train(was_training)
        self.flat_output = want.dim() == 1
        _, _, got = self.forward(x.reshape(n_probe, -1).contiguous(), 1)
        want = want.reshape(n_probe, -1).to(torch.float32)
        if want.shape != got.shape:
            raise cabi.BnnpError('model output shape %s does not match the traced network (%s)'
                                 % (tuple(want.shape), tuple(got.shape)))
        err = (got - want).abs().max().item()
        scale = max(want.abs().max().item(), got.abs().max().item())
        # the split-bf16 / fp32 forward agrees with torch to ~1e-5 of the outputs' own scale (TF32 convolutions in the
        # caller's torch settings would cost ~1e-3: the probe runs with them off); also catches NaN
        if not err <= 2e-4 * scale + 1e-6:
            raise cabi.BnnpError('model(x) differs from its layer trace (max |diff| %.3g): the model applies '
                                 'operations outside its leaf modules (functional activations, skip '
                                 'connections, ...), which this engine cannot represent' % err)
        self._verified = True

    def prepare(self, X):
        """Validate / flatten the input batch, (re)build the descriptor, refresh weight pointers."""
        cabi.require_device(self.params[0])
        X = cabi.f32(X, self.device)
        in_shape = tuple(X.shape[1:])
        if self._sig != in_shape:
            self._build(in_shape)
        for op, mod in zip(self.ops, self.mods):
            if op.kind in (cabi.OP_LINEAR, cabi.OP_CONV2D):
                w = mod.weight
                if w.dtype != torch.float32 or not w.is_contiguous():
                    raise cabi.BnnpError('parameters must be contiguous float32')
                op.weight = w.data_ptr()
                op.bias = mod.bias.data_ptr() if mod.bias is not None else None
        if not self._verified:
            self._verify(in_shape)
        return X.reshape(X.shape[0], -1)

    def plan(self, N, V):
        key = (N, V)
        if key not in self._plans:
            pl = cabi.Plan()
            check(lib.bnnp_net_plan(self.ops, self.n_ops, N, V, C.byref(pl)), 'bnnp_net_plan')
            self._plans[key] = pl
        return self._plans[key]

    def workspace(self, plan):
        n = int(plan.total_floats) + 64
        if self._ws is None or self._ws.numel() < n:
            self._ws = None
            self._ws = torch.empty(n, dtype=torch.float32, device=self.device)
        return self._ws

    def scratch(self, n_floats):
        n = int(n_floats) + 64
        if self._scratch is None or self._scratch.numel() < n:
            self._scratch = None
            self._scratch = torch.empty(n, dtype=torch.float32, device=self.device)
        return self._scratch

    def release(self):
        self._ws = None
        self._scratch = None

    # ------------------------------------------------------------------ passes
    def forward(self, X2d, V):
        """Forward N examples; returns (plan, ws, logits [N,K])."""
        N = X2d.shape[0]
        pl = self.plan(N, V)
        ws = self.workspace(pl)
        f = torch.empty(N, self.K, dtype=torch.float32, device=self.device)
        check(lib.bnnp_net_forward(self.ops, self.n_ops, C.byref(pl), ptr(X2d), N, ptr(ws), ptr(f),
                                   stream()), 'bnnp_net_forward')
        return pl, ws, f

    def backward(self, pl, ws, seed, N, V, flags=0):
        check(lib.bnnp_net_backward_ex(self.ops, self.n_ops, C.byref(pl), ptr(seed), N, V, ptr(ws), flags,
                                       stream()), 'bnnp_net_backward_ex')

    def identity_seed(self, N):
        seed = torch.empty(N, self.K, self.K, dtype=torch.float32, device=self.device)
        check(lib.bnnp_identity_seed(N, self.K, ptr(seed), stream()), 'bnnp_identity_seed')
        return seed

    def jacobian_factors(self, X):
        """forward + identity-seeded backward: the (a_l, g_l) factors of J for a batch."""
        X2d = self.prepare(X)
        N = X2d.shape[0]
        pl, ws, f = self.forward(X2d, self.K)
        self.backward(pl, ws, self.identity_seed(N), N, self.K)
        return X2d, pl, ws, f

    def sqrt_factors(self, X, lik_code, kfac_only=False, diag_only=False):
        """forward + backward seeded with the loss-Hessian square root (BackPACK convention).  ``kfac_only``: the
        consumer is the KFAC accumulation, which needs only position sums at a first conv layer; ``diag_only``: the
        consumer is the diagonal accumulation, which takes a first conv layer's gradient as the bf16 planes of its
        tensor-core kernel.  Either returns the flags to hand to ``bnnp_kfac_accum_ex`` / ``bnnp_ggn_diag_accum_ex`` as
        a fifth element."""
        X2d = self.prepare(X)
        N = X2d.shape[0]
        flags = 0
        if kfac_only and lib.bnnp_net_first_conv_possum_ok(self.ops, self.n_ops):
            flags = cabi.BWD_FIRST_CONV_POSSUM
        elif diag_only and lib.bnnp_net_first_conv_planes_ok(self.ops, self.n_ops, N, self.K):
            flags = cabi.BWD_FIRST_CONV_PLANES
        pl, ws, f = self.forward(X2d, self.K)
        seed = torch.empty(N, self.K, self.K, dtype=torch.float32, device=self.device)
        check(lib.bnnp_lik_sqrt_seed(lik_code, ptr(f), N, self.K, ptr(seed), stream()), 'bnnp_lik_sqrt_seed')
        self.backward(pl, ws, seed, N, self.K, flags)
        if kfac_only or diag_only:
            return X2d, pl, ws, f, flags
        return X2d, pl, ws, f

    def materialise_jacobians(self, X):
        X2d, pl, ws, f = self.jacobian_factors(X)
        N = X2d.shape[0]
        Js = torch.empty(N, self.P, self.K, dtype=torch.float32, device=self.device)
        check(lib.bnnp_jacobians(self.ops, self.n_ops, C.byref(pl), N, ptr(X2d), ptr(ws), ptr(Js), stream()),
              'bnnp_jacobians')
        return Js, f


    def output_jacobian(self, X, output):
        """(J [N,P], f [N,K]) for ONE network output: forward + a one-column backward seeded with e_output
        (preds/gradients.py:49-77 ``Jacobians_gp`` with a single output)."""
        X2d = self.prepare(X)
        N = X2d.shape[0]
        pl, ws, f = self.forward(X2d, 1)
        seed = torch.zeros(N, 1, self.K, dtype=torch.float32, device=self.device)
        seed[:, 0, int(output)] = 1.0
        self.backward(pl, ws, seed, N, 1)
        J = torch.empty(N, self.P, dtype=torch.float32, device=self.device)
        check(lib.bnnp_jacobians_cols(self.ops, self.n_ops, C.byref(pl), N, 1, ptr(X2d), ptr(ws), ptr(J), stream()),
              'bnnp_jacobians_cols')
        return J, f


_NETS = weakref.WeakKeyDictionary()        # model -> Net; entries die with the model


def net_for(model):
    """One cached ``Net`` per model object (the model keeps its own parameters; we hold pointers)."""
    try:
        ent = _NETS.get(model)
    except TypeError:                       # unhashable / non-weakrefable "model" objects: no caching
        return Net(model)
    if ent is None or ent.device != next(model.parameters()).device:
        ent = Net(model)
        _NETS[model] = ent
    return ent


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*[p if p else None for p in ptrs])
    return arr


def float_array(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])
