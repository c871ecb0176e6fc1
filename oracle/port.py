"""CPU restatement ("port") of the Laplace-GGN hot path of aleximmer/BNN-predictions.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module, and only as
the checker / the timed CPU baseline.  The product package (``bnn-predictions_b200/preds_b200``)
never imports it and has no CPU fallback.

Parity status: PINNED.  Every function below is checked (``tests/test_oracle_golden.py``)
against fixtures in ``tests/golden/`` that were produced by running the UNMODIFIED reference
(``/root/reference/preds``) through ``oracle/refharness`` with ``oracle/gen_golden.py``; in the
build container the same tests also compare live against the reference.  The one part of the
reference's arithmetic that lives in an absent third-party dependency is BackPACK 1.1.1
(``requirements.txt:4``): ``BatchGrad`` is pinned by the reference's own
``tests/test_jacobians.py:60-66`` (vs. pure autograd); ``DiagGGNExact`` by the identity
``diag == diag(full GGN)``; ``KFLR`` for Linear layers and all biases by the N=1 identity
``G (x) A == exact block``.  Conv-weight KFLR values (spatial pre-sum in G) follow the published
BackPACK ``HBPConv2d`` definition and are *unpinned by any reference test* (SURVEY.md 8c).

The callers either side of the path (``post_process``, ``diagonal_ggn``,
``linear_sampling_predictive``, ``laplace_refine`` / ``vi_refine`` / ``vi_diag_refine``, ``linear_regression_predictive``,
``classification_metrics``) are restated at the end of the file and PINNED the same way
(``tests/golden/ggn_simlp_{cls,bern}.npz``, ``extras.npz``; ``tests/test_oracle_ggn_golden.py``,
``tests/test_oracle_extras_golden.py``).

Everything is float32 torch on the CPU, written as plain functions with explicit
standard-normal draws (``eps``) so the CUDA path can be fed identical numbers.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

CLAMP_EPS = 1e-7          # preds/likelihoods.py:4


# ------------------------------------------------------------------------------------------
# models  (preds/models.py:91-122, :239-285; deepobs tfconv2d/tfmaxpool2d semantics)
# ------------------------------------------------------------------------------------------
class SamePad2d(nn.Module):
    """TF-"same" zero padding in front of a conv/pool with padding=0 (Appendix A of SURVEY.md):
    out = ceil(in/stride); pad = max((out-1)*stride + k - in, 0); left/top = pad//2."""

    def __init__(self, kernel, stride):
        super().__init__()
        self.kernel, self.stride = kernel, stride

    def pads(self, H, W):
        k, s = self.kernel, self.stride
        ph = max((math.ceil(H / s) - 1) * s + k - H, 0)
        pw = max((math.ceil(W / s) - 1) * s + k - W, 0)
        return (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)

    def forward(self, x):
        return F.pad(x, self.pads(x.shape[2], x.shape[3]))


def make_mlps(input_size, hidden_sizes, output_size, activation='tanh', flatten=False, bias=True):
    """Sequential MLP with module activations; restates preds/models.py:91-122."""
    act = {'relu': nn.ReLU, 'tanh': nn.Tanh}[activation]
    layers = [nn.Flatten()] if flatten else []
    sizes = [input_size] + list(hidden_sizes)
    for a, b in zip(sizes[:-1], sizes[1:]):
        layers += [nn.Linear(a, b, bias=bias), act()]
    layers.append(nn.Linear(sizes[-1], output_size if output_size is not None else 1, bias=bias))
    net = nn.Sequential(*layers)
    net.output_size = output_size if output_size is not None else 1
    return net


class FunctionalMLP(nn.Module):
    """MLP whose activation is a plain function and whose single output is flattened to [N];
    restates preds/models.py:124-176 (``MLP`` / ``SiMLP``)."""

    def __init__(self, input_size, hidden_sizes, output_size, activation='tanh'):
        super().__init__()
        self.input_size, self.output_size = input_size, (1 if output_size is None else output_size)
        self.act = {'relu': torch.relu, 'tanh': torch.tanh, 'sigmoid': torch.sigmoid}[activation]
        sizes = [input_size] + list(hidden_sizes)
        self.hidden_layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:])])
        self.output_layer = nn.Linear(sizes[-1], self.output_size)

    def forward(self, x):
        h = x.view(-1, self.input_size)
        for layer in self.hidden_layers:
            h = self.act(layer(h))
        z = self.output_layer(h)
        return z.flatten() if self.output_size == 1 else z


def make_simlp(input_size, output_size, n_layers, n_units, activation='tanh'):
    return FunctionalMLP(input_size, [n_units] * n_layers, output_size, activation)


def make_cifar10net(in_channels=3, n_out=10, use_tanh=False):
    """DeepOBS 3c3d network; restates preds/models.py:239-285 (flat module list)."""
    act = nn.Tanh if use_tanh else nn.ReLU
    net = nn.Sequential(
        nn.Conv2d(in_channels, 64, 5), nn.ReLU(), SamePad2d(3, 2), nn.MaxPool2d(3, 2),
        nn.Conv2d(64, 96, 3), nn.ReLU(), SamePad2d(3, 2), nn.MaxPool2d(3, 2),
        SamePad2d(3, 1), nn.Conv2d(96, 128, 3), nn.ReLU(), SamePad2d(3, 2), nn.MaxPool2d(3, 2),
        nn.Flatten(),
        nn.Linear(3 * 3 * 128, 512), act(), nn.Linear(512, 256), act(), nn.Linear(256, n_out))
    for m in net.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.constant_(m.bias, 0.0)
            nn.init.xavier_normal_(m.weight)
        if isinstance(m, nn.Linear):
            nn.init.constant_(m.bias, 0.0)
            nn.init.xavier_uniform_(m.weight)
    net.output_size = n_out
    return net


def make_cifar100net(in_channels=3, n_out=100):
    """DeepOBS All-CNN-C; restates preds/models.py:182-236 (flat module list; TF-"same" padding as an explicit
    SamePad2d in front of the padding-0 conv, conv7 unpadded, no ReLU after conv9)."""
    spec = [(in_channels, 96, 3, 1, True), (96, 96, 3, 1, True), (96, 96, 3, 2, True),
            (96, 192, 3, 1, True), (192, 192, 3, 1, True), (192, 192, 3, 2, True),
            (192, 192, 3, 1, False), (192, 192, 1, 1, True), (192, n_out, 1, 1, True)]
    mods = []
    for i, (cin, cout, k, s, same) in enumerate(spec):
        if same:
            mods.append(SamePad2d(k, s))
        mods.append(nn.Conv2d(cin, cout, k, stride=s))
        if i + 1 < len(spec):
            mods.append(nn.ReLU())
    mods += [nn.AvgPool2d(6), nn.Flatten()]
    net = nn.Sequential(*mods)
    for m in net.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.constant_(m.bias, 0.1)
            nn.init.xavier_normal_(m.weight)
    net.output_size = n_out
    return net


def n_params(model):
    return sum(p.numel() for p in model.parameters())


def get_theta(model):
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()])


def set_theta(model, theta):
    off = 0
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(theta[off:off + p.numel()].reshape(p.shape))
            off += p.numel()


# ------------------------------------------------------------------------------------------
# likelihoods  (preds/likelihoods.py)
# ------------------------------------------------------------------------------------------
class CategoricalLh:
    name = 'categorical'

    def inv_link(self, f):                       # :95-96
        return torch.softmax(f, dim=-1)

    def hessian(self, f):                        # :90-93  (clamped softmax)
        p = self.inv_link(f).clamp(CLAMP_EPS, 1 - CLAMP_EPS)
        return torch.diag_embed(p) - p.unsqueeze(2) * p.unsqueeze(1)

    def residual(self, y, f):                    # :84-88
        onehot = torch.zeros_like(f)
        onehot[torch.arange(len(y)), y.long()] = 1
        return onehot - self.inv_link(f)

    def hess_factor(self):                       # :98-99  nn_loss -> (CE(sum), 1)
        return 1.0

    def loss_hessian_sqrt(self, f):
        """BackPACK CrossEntropyLoss sqrt-Hessian, UNclamped softmax (Appendix A)."""
        p = torch.softmax(f, dim=1)
        sq = p.sqrt()
        return torch.diag_embed(sq) - p.unsqueeze(2) * sq.unsqueeze(1)


class GaussianLh:
    name = 'gaussian'

    def __init__(self, sigma_noise=1.0):
        self.sigma_noise = sigma_noise

    def inv_link(self, f):                       # :54-55
        return f

    def hessian(self, f):                        # :50-52
        assert f.size(1) == 1
        return torch.ones_like(f).unsqueeze(-1) / self.sigma_noise ** 2

    def residual(self, y, f):                    # :47-48
        return (y - f) / self.sigma_noise ** 2

    def hess_factor(self):                       # :57-58  nn_loss -> (MSE(sum), 1/(2 s^2))
        return 1.0 / (2 * self.sigma_noise ** 2)

    def loss_hessian_sqrt(self, f):
        """BackPACK MSELoss(sum) sqrt-Hessian: sqrt(2) I."""
        N, K = f.shape
        return math.sqrt(2.0) * torch.eye(K, dtype=f.dtype).expand(N, K, K).clone()


class BernoulliLh:
    name = 'bernoulli'

    def inv_link(self, f):                       # :71-72
        return torch.sigmoid(f)

    def hessian(self, f):                        # :67-69
        p = self.inv_link(f).clamp(CLAMP_EPS, 1 - CLAMP_EPS)
        return p * (1 - p)

    def residual(self, y, f):                    # :25-26
        return y - self.inv_link(f)

    def hess_factor(self):                       # :74-75
        raise ValueError('No extendable nn loss for backpack in Bernoulli case')


# ------------------------------------------------------------------------------------------
# per-sample layer quantities (BackPACK BatchGrad / DiagGGNExact / KFLR, Appendix A)
# ------------------------------------------------------------------------------------------
def _param_layers(model):
    return [m for m in model.modules() if isinstance(m, (nn.Linear, nn.Conv2d))]


def _forward_recording(model, X):
    """Forward pass keeping every parameterised layer's input and (graph-attached) output."""
    rec, handles = [], []
    for m in _param_layers(model):
        handles.append(m.register_forward_hook(
            lambda mod, inp, out, rec=rec: rec.append((mod, inp[0].detach(), out))))
    try:
        f = model(X)
    finally:
        for h in handles:
            h.remove()
    return f, rec


def _unfold(m, a):
    return F.unfold(a, m.kernel_size, dilation=m.dilation, padding=m.padding, stride=m.stride)


def jacobians(model, X):
    """Js [N,P,K] (K last; [N,P] when K == 1) and f [N,K]; restates preds/gradients.py:20-46
    with BackPACK ``BatchGrad`` written out per layer (Linear: g (x) a; Conv2d: unfold-einsum)."""
    K = model.output_size
    N = X.shape[0]
    cols, f_out = [], None
    for k in range(K):
        f, rec = _forward_recording(model, X)
        if f_out is None:
            f_out = f.detach()
        target = f[:, k].sum() if K > 1 else f.sum()
        grads = torch.autograd.grad(target, [out for (_, _, out) in rec])
        per_param = {}
        for (m, a, _), g in zip(rec, grads):
            if isinstance(m, nn.Linear):
                per_param[m.weight] = (g.unsqueeze(2) * a.unsqueeze(1)).reshape(N, -1)
                if m.bias is not None:
                    per_param[m.bias] = g
            else:
                go = g.reshape(N, g.shape[1], -1)
                per_param[m.weight] = torch.einsum('nil,nol->noi', _unfold(m, a), go).reshape(N, -1)
                if m.bias is not None:
                    per_param[m.bias] = go.sum(2)
        cols.append(torch.cat([per_param[p] for p in model.parameters()], dim=1))
    if K > 1:
        return torch.stack(cols, dim=2), f_out
    return cols[0], f_out


def ggn_bundle(model, lh, X, y=None):
    """(Js [m,p,k], Hess [m,k,k], rs [m,k] | None, f); restates preds/optimizers.py:8-25."""
    Js, f = jacobians(model, X)
    Hess = lh.hessian(f)
    m, p = Js.shape[:2]
    rs = lh.residual(y, f) if y is not None else None
    if Js.dim() == 2:
        Js, Hess = Js.reshape(m, p, 1), Hess.reshape(m, 1, 1)
        rs = rs.reshape(m, 1) if rs is not None else None
    return Js, Hess, rs, f


def _sqrt_ggn_backprop(model, lh, X):
    """S_l[v, n, out, (H, W)] for every parameterised layer: the loss-Hessian square root
    backpropagated to the layer outputs (BackPACK second-order convention, Appendix A)."""
    f, rec = _forward_recording(model, X)
    S = lh.loss_hessian_sqrt(f.detach())
    outs = [out for (_, _, out) in rec]
    per_v = [torch.autograd.grad(f, outs, grad_outputs=S[:, :, v], retain_graph=True)
             for v in range(S.shape[2])]
    Sl = [torch.stack([per_v[v][i] for v in range(S.shape[2])]).detach() for i in range(len(rec))]
    return rec, Sl


def diag_ggn_batch(model, lh, X):
    """Flat exact GGN diagonal of one batch (no hess_factor); BackPACK ``DiagGGNExact`` as read
    by preds/laplace.py:14-15,52-59."""
    rec, Sls = _sqrt_ggn_backprop(model, lh, X)
    out = {}
    for (m, a, _), Sl in zip(rec, Sls):
        V, N = Sl.shape[:2]
        if isinstance(m, nn.Linear):
            out[m.weight] = torch.einsum('vno,ni->oi', Sl ** 2, a ** 2)
            if m.bias is not None:
                out[m.bias] = (Sl ** 2).sum((0, 1))
        else:
            So = Sl.reshape(V, N, Sl.shape[2], -1)
            JS = torch.einsum('nil,vnol->vnoi', _unfold(m, a), So)
            out[m.weight] = (JS ** 2).sum((0, 1))
            if m.bias is not None:
                out[m.bias] = (So.sum(3) ** 2).sum((0, 1))
    return torch.cat([out[p].reshape(-1) for p in model.parameters()])


def kflr_batch(model, lh, X):
    """Per-parameter Kronecker factors of one batch, [G, A] for weights / [G] for biases;
    BackPACK ``KFLR`` as read by preds/laplace.py:73-78 (G sums over the batch, A is a batch mean)."""
    rec, Sls = _sqrt_ggn_backprop(model, lh, X)
    out = {}
    for (m, a, _), Sl in zip(rec, Sls):
        V, N = Sl.shape[:2]
        if isinstance(m, nn.Linear):
            G = torch.einsum('vni,vnj->ij', Sl, Sl)
            A = a.t() @ a / N
        else:
            s = Sl.reshape(V, N, Sl.shape[2], -1).sum(3)
            G = torch.einsum('vni,vnj->ij', s, s)
            U = _unfold(m, a)
            A = torch.einsum('bik,bjk->ij', U, U) / N
        out[m.weight] = [G, A]
        if m.bias is not None:
            out[m.bias] = [G]
    return [out[p] for p in model.parameters()]


# ------------------------------------------------------------------------------------------
# Kronecker-factored posterior  (preds/kron.py)
# ------------------------------------------------------------------------------------------
def symeig(M):
    """Eigendecomposition from the UPPER triangle, eigenvalues clamped at 0, NaN -> 0;
    restates preds/kron.py:146-173 (the jitter retry only triggers on LAPACK failure)."""
    try:
        L, W = torch.linalg.eigh(M, UPLO='U')
    except RuntimeError:
        L, W = torch.linalg.eigh(M + torch.eye(M.shape[0], dtype=M.dtype), UPLO='U')
        L = L - 1.0
    L = L.clamp(min=0.0)
    L[torch.isnan(L)] = 0.0
    W[torch.isnan(W)] = 0.0
    return L, W


class KronState:
    """Plain container: qhs (precision factors), Ws/Lams (eigen), deltas, dampen."""

    def __init__(self, model, delta, dampen=False):
        self.qhs = []                                  # preds/kron.py:25-43
        for p in model.parameters():
            if p.dim() == 1:
                self.qhs.append([torch.zeros(p.shape[0], p.shape[0])])
            else:
                m, n = p.shape[0], p[0].numel()
                self.qhs.append([torch.zeros(m, m), torch.zeros(n, n)])
        d = torch.as_tensor(delta)                     # :45-56
        if d.dim() == 0 or (d.dim() == 1 and len(d) == 1):
            self.deltas = d.reshape(-1)[:1].repeat(len(self.qhs))
        elif d.dim() == 1:
            if len(d) != len(self.qhs):
                raise ValueError('Deltas do not match with parameter groups.')
            self.deltas = d
        else:
            raise ValueError('Invalid dimensionality of delta.')
        self.dampen = dampen
        self.Ws = self.Lams = None

    def update(self, krons, factor=1.0, batch_factor=1.0):      # :58-69
        for kr, qh in zip(krons, self.qhs):
            if len(kr) == 1:
                qh[0] += kr[0] * factor
            elif len(kr) == 2:
                qh[0] += kr[0] * factor
                qh[1] += kr[1] * batch_factor
            else:
                raise ValueError('..')

    def decompose(self):                                        # :71-83
        eig = [[symeig(m) for m in qh] for qh in self.qhs]
        self.Lams = [[e[0] for e in pe] for pe in eig]
        self.Ws = [[e[1] for e in pe] for pe in eig]

    def sample(self, eps_list):
        """theta-samples [S,P] ~ N(0, (G (x) A + delta I)^-1) from injected draws, one per
        parameter in parameters() order: [p,S] for 1-factor, [S,m,n] for 2-factor; :85-110."""
        parts = []
        for l, w, delta, eps in zip(self.Lams, self.Ws, self.deltas, eps_list):
            if len(l) == 1:
                scale = torch.sqrt(1 / (l[0] + delta)).reshape(-1, 1)
                parts.append((w[0] @ (scale * (w[0].t() @ eps))).t())
            else:
                (l1, l2), (W1, W2) = l, w
                if self.dampen:
                    D = torch.sqrt(1 / torch.outer(l1 + torch.sqrt(delta), l2 + torch.sqrt(delta)))
                else:
                    D = torch.sqrt(1 / (torch.outer(l1, l2) + delta))
                z = (W1.t() @ eps @ W2) * D.unsqueeze(0)
                z = W1 @ z @ W2.t()
                parts.append(z.reshape(eps.shape[0], -1))
        return torch.cat(parts, dim=1)

    def functional_variance(self, Js):
        """J Sigma J^T for Js [B,K,P]; restates preds/kron.py:112-143 (hess_factor = 1)."""
        B, K, P = Js.shape
        Jf = Js.reshape(B * K, P)
        cur, SJ = 0, []
        for l, w, delta in zip(self.Lams, self.Ws, self.deltas):
            if len(l) == 1:
                p = len(l[0])
                inv = (1 / (l[0] + delta)).reshape(-1, 1)
                Jp = Jf[:, cur:cur + p].t()
                SJ.append((w[0] @ (inv * (w[0].t() @ Jp))).t())
            else:
                (l1, l2), (W1, W2) = l, w
                m, n = len(l1), len(l2)
                p = m * n
                if self.dampen:
                    Dinv = 1 / torch.outer(l1 + torch.sqrt(delta), l2 + torch.sqrt(delta))
                else:
                    Dinv = 1 / (torch.outer(l1, l2) + delta)
                Jp = Jf[:, cur:cur + p].reshape(B * K, m, n)
                Jp = (W1.t() @ Jp @ W2) * Dinv.unsqueeze(0)
                SJ.append((W1 @ Jp @ W2.t()).reshape(B * K, p))
            cur += p
        SJ = torch.cat(SJ, dim=1).reshape(B, K, P)
        return torch.bmm(Js, SJ.transpose(1, 2))


# ------------------------------------------------------------------------------------------
# Laplace.infer  (preds/laplace.py:30-82, :115-135)
# ------------------------------------------------------------------------------------------
def expand_prior_precision(prior_prec, P, cov_type):
    pp = torch.as_tensor(prior_prec, dtype=torch.float32)
    if pp.dim() == 0:
        d = torch.ones(P) * pp
        return torch.diag(d) if cov_type == 'full' else d
    if pp.dim() == 1:
        return torch.diag(pp) if cov_type == 'full' else pp.clone()
    if pp.dim() == 2:
        if cov_type == 'diag':
            raise ValueError('Will not diagonalize non-diagonal prior.')
        return pp.clone()
    raise ValueError('Invalid shape for prior precision')


def ggn_full_accumulate(model, lh, batches, precision):
    """precision += sum_batches einsum('mpk,mkl,mql->pq'); preds/laplace.py:39-42."""
    for X, y in batches:
        Js, Hess, _, _ = ggn_bundle(model, lh, X, y)
        precision += torch.einsum('mpk,mkl,mql->pq', Js, Hess, Js)
    return precision


def factorise_full(precision):
    """chol(H) -> Sigma = cholesky_inverse -> chol(Sigma) (lower); preds/laplace.py:43-45."""
    chol = torch.linalg.cholesky(precision)
    Sigma = torch.cholesky_inverse(chol, upper=False)
    return torch.linalg.cholesky(Sigma)


def infer(model, prior_prec, lh, batches, cov_type='full', dampen_kron=False):
    """Returns dict(mu, cov_type, and one of Sigma_chol [P,P] / Sigma_chol [P] / kron)."""
    model.eval()
    batches = list(batches)
    P = n_params(model)
    post = {'mu': get_theta(model), 'cov_type': cov_type}
    if cov_type == 'full':
        H = ggn_full_accumulate(model, lh, batches, expand_prior_precision(prior_prec, P, 'full'))
        post['precision'] = H.clone()
        post['Sigma_chol'] = factorise_full(H)
    elif cov_type == 'diag':                                   # :47-60
        hf = lh.hess_factor()
        prec = expand_prior_precision(prior_prec, P, 'diag')
        for X, _ in batches:
            prec = prec + hf * diag_ggn_batch(model, lh, X)
        post['precision'] = prec
        post['Sigma_chol'] = torch.sqrt(1 / prec)
    elif cov_type == 'kron':                                   # :62-80
        hf = lh.hess_factor()
        kron = KronState(model, prior_prec, dampen=dampen_kron)
        N = sum(len(b[0]) for b in batches)
        for X, y in batches:
            kron.update(kflr_batch(model, lh, X), factor=hf, batch_factor=len(y) / N)
        kron.decompose()
        post['kron'] = kron
    return post


# ------------------------------------------------------------------------------------------
# predictives  (preds/predictives.py:12-28, :50-76)
# ------------------------------------------------------------------------------------------
def glm_moments(model, X, post):
    """f_mu [N,K], f_var [N,K,K]; restates preds/predictives.py:52-65."""
    theta_star = get_theta(model)
    Js, f = jacobians(model, X)
    Js = Js.transpose(1, 2) if Js.dim() > 2 else Js.unsqueeze(1)
    f_mu = f + Js @ (post['mu'] - theta_star)
    if post['cov_type'] == 'kron':
        f_var = post['kron'].functional_variance(Js)
    elif post['cov_type'] == 'full':
        Sigma = post['Sigma_chol'] @ post['Sigma_chol'].t()          # preds/laplace.py:111
        f_var = torch.einsum('nkp,pq,ncq->nkc', Js, Sigma, Js)
    else:
        f_var = torch.einsum('nkp,p,ncp->nkc', Js, post['Sigma_chol'].square(), Js)
    return f_mu, f_var


def glm_predictive(model, lh, X, post, eps):
    """softmax(f_mu + chol(f_var) eps) -> [S,N,K] for injected eps [S,N,K]; for GaussianLh
    returns (f_mu, f_var) like preds/predictives.py:66-67.  Cholesky failure falls back to
    the reference's diagonal branch (variance used as Normal scale, :73-76)."""
    f_mu, f_var = glm_moments(model, X, post)
    if isinstance(lh, GaussianLh):
        return f_mu, f_var
    try:
        L = torch.linalg.cholesky(f_var)
        fs = f_mu.unsqueeze(0) + torch.einsum('nkl,snl->snk', L, eps)
    except RuntimeError:
        scale = f_var.diagonal(dim1=1, dim2=2).clamp(1e-5)
        fs = f_mu.unsqueeze(0) + scale.unsqueeze(0) * eps
    return lh.inv_link(fs)


def bnn_weight_samples(post, eps):
    """theta samples [S,P]; eps is [P,S] (full/diag) or the per-parameter list (kron);
    restates preds/predictives.py:14-20."""
    if post['cov_type'] == 'kron':
        covs = post['kron'].sample(eps)
    elif post['cov_type'] == 'full':
        covs = (post['Sigma_chol'] @ eps).t()
    else:
        covs = (post['Sigma_chol'].reshape(-1, 1) * eps).t()
    return post['mu'] + covs


def bnn_predictive(model, lh, X, post, eps):
    """S forward passes at sampled weights -> [S,N,K]; restates preds/predictives.py:21-28."""
    theta_star = get_theta(model)
    samples = bnn_weight_samples(post, eps)
    preds = []
    with torch.no_grad():
        for s in range(samples.shape[0]):
            set_theta(model, samples[s])
            preds.append(lh.inv_link(model(X)).detach())
    set_theta(model, theta_star)
    return torch.stack(preds)


def kron_eps_shapes(model, S):
    """Shapes of the draws Kron.sample consumes, in parameters() order (preds/kron.py:92,106)."""
    shapes = []
    for p in model.parameters():
        if p.dim() == 1:
            shapes.append((p.shape[0], S))
        else:
            shapes.append((S, p.shape[0], p[0].numel()))
    return shapes


# ------------------------------------------------------------------------------------------
# LaplaceGGN.post_process, get_diagonal_ggn, linear_sampling_predictive
# ------------------------------------------------------------------------------------------
def post_process(model, lh, batches, prior_prec_matrix, prior_mu):
    """Posterior of the linearised model; restates preds/optimizers.py:81-104.
    Returns dict(precision, Sigma, Sigma_chol, mu, G)."""
    theta_star = get_theta(model)
    P = len(theta_star)
    JLJ, G = torch.zeros(P, P), torch.zeros(P)
    for X, y in batches:
        Js, Hess, rs, _ = ggn_bundle(model, lh, X, y)
        JLJ += torch.einsum('mpk,mkl,mql->pq', Js, Hess, Js)
        G += torch.einsum('mpk,mk->p', Js, rs)
    precision = JLJ + prior_prec_matrix
    chol = torch.linalg.cholesky(precision)
    Sigma = torch.cholesky_inverse(chol, upper=False)
    b = G + JLJ @ theta_star + prior_prec_matrix @ prior_mu
    mu = torch.cholesky_solve(b.reshape(-1, 1), chol, upper=False).flatten()
    return {'precision': precision, 'Sigma': Sigma, 'Sigma_chol': torch.linalg.cholesky(Sigma), 'mu': mu, 'G': G}


def diagonal_ggn(precision):
    """(Sigma, Sigma_chol) as diagonal matrices; restates preds/optimizers.py:107-112."""
    Sigma_diag = 1 / torch.diag(precision)
    return torch.diag(Sigma_diag), torch.diag(torch.sqrt(Sigma_diag))


def linear_sampling_predictive(model, lh, X, mu, Sigma_chol, eps, no_link=False):
    """link(f + J (theta_s - theta*)) for theta_s = mu + Sigma_chol eps, eps [P,S];
    restates preds/predictives.py:31-47 ([S,N,K]; [S,N] when the model flattens a single output)."""
    theta_star = get_theta(model)
    Js, f = jacobians(model, X)
    if Js.dim() > 2:
        Js = Js.transpose(1, 2)
    offset = f - Js @ theta_star
    covs = (Sigma_chol @ eps).t() if Sigma_chol.dim() == 2 else (Sigma_chol.reshape(-1, 1) * eps).t()
    samples = mu + covs
    link = (lambda x: x) if no_link else lh.inv_link
    return torch.stack([link(offset + Js @ samples[i]) for i in range(samples.shape[0])])


# ------------------------------------------------------------------------------------------
# delta-method regression predictive and the metric reductions
# ------------------------------------------------------------------------------------------
def linear_regression_predictive(model, lh, X, mu, Sigma_chol):
    """(mu_star, var_f, var_noise); restates preds/predictives.py:79-94."""
    theta_star = get_theta(model)
    Sigma = Sigma_chol @ Sigma_chol.t() if Sigma_chol.dim() == 2 else torch.diag(Sigma_chol ** 2)
    Js, Hess, _, f = ggn_bundle(model, lh, X)
    s2 = lh.sigma_noise ** 2 if lh.name == 'gaussian' else 1.0      # sigma_2_dispersion, preds/likelihoods.py:20-22,40-42
    Lams = Hess * s2                                                # get_Lams_Vys, preds/likelihoods.py:7-12
    Vys = Lams * s2
    Jgs = torch.bmm(Js, Lams)
    lin_pred = torch.einsum('mpk,p->mk', Jgs, mu - theta_star).reshape(*f.shape)
    mu_star = lh.inv_link(f) + lin_pred
    var_f = torch.bmm(Jgs.transpose(1, 2) @ Sigma, Jgs).squeeze()
    return mu_star, var_f, Vys.squeeze()


def classification_metrics(p, y, lik, bins=10):
    """{'nll','acc','ece'}; restates preds/utils.py:6-52 (lik: 'categorical' with p [N,K], or 'bernoulli'
    with p [N]) in plain fp32/fp64 tensor code."""
    eps = torch.finfo(torch.float32).eps
    N = len(y)
    if lik == 'categorical':
        pn = (p / p.sum(-1, keepdim=True)).clamp(eps, 1 - eps)            # Categorical(probs=p).logits
        nll = -pn.log().gather(1, y.long().reshape(-1, 1)).mean().item()
        acc = (p.argmax(-1) == y).sum().item() / N
        probs = p
    else:
        pc = p.clamp(eps, 1 - eps)                                         # Bernoulli(probs=p).log_prob(y)
        nll = -(y * pc.log() + (1 - y) * torch.log1p(-pc)).mean().item()
        acc = ((p >= 0.5).float() == y).float().sum().item() / N
        probs = torch.stack([1 - p, p]).t()
    conf, pred = probs.max(1)
    hit = pred.eq(y.long())
    edges = torch.linspace(0, 1, bins + 1)
    ece = 0.0
    for lo, hi in zip(edges[:-1], edges[1:]):
        in_bin = (conf > lo) & (conf <= hi)
        if in_bin.any():
            ece += abs(conf[in_bin].double().mean().item() - hit[in_bin].double().mean().item()) * in_bin.double().mean().item()
    return {'nll': nll, 'acc': acc, 'ece': ece}


# ------------------------------------------------------------------------------------------
# GLM refinement on a fixed Jacobian (preds/refine.py)
# ------------------------------------------------------------------------------------------
def _glm_parts(model, X):
    J, f = jacobians(model, X)
    if J.dim() == 3:
        J = J.transpose(1, 2)                       # [N,K,P]
    theta_star = get_theta(model)
    return J, f - J @ theta_star, theta_star


def _log_lik(lh, y, f):
    """preds/likelihoods.py:31-32,44-45,64-65,81-82 (sum of log-probabilities)."""
    if lh.name == 'categorical':
        return torch.distributions.Categorical(logits=f).log_prob(y).sum()
    if lh.name == 'bernoulli':
        return torch.distributions.Bernoulli(logits=f).log_prob(y).sum()
    return torch.distributions.Normal(f, lh.sigma_noise).log_prob(y).sum()


def _jlj(J, Lams):
    if J.dim() == 3:
        return torch.einsum('mkp,mkl,mlq->pq', J, Lams, J)
    return (J.T * Lams.reshape(1, -1)) @ J


def _jtr(J, rs):
    return torch.einsum('mkp,mk->p', J, rs) if J.dim() == 3 else J.T @ rs


def laplace_refine(model, X, y, lh, prior_prec, n_epochs=1000, lr=1e-3):
    """(mu, Sigma_chol, Sigma_chol_diag, losses); restates preds/refine.py:9-40."""
    J, f_offset, theta_star = _glm_parts(model, X)
    mu = theta_star.clone().requires_grad_(True)
    opt = torch.optim.Adam([mu], lr=lr)
    losses = []
    for _ in range(n_epochs):
        opt.zero_grad()
        loss = -_log_lik(lh, y, f_offset + J @ mu) + 0.5 * mu @ mu * prior_prec
        loss.backward()
        losses.append(loss.item())
        opt.step()
    mu = mu.detach()
    precision = _jlj(J, lh.hessian(f_offset + J @ mu)) + prior_prec * torch.eye(len(mu))
    chol = torch.linalg.cholesky(precision)
    Sigma_chol = torch.linalg.cholesky(torch.cholesky_inverse(chol, upper=False))
    return mu, Sigma_chol, torch.diag(torch.sqrt(1 / torch.diag(precision))), losses


def vi_refine(model, post, prior_prec_matrix, X, y, lh, eps, lr=1e-2):
    """(mu, Sigma_chol, losses) after len(eps) epochs; restates preds/refine.py:42-81.
    ``post`` = dict(precision, Sigma_chol) of the LaplaceGGN posterior; eps [n_epochs, P]."""
    from torch.distributions import MultivariateNormal, kl_divergence
    beta = lr
    J, f_offset, theta_star = _glm_parts(model, X)
    mu = theta_star.clone()
    Sigma_chol, prec = post['Sigma_chol'].clone(), post['precision'].clone()
    p = MultivariateNormal(torch.zeros_like(mu), scale_tril=torch.diag(torch.sqrt(1 / torch.diag(prior_prec_matrix))))
    losses = []
    for e in eps:
        f_t = f_offset + J @ (mu + Sigma_chol @ e)
        H = _jlj(J, lh.hessian(f_t))
        g = _jtr(J, lh.residual(y, f_t)) - prior_prec_matrix @ mu
        prec = (1 - beta) * prec + beta * (H + prior_prec_matrix)
        pchol = torch.linalg.cholesky(prec)
        mu = mu + beta * torch.cholesky_solve(g.reshape(-1, 1), pchol, upper=False).squeeze()
        Sigma_chol = torch.linalg.cholesky(torch.cholesky_inverse(pchol, upper=False))
        q = MultivariateNormal(loc=mu, scale_tril=Sigma_chol)
        losses.append((-_log_lik(lh, y, f_t) + kl_divergence(q, p)).item())
    return mu, Sigma_chol, losses


def vi_diag_refine(model, post, prior_prec_matrix, X, y, lh, eps, lr=1e-3):
    """(mu, diag-matrix Sigma_chol, losses); restates preds/refine.py:84-121."""
    from torch.distributions import Normal, kl_divergence
    beta = lr
    J, f_offset, theta_star = _glm_parts(model, X)
    mu = theta_star.clone()
    prec = torch.diag(post['precision']).clone()
    prior = torch.diag(prior_prec_matrix)
    sigma_chol = torch.sqrt(1 / prec)
    m_prior = Normal(loc=torch.zeros_like(mu), scale=1 / torch.sqrt(prior))
    losses = []
    for e in eps:
        f_t = f_offset + J @ (mu + sigma_chol * e)
        H = torch.diag(_jlj(J, lh.hessian(f_t)))
        g = _jtr(J, lh.residual(y, f_t)) - prior * mu
        prec = (1 - beta) * prec + beta * (H + prior)
        mu = mu + beta * g / prec
        sigma_chol = torch.sqrt(1 / prec)
        losses.append((-_log_lik(lh, y, f_t) + kl_divergence(Normal(loc=mu, scale=sigma_chol), m_prior).sum()).item())
    return mu, torch.diag(sigma_chol), losses


# ---------------------------------------------------------------------------------------------------------------------
# Function-space inference: preds/gps.py:664-927 gp_quantities, preds/laplace.py:138-207
# FunctionaLaplace, preds/gps.py:384-447 GP_predictive, preds/gps.py:930-952 mc_preds.  Pinned against the unmodified
# reference by tests/golden/gp_*.npz (oracle/gen_golden.py, run through refharness.cuda_standins).
# ---------------------------------------------------------------------------------------------------------------------
def gp_quantities(model, X_train, X_test, prior_var=None, outputs=None, only_pred_var=True):
    """The dictionary of preds/gps.py:910-925 for one prior-variance group (kernel slot axis dropped, as
    ``collapse_groups=True`` does): per output c K_train[c] = J_c J_c^T (preds/gps.py:64-88 ``Parallel_calc_kernel`` summed
    over the parameter slices), K_train_test, K_test = its diagonal (``only_pred_var``) or the full block,
    f_gp = J_c theta (``Parallel_calc_gp_function``, :91-97), f_star = f^T[outputs].  ``prior_var`` scales the kernels
    (``with_prior_prec=True``); FunctionaLaplace passes ``with_prior_prec=False``."""
    theta = get_theta(model)
    Jtr, ftr = jacobians(model, X_train)              # [n, P, C]
    Jte, fte = jacobians(model, X_test)
    if Jtr.dim() == 2:
        Jtr, Jte = Jtr.unsqueeze(2), Jte.unsqueeze(2)
    ftr, fte = ftr.reshape(len(X_train), -1), fte.reshape(len(X_test), -1)
    if outputs is not None:
        Jtr, Jte, ftr, fte = Jtr[:, :, outputs], Jte[:, :, outputs], ftr[:, outputs], fte[:, outputs]
    s = 1.0 if prior_var is None else float(prior_var)
    return {
        'K_train': s * torch.einsum('npc,mpc->cnm', Jtr, Jtr),
        'K_train_test': s * torch.einsum('npc,mpc->cnm', Jtr, Jte),
        'K_test': s * (torch.einsum('mpc,mpc->cm', Jte, Jte) if only_pred_var else torch.einsum('npc,mpc->cnm', Jte, Jte)),
        'f_gp_train': torch.einsum('npc,p->cn', Jtr, theta),
        'f_gp_test': torch.einsum('mpc,p->cm', Jte, theta),
        'f_star_train': ftr.t().contiguous(),
        'f_star_test': fte.t().contiguous(),
    }


def gp_predictive_moments(q, prior_prec, indep=False):
    """(mean [m, C], covariance [m, C, C]) of the distribution FunctionaLaplace.predictive_samples draws from
    (preds/laplace.py:180-205): the kernels divided by the prior precision; ``indep`` -> per-class binary GPs with
    W = p (1 - p) (:186-199); otherwise GP_predictive's categorical branch (preds/gps.py:411-426, E / M / b / c of
    ``Parallel_calc_E_z`` :99-110 and ``Parallel_calc_b_c_pred`` :113-122); the mean is replaced by f_star_test (:205)."""
    K_nn, K_nm, K_m = q['K_train'] / prior_prec, q['K_train_test'] / prior_prec, q['K_test'] / prior_prec
    C, n, m = K_nm.shape
    mean = q['f_star_test'].t()
    cov = torch.zeros(m, C, C)
    if indep:
        p = torch.softmax(q['f_star_train'], dim=0)
        w = torch.sqrt(p * (1 - p))
        B = w.unsqueeze(2) * K_nn * w.unsqueeze(1)
        B.diagonal(dim1=1, dim2=2).add_(1.0)
        Kinv = w.unsqueeze(2) * torch.inverse(B) * w.unsqueeze(1)
        gain = torch.einsum('cnm,cnp,cpm->cm', K_nm, Kinv, K_nm)
        cov.diagonal(dim1=1, dim2=2).copy_((K_m - gain).t())
        return mean, cov
    pi = torch.softmax(q['f_star_train'], dim=0)
    d = pi.sqrt()
    A = K_nn * d.unsqueeze(1) * d.unsqueeze(2)
    A.diagonal(dim1=1, dim2=2).add_(1.0)
    L = torch.linalg.cholesky(A)
    E = torch.cholesky_solve(torch.diag_embed(d), L) * d.unsqueeze(2)
    M = torch.linalg.cholesky(E.sum(0))
    b = torch.einsum('cnm,cni->cmi', E, K_nm)
    c_ = torch.cholesky_solve(b, M)
    cov = torch.einsum('mcn,mnd->mcd', b.permute(2, 0, 1), c_.permute(2, 1, 0)).contiguous()
    diag_term = torch.einsum('mcn,mnc->mc', K_nm.permute(2, 0, 1), b.permute(2, 1, 0))
    cov.diagonal(dim1=1, dim2=2).add_(K_m.t() - diag_term)
    return mean, cov


def gp_mc_mean(mean, cov, eps):
    """mc_preds' Monte-Carlo mean of softmax(f) (preds/gps.py:930-952) for injected eps [S, m, C]:
    f = mean + chol(cov) eps (MultivariateNormal) -- for a diagonal cov that is Normal(mean, sqrt(var)) (:199)."""
    L = torch.linalg.cholesky(cov)
    fs = mean.unsqueeze(0) + torch.einsum('mcd,smd->smc', L, eps)
    return torch.softmax(fs, dim=-1).mean(0)
