"""Size-independent properties at BASELINE.json's full size (config 5: P = 59 635, K = 10), where the
CPU oracle cannot finish: symmetry, a checksum of checksums (trace of the full GGN == sum of the
diagonal-GGN path), and linearity (H v against an independent fp64 evaluation of
sum_n J_n^T Lambda_n (J_n v) from the layer factors)."""
import ctypes as C

import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu


def test_cfg5_full_ggn_properties():
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip('needs ~45 GB of free device memory')
    from preds_b200 import Laplace, CategoricalLh, MLPS
    from preds_b200._engine import net_for
    torch.manual_seed(71)
    model = MLPS(784, [75], 10, 'tanh', flatten=True).cuda()
    P = sum(p.numel() for p in model.parameters())
    assert P == 59635
    g = torch.Generator().manual_seed(15)
    N = 6000                                    # > one 4096-example accumulation segment, ragged tail
    X = torch.randn(N, 1, 28, 28, generator=g)
    y = torch.randint(10, (N,), generator=g)
    lh = CategoricalLh()
    lap = Laplace(model, 0.0, lh)
    lap.ggn_engine = 2
    lap.infer([(X[:2500], y[:2500]), (X[2500:], y[2500:])], cov_type='full', factorise=False)
    H = lap.precision
    # (1) symmetry (exact after the mirror)
    idx = torch.randint(0, P, (4096,), device='cuda')
    jdx = torch.randint(0, P, (4096,), device='cuda')
    assert torch.equal(H[idx, jdx], H[jdx, idx])
    # (2) checksum of checksums: trace(full) == sum(diag path); the diag path uses BackPACK's
    #     unclamped softmax, the full path the clamped one -> agree to ~1e-5
    lapd = Laplace(model, 0.0, lh)
    lapd.infer([(X, y)], cov_type='diag', factorise=False)
    d_full = H.diagonal().double()
    d_diag = lapd.precision.double()
    assert float((d_full - d_diag).abs().max() / d_diag.abs().max()) < 1e-4
    assert abs(float(d_full.sum() / d_diag.sum()) - 1) < 5e-5
    # (3) linearity: H v vs fp64 sum_n J_n^T Lambda_n J_n v from the factors (independent of the gram kernel)
    net = net_for(model)
    X2d, pl, ws, f = net.jacobian_factors(X.cuda())
    K, M, D = net.K, pl.Mtot, pl.Dtot
    A = ws[pl.acat_off: pl.acat_off + N * D].reshape(N, D).double()
    G = ws[pl.gcat_off: pl.gcat_off + N * K * M].reshape(N, K, M).double()
    Lam = lh._hessian_nkk(f).double()
    v = torch.randn(P, generator=torch.Generator().manual_seed(5)).cuda().double()
    W1, b1 = v[:58800].reshape(75, 784), v[58800:58875]
    W2, b2 = v[58875:59625].reshape(10, 75), v[59625:]
    c2, u2 = pl.op[2].lin_col, pl.op[2].lin_unit        # layer slices are padded to multiples of 4
    a1, a2 = A[:, :784], A[:, c2:c2 + 75]
    g1, g2 = G[:, :, :75], G[:, :, u2:u2 + 10]
    Jv = torch.einsum('nki,ni->nk', g1, a1 @ W1.T + b1) + torch.einsum('nki,ni->nk', g2, a2 @ W2.T + b2)
    w = torch.einsum('nkl,nl->nk', Lam, Jv)
    t1 = torch.einsum('nk,nki->ni', w, g1)
    t2 = torch.einsum('nk,nki->ni', w, g2)
    ref = torch.cat([(t1.T @ a1).reshape(-1), t1.sum(0), (t2.T @ a2).reshape(-1), t2.sum(0)])
    Hv = (H.double() @ v) if False else torch.cat([(H[i:i + 4096].double() @ v) for i in range(0, P, 4096)])
    assert float((Hv - ref).abs().max() / ref.abs().max()) < 1e-4
    net.release()


def test_cfg5_weight_sampling_tensor_core():
    """theta_s = mu + L eps at P = 59 635 (tensor-core path) against torch on a row subset."""
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip('needs ~30 GB of free device memory')
    from preds_b200.predictives import sample_weights
    P, S = 59635, 24
    g = torch.Generator(device='cuda').manual_seed(3)
    L = torch.tril(torch.randn(P, P, device='cuda', generator=g)) / P ** 0.5
    mu = torch.randn(P, device='cuda', generator=g)
    eps = torch.randn(P, S, device='cuda', generator=g)
    out = sample_weights(mu, L, S, eps=eps)                     # [S, P]
    rows = torch.randint(0, P, (512,), device='cuda', generator=g)
    ref = mu[rows].double().unsqueeze(0) + (L[rows].double() @ eps.double()).T
    assert float((out[:, rows].double() - ref).abs().max() / ref.abs().max()) < 2e-5


@pytest.mark.parametrize('cov', ['kron', 'diag'])
def test_glm_kron_diag_tensor_core_paths(cov):
    """Kron / diag GLM moments at a size where every contraction takes the tcgen05 branch (and the fused
    warp-per-input reduction) against the fp32 SIMT kernels with the unfused reductions on the same inputs
    (library option glm_simt; the SIMT path is the one pinned against the reference goldens at small sizes), plus the K x K
    symmetry and positive diagonal that f_var = J Sigma J^T must have."""
    from preds_b200 import MLPS, CategoricalLh, Laplace
    from preds_b200.predictives import glm_moments
    from preds_b200.kron import Kron
    torch.manual_seed(71)
    model = MLPS(784, [256, 128], 10, 'tanh', flatten=True).cuda()
    g = torch.Generator().manual_seed(15)
    X, y = torch.randn(512, 1, 28, 28, generator=g).cuda(), torch.randint(10, (512,), generator=g).cuda()
    Xte = torch.randn(384, 1, 28, 28, generator=g).cuda()
    lap = Laplace(model, 1.0, CategoricalLh())
    lap.infer([(X, y)], cov_type=cov)
    Sigma = lap.Sigma_chol if isinstance(lap.Sigma_chol, Kron) else lap.Sigma
    fmu, fvar = glm_moments(Xte, model, lap.mu, Sigma)
    from preds_b200._cabi import lib
    lib.bnnp_set_option(b'glm_simt', 1)
    try:
        fmu_ref, fvar_ref = glm_moments(Xte, model, lap.mu, Sigma)
    finally:
        lib.bnnp_set_option(b'glm_simt', 0)
    assert torch.equal(fmu, fmu_ref)
    assert cases.rel_err(fvar.cpu(), fvar_ref.cpu()) < 3e-5
    assert cases.rel_err(fvar.cpu(), fvar.transpose(1, 2).cpu()) < 1e-6
    assert bool((fvar.diagonal(dim1=1, dim2=2) > 0).all())


@pytest.mark.parametrize('arch', [(784, [75], 10, 2500), (300, [40, 24], 7, 1100), (500, [33], 4, 700)])
def test_gram_wide_kernel_is_bitwise_the_single_cta_kernel(arch):
    """The default full-GGN kernel (CTA pair, tcgen05.mma.cta_group::2 over two row tiles, ONE column unit x two column
    tiles sharing the generated A operand, cross-CTA barriers with a relaxed remote arrive) must reproduce the single-CTA
    kernel (two units sharing the B tile) bit for bit: same MMA sequence per accumulator, and every element of H
    receives exactly one add per launch.  Run twice to catch a rare ordering problem; odd tile counts, a single
    (unpaired) column tile and a ragged last K-segment (N not a multiple of 64) included."""
    from preds_b200 import MLPS, CategoricalLh, Laplace
    from preds_b200._cabi import lib
    din, hid, K, N = arch
    torch.manual_seed(11)
    model = MLPS(din, hid, K, 'tanh').cuda()
    g = torch.Generator().manual_seed(12)
    X, y = torch.randn(N, din, generator=g).cuda(), torch.randint(K, (N,), generator=g).cuda()

    def run():
        lap = Laplace(model, 1.0, CategoricalLh())
        lap.ggn_engine = 2
        lap.infer([(X, y)], cov_type='full', factorise=False)
        return lap.precision

    wide = [run() for _ in range(2)]
    lib.bnnp_set_option(b'gram_kernel', 1)
    try:
        single = run()
    finally:
        lib.bnnp_set_option(b'gram_kernel', 0)
    assert torch.equal(wide[0], wide[1])
    assert torch.equal(wide[0], single)
    assert torch.equal(single, single.T)


def test_row_chunked_accumulation_is_bitwise_the_whole_matrix():
    """bnnp_ggn_full_accum_rows over the chunks of bnnp_ggn_full_row_chunks (the data-parallel mode accumulates the last
    super-batch this way so that finished row chunks can be exchanged early) == one call over all rows, bit for bit."""
    import ctypes as C
    from preds_b200 import MLPS, CategoricalLh
    from preds_b200._cabi import lib, check, ptr, stream
    from preds_b200._engine import net_for
    torch.manual_seed(11)
    model = MLPS(300, [40, 24], 7, 'tanh').cuda()
    g = torch.Generator().manual_seed(12)
    X = torch.randn(1100, 300, generator=g).cuda()
    net = net_for(model)
    X2d, pl, ws, f = net.jacobian_factors(X)
    N, P = X2d.shape[0], net.P
    Lam = CategoricalLh()._hessian_nkk(f)
    sc = net.scratch(lib.bnnp_ggn_full_scratch_floats(C.byref(pl), N, 2))
    H0 = torch.zeros(P, P, device='cuda')
    check(lib.bnnp_ggn_full_accum(net.ops, net.n_ops, C.byref(pl), N, ptr(ws), ptr(Lam), ptr(H0), ptr(sc), 2, stream()))
    for want_chunks in (2, 5):
        bounds = (C.c_int64 * (want_chunks + 1))()
        n = lib.bnnp_ggn_full_row_chunks(net.ops, net.n_ops, C.byref(pl), N, 2, want_chunks, bounds)
        assert 1 < n <= want_chunks and bounds[0] == 0 and bounds[n] == P
        assert all(bounds[i] < bounds[i + 1] for i in range(n))
        H1 = torch.zeros(P, P, device='cuda')
        for c in range(n):
            check(lib.bnnp_ggn_full_accum_rows(net.ops, net.n_ops, C.byref(pl), N, ptr(ws), ptr(Lam), ptr(H1), ptr(sc), 2,
                                               bounds[c], bounds[c + 1], int(c > 0), stream()))
            # rows of later chunks are untouched so far (what the early exchange relies on)
            if c + 1 < n:
                assert float(H1[bounds[c + 1]:].abs().max()) == 0
        assert torch.equal(torch.tril(H1), torch.tril(H0))
    # an illegal boundary is refused, not rounded
    with pytest.raises(ValueError):
        check(lib.bnnp_ggn_full_accum_rows(net.ops, net.n_ops, C.byref(pl), N, ptr(ws), ptr(Lam), ptr(H1), ptr(sc), 2,
                                           0, 100, 1, stream()))
    net.release()


@pytest.mark.parametrize('engine,tol', [('trsm', 3e-5), ('cublas', 3e-5), ('tc', 5e-4)])
def test_blocked_spd_inverse_matches_potri(engine, tol, monkeypatch):
    """dense.spd_inverse_from_chol (blocked triangular inverse + blocked Gram, used for the one-off factorisation above
    dense.MIN_P) against torch.cholesky_inverse on a moderately conditioned SPD matrix; odd size so that the recursion
    and the block loops see ragged tails.  Stated bounds: fp32-library engines 3e-5 relative, split-bf16 engine 5e-4."""
    from preds_b200 import dense
    P = 9001
    g = torch.Generator(device='cuda').manual_seed(0)
    A = torch.randn(P, 1500, device='cuda', generator=g)
    H = A @ A.T
    H.diagonal().add_(float(H.diagonal().mean()) * 0.02)
    L = torch.linalg.cholesky(H)
    want = torch.cholesky_inverse(L, upper=False)
    monkeypatch.setattr(dense, 'ENGINE', engine)
    got = dense.spd_inverse_from_chol(L)
    assert torch.equal(got, got.T)
    assert cases.rel_err(got.cpu(), want.cpu()) < tol
    # below MIN_P the library call is used unchanged
    small = torch.linalg.cholesky(H[:300, :300].contiguous())
    assert cases.rel_err(dense.spd_inverse_from_chol(small).cpu(), torch.cholesky_inverse(small, upper=False).cpu()) < 1e-6


@pytest.mark.parametrize('dampen', [0, 1])
def test_kron_weight_sampling_tensor_core(dampen):
    """bnnp_kron_sample_group at a size that takes the tcgen05 branch (S*m*n*min(m,n) >= 2^27) against an fp64
    evaluation of preds/kron.py:94-109:  W1 ((W1^T eps W2) * D) W2^T  with D = sqrt(1/(l1 l2^T + delta)) (or the dampened
    form), written behind an offset into theta_s [S, P]."""
    from preds_b200 import _cabi as cabi
    m, n, S, off, delta = 192, 160, 40, 37, 0.7
    P = off + m * n + 11
    g = torch.Generator().manual_seed(21)
    W1 = torch.linalg.qr(torch.randn(m, m, generator=g, dtype=torch.float64))[0]
    W2 = torch.linalg.qr(torch.randn(n, n, generator=g, dtype=torch.float64))[0]
    l1, l2 = torch.rand(m, generator=g, dtype=torch.float64) * 5, torch.rand(n, generator=g, dtype=torch.float64) * 3
    eps = torch.randn(S, m, n, generator=g, dtype=torch.float64)
    mu = torch.randn(P, generator=g, dtype=torch.float64)
    if dampen:
        D = torch.sqrt(1 / ((l1 + delta ** 0.5).unsqueeze(1) * (l2 + delta ** 0.5).unsqueeze(0)))
    else:
        D = torch.sqrt(1 / (l1.unsqueeze(1) * l2.unsqueeze(0) + delta))
    want = W1 @ ((W1.T @ eps @ W2) * D) @ W2.T + mu[off:off + m * n].reshape(m, n)
    dev = 'cuda'
    f = lambda t: t.to(dev, torch.float32).contiguous()
    out = torch.zeros(S, P, device=dev)
    scratch = torch.empty(2 * S * m * n + m * n + 64, device=dev)
    dW1, dl1, dW2, dl2, deps, dmu = (f(t) for t in (W1, l1, W2, l2, eps, mu))      # keep the device copies alive
    cabi.check(cabi.lib.bnnp_kron_sample_group(cabi.ptr(dW1), cabi.ptr(dl1), cabi.ptr(dW2), cabi.ptr(dl2), m, n, delta,
                                               dampen, cabi.ptr(deps), S, cabi.ptr(dmu), cabi.ptr(out), P, off,
                                               cabi.ptr(scratch), cabi.stream()))
    got = out[:, off:off + m * n].reshape(S, m, n)
    assert cases.rel_err(got.cpu(), want) < 2e-5
    assert float(out[:, :off].abs().max()) == 0 and float(out[:, off + m * n:].abs().max()) == 0     # nothing else touched
