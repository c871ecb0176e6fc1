"""The factorisation chain on the library's own tensor-core GEMM (csrc/gemm3.cu, preds_b200/factor.py) against float64 and
against the library calls the reference makes (preds/laplace.py:43-45: torch.cholesky, cholesky_inverse, cholesky)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _planes(M, K):
    from preds_b200.factor import _Planes
    return _Planes(M, K, 'cuda', zero=True)


@pytest.mark.parametrize('shape', [(300, 200, 1000), (129, 17, 4096), (1024, 1024, 1024), (77, 513, 33), (640, 640, 257)])
def test_gemm3_nt_against_float64(shape):
    """bnnp_gemm3_nt: error relative to the result's scale no worse than twice that of an fp32 library GEMM (and < 1e-6);
    ragged M / N / K tails, alpha / beta."""
    from preds_b200.factor import _gemm3
    M, N, K = shape
    g = torch.Generator().manual_seed(11)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    C = torch.randn(M, N, generator=g).cuda()
    want = 0.7 * (A.double() @ B.double().T) - 0.3 * C.double()
    lib32 = 0.7 * (A @ B.T) - 0.3 * C
    pa, pb = _planes(M, K), _planes(N, K)
    pa.fill(A.data_ptr(), K, M, K)
    pb.fill(B.data_ptr(), K, N, K)
    _gemm3(M, N, K, 0.7, pa, pa.addr(), pb, pb.addr(), -0.3, C.data_ptr(), N)
    scale = float(want.abs().max())
    err = float((C.double() - want).abs().max()) / scale
    err32 = float((lib32.double() - want).abs().max()) / scale
    assert err < 1e-6 and err <= 2.0 * err32 + 1e-7, (err, err32)


def test_gemm3_positive_sums_are_unbiased():
    """Sums of one sign are where truncating accumulator adds show (scripts/tmem_trunc_probe.py: -8e-7 at K = 1024 in one
    accumulation): the chunked main accumulation keeps the mean signed error below 3e-8."""
    from preds_b200.factor import _gemm3
    M = N = 256
    K = 4096
    g = torch.Generator().manual_seed(3)
    A = (torch.rand(M, K, generator=g) + 0.5).cuda()
    B = (torch.rand(N, K, generator=g) + 0.5).cuda()
    C = torch.zeros(M, N, device='cuda')
    pa, pb = _planes(M, K), _planes(N, K)
    pa.fill(A.data_ptr(), K, M, K)
    pb.fill(B.data_ptr(), K, N, K)
    _gemm3(M, N, K, 1.0, pa, pa.addr(), pb, pb.addr(), 0.0, C.data_ptr(), N)
    want = A.double() @ B.double().T
    rel = (C.double() - want) / want
    assert abs(float(rel.mean())) < 3e-8 and float(rel.abs().max()) < 5e-7, (float(rel.mean()), float(rel.abs().max()))


def test_gemm3_lower_and_triangular_k():
    """lower = 1 touches only tiles on / below the diagonal; ktri = 2 skips the zero part of an upper-triangular operand
    (the Sigma = Y Y^T step), segment offsets included."""
    from preds_b200.factor import _gemm3
    n = 700
    g = torch.Generator().manual_seed(5)
    Y = torch.triu(torch.randn(n, n, generator=g)).cuda()
    C = torch.full((n, n), 7.0, device='cuda')
    py = _planes(n, n)
    py.fill(Y.data_ptr(), n, n, n)
    seg = 256
    C.zero_()
    C += 7.0
    for k0 in range(0, n, seg):
        kk = min(seg, n - k0)
        rows = min(n, k0 + kk)
        _gemm3(rows, rows, kk, 1.0, py, py.addr(0, k0), py, py.addr(0, k0), 1.0, C.data_ptr(), n, lower=1, ktri=2, koff=-k0)
    want = Y.double() @ Y.double().T + 7.0
    got = C.double()
    low = torch.tril(torch.ones(n, n, dtype=torch.bool, device='cuda'))
    assert float((got - want)[low].abs().max() / want.abs().max()) < 1e-6
    # tiles strictly above the diagonal were never written
    blk = torch.arange(n, device='cuda') // 128
    above = blk[:, None] < blk[None, :]
    assert bool((C[above] == 7.0).all())


@pytest.mark.parametrize('P', [4097, 4099, 6000])
def test_factorisation_chain_against_library(P):
    """cholesky_lower / inverse_from_chol against torch.linalg.cholesky / cholesky_inverse on an SPD matrix with a
    prior-dominated spectrum: the factors are no further from a float64 factorisation than 2x the library's, and so is the
    residual |H Sigma - I|.  P = 4097: a last panel of a single row / column (K = 1 products, 1 x 1 diagonal block)."""
    from preds_b200 import factor
    g = torch.Generator(device='cuda').manual_seed(1)
    A = torch.randn(P, 512, device='cuda', generator=g)
    H = A @ A.T
    H.diagonal().add_(float(H.diagonal().mean()) * 0.02)
    L0 = torch.linalg.cholesky(H)
    L64 = torch.linalg.cholesky(H.double())
    L1 = factor.cholesky_lower(H.clone())
    assert bool((torch.triu(L1, 1) == 0).all())
    e1 = float((L1 - L64).abs().max() / L64.abs().max())
    e0 = float((L0 - L64).abs().max() / L64.abs().max())
    assert e1 <= 2.0 * e0 + 1e-6 and e1 < 1e-4, (e1, e0)          # two fp32 factorisations differ by the conditioning
    S0 = torch.cholesky_inverse(L0)
    S1 = factor.inverse_from_chol(L1)
    assert torch.equal(S1, S1.T)
    eye = torch.eye(P, device='cuda', dtype=torch.float64)
    r0 = float((H.double() @ S0.double() - eye).abs().max())
    r1 = float((H.double() @ S1.double() - eye).abs().max())
    assert r1 <= 2.0 * r0 + 1e-6, (r1, r0)
    S64 = torch.cholesky_inverse(L64)
    e1 = float((S1 - S64).abs().max() / S64.abs().max())
    e0 = float((S0 - S64).abs().max() / S64.abs().max())
    assert e1 <= 2.0 * e0 + 1e-6 and e1 < 1e-3, (e1, e0)
    C0 = torch.linalg.cholesky(S0)
    C64 = torch.linalg.cholesky(S0.double())
    C1 = factor.cholesky(S0)          # (the library hands back a column-major Sigma)
    e1 = float((C1 - C64).abs().max() / C64.abs().max())
    e0 = float((C0 - C64).abs().max() / C64.abs().max())
    assert e1 <= 2.0 * e0 + 1e-6 and e1 < 1e-3, (e1, e0)


def test_cholesky_lower_rejects_indefinite():
    from preds_b200 import factor
    g = torch.Generator(device='cuda').manual_seed(2)
    A = torch.randn(4200, 4200, device='cuda', generator=g)
    H = A + A.T
    with pytest.raises(torch.linalg.LinAlgError):
        factor.cholesky_lower(H)
