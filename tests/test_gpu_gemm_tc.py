"""The generic tensor-core GEMM (tcgen05, split-bf16, generated operands) against fp64 matmul:
all four transpose combinations, ragged sizes (tile tails in M, N and K), split-K and beta != 0.
Bound: 2e-5 relative to max |C| (split-bf16 carries ~2^-17 per term; fp32 SIMT gives ~1e-6)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 256, 64), (1, 1, 1), (130, 17, 33), (300, 520, 1000), (64, 64, 20000), (1000, 75, 784),
          (513, 257, 95), (2048, 10, 75)]


@pytest.mark.parametrize('ta,tb', [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize('M,N,K', SHAPES)
def test_sgemm_tc(M, N, K, ta, tb):
    from preds_b200._cabi import lib, check, ptr, stream
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N * 3 + K + ta * 2 + tb)
    A = torch.randn((K, M) if ta else (M, K), device='cuda', generator=g)
    B = torch.randn((N, K) if tb else (K, N), device='cuda', generator=g)
    C0 = torch.randn(M, N, device='cuda', generator=g)
    Aop = A.T if ta else A
    Bop = B.T if tb else B
    for alpha, beta, use_scratch in [(1.0, 0.0, False), (0.5, 2.0, True)]:
        C = C0.clone()
        sc = None
        if use_scratch:
            sc = torch.empty(int(lib.bnnp_sgemm_tc_scratch_floats(M, N, K)), device='cuda')
        check(lib.bnnp_sgemm_tc(ta, tb, M, N, K, alpha, ptr(A), A.shape[1], ptr(B), B.shape[1], beta, ptr(C), N,
                                ptr(sc), stream()), 'bnnp_sgemm_tc')
        ref = alpha * (Aop.double() @ Bop.double()) + beta * C0.double()
        err = float((C.double() - ref).abs().max() / ref.abs().max())
        # without scratch a single TMEM accumulator sums all of K and its truncating adds show
        # (~3e-9 relative per K element); with scratch split-K caps each accumulation at 4096
        bound = 2e-5 if (use_scratch or K <= 4096) else 2e-4
        assert err < bound, (M, N, K, ta, tb, alpha, beta, err)


def test_sgemm_simt_matches_too():
    from preds_b200._cabi import lib, check, ptr, stream
    A = torch.randn(200, 300, device='cuda')
    B = torch.randn(300, 150, device='cuda')
    C = torch.zeros(200, 150, device='cuda')
    check(lib.bnnp_sgemm(0, 0, 200, 150, 300, 1.0, ptr(A), 300, ptr(B), 150, 0.0, ptr(C), 150, stream()))
    assert float((C - A @ B).abs().max()) < 1e-3


@pytest.mark.parametrize('shape', [(300, 200, 1000), (129, 17, 9000), (10, 10, 40960), (1024, 1024, 5000), (77, 513, 33)])
def test_gemm_nt_split_tma(shape):
    """bnnp_gemm_nt_split (TMA-fed, operands pre-split into bf16 hi/lo by bnnp_split_bf16 / bnnp_split_bf16_t) against
    fp64: ragged M / N / K tails, K > 4096 (split-K partials, one TMEM accumulation <= 4096 k), alpha / beta."""
    from preds_b200._cabi import lib, check, ptr, stream
    M, N, K = shape
    g = torch.Generator().manual_seed(7)
    A = torch.randn(M, K, generator=g).cuda()
    Bt = torch.randn(K, N, generator=g).cuda()                # B given transposed: exercises bnnp_split_bf16_t
    C = torch.randn(M, N, generator=g).cuda()
    Kp = (K + 7) // 8 * 8
    ah, al = (torch.empty(M, Kp, dtype=torch.bfloat16, device='cuda') for _ in range(2))
    bh, bl = (torch.empty(N, Kp, dtype=torch.bfloat16, device='cuda') for _ in range(2))
    check(lib.bnnp_split_bf16(ptr(A), K, M, K, ptr(ah), ptr(al), stream()))
    check(lib.bnnp_split_bf16_t(ptr(Bt), N, K, N, Kp, ptr(bh), ptr(bl), stream()))
    sc = torch.empty(int(lib.bnnp_gemm_nt_scratch_floats(M, N, K)), device='cuda')
    want = 0.7 * (A.double() @ Bt.double()) - 0.3 * C.double()
    check(lib.bnnp_gemm_nt_split(M, N, K, 0.7, ptr(ah), ptr(al), Kp, ptr(bh), ptr(bl), Kp, -0.3, ptr(C), N, ptr(sc), stream()))
    assert float((C.double() - want).abs().max() / want.abs().max()) < 2e-5
