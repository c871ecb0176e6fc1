"""Dense one-off factorisation helpers for the full-covariance posterior (preds/laplace.py:43-45).

``cholesky_inverse`` (cuSOLVER ``potri``) is the slow third of ``chol -> inverse -> chol``: 7.1 s of 10.2 s at
P = 59 635, against 1.5 s for each Cholesky.  ``spd_inverse_from_chol`` computes the same Sigma = L^-T L^-1 from
GEMM-rich pieces that run on the split-bf16 tensor-core GEMM of the library (``bnnp_sgemm_tc``, fp32-grade):

  1. L^-1 by recursive block inversion  [[A,0],[B,C]]^-1 = [[A^-1,0],[-C^-1 B A^-1, C^-1]]  (leaf blocks by trsm);
  2. Sigma = L^-T L^-1 block by block, skipping the zero blocks of the triangular factor, lower triangle only,
     mirrored at the end.
"""
import torch

from ._cabi import lib, check, ptr, stream

# Engine of the blocked inverse (measured at P = 59 635 against potri: 7.07 s, residual |H S - I|_max 2.2e-4):
#   'trsm'   (default) off-diagonal blocks by two cuBLAS triangular solves against the factor, Gram by fp32 library GEMMs
#            accumulated in 4096-row segments: 2.9 s, residual 3.2e-4, 9e-6 relative difference to potri;
#   'cublas' off-diagonal blocks as products with the already inverted sub-blocks (fp32 library GEMMs): 3.7 s;
#   'tc'     the same products on the library's split-bf16 tensor-core GEMM: 1.5 s, but the ~2^-17 product error is
#            amplified by the conditioning (residual 1.8e-3, 7.7e-5 relative difference at P = 12 000) -- opt-in only.
# (A single fp32 accumulation over all P rows in the Gram step costs 10x in the residual: 2.5e-3.)
ENGINE = 'trsm'
LEAF = 2048          # leaf size of the recursive triangular inverse (solved by cuBLAS trsm)
BLOCK = 4096         # block size of the triangular product
MIN_P = 8192         # below this the library call is used as is


KSEG = 4096          # one tensor-core accumulation never spans more (TMEM adds truncate); segments add through C


def _gemm(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    """C = alpha op(A) op(B) + beta C with the K range walked in KSEG segments (no split-K workspace)."""
    for k0 in range(0, K, KSEG):
        kk = min(KSEG, K - k0)
        a = A + 4 * (k0 * lda if ta else k0)            # op(A) is M x K: stored K x M when transposed
        b = B + 4 * (k0 if tb else k0 * ldb)            # op(B) is K x N: stored N x K when transposed
        check(lib.bnnp_sgemm_tc(ta, tb, M, N, kk, alpha, a, lda, b, ldb, beta if k0 == 0 else 1.0, Cm, ldc, None, stream()),
              'bnnp_sgemm_tc')


def _addr(t, r, c):
    return t.data_ptr() + (r * t.stride(0) + c) * 4


def tri_inverse_lower(L):
    """X = L^-1 for a lower-triangular L [P,P] (row-major, contiguous); the strict upper triangle of X is zero."""
    P = L.shape[0]
    L = L.contiguous()
    X = torch.zeros(P, P, dtype=torch.float32, device=L.device)
    half = (P // 2 + 127) // 128 * 128
    T = torch.empty(max(1, (P - half) * half), dtype=torch.float32, device=L.device)
    ld = P

    def rec(lo, hi):
        n = hi - lo
        if n <= LEAF:
            eye = torch.eye(n, dtype=torch.float32, device=L.device)
            X[lo:hi, lo:hi] = torch.linalg.solve_triangular(L[lo:hi, lo:hi], eye, upper=False)
            return
        mid = lo + (n // 2 + 127) // 128 * 128
        rec(lo, mid)
        rec(mid, hi)
        n1, n2 = mid - lo, hi - mid
        # T = X22 @ L21 ;  X21 = -T @ X11
        if ENGINE == 'trsm':
            # X21 = -L22^-1 L21 L11^-1 by two triangular solves against the factor itself (no computed inverse enters:
            # the errors of the sub-blocks do not compound)
            Tv = torch.linalg.solve_triangular(L[mid:hi, mid:hi], L[mid:hi, lo:mid], upper=False, left=True)
            X[mid:hi, lo:mid] = torch.linalg.solve_triangular(L[lo:mid, lo:mid], Tv, upper=False, left=False).neg_()
            return
        if ENGINE == 'cublas':
            Tv = T[:n2 * n1].view(n2, n1)
            torch.mm(X[mid:hi, mid:hi], L[mid:hi, lo:mid], out=Tv)
            X[mid:hi, lo:mid] = torch.mm(Tv, X[lo:mid, lo:mid]).neg_()
            return
        _gemm(0, 0, n2, n1, n2, 1.0, _addr(X, mid, mid), ld, _addr(L, mid, lo), ld, 0.0, T.data_ptr(), n1)
        _gemm(0, 0, n2, n1, n1, -1.0, T.data_ptr(), n1, _addr(X, lo, lo), ld, 0.0, _addr(X, mid, lo), ld)

    rec(0, P)
    return X


def gram_of_lower(X):
    """S = X^T X for a lower-triangular X [P,P]: S[I,J] = sum_{K >= I} X[K,I]^T X[K,J] for block columns J <= I (the blocks
    above the diagonal of X are zero and never touched); the upper triangle of S is mirrored from the lower one."""
    P = X.shape[0]
    S = torch.empty(P, P, dtype=torch.float32, device=X.device)
    ld = P
    for i0 in range(0, P, BLOCK):
        bi = min(BLOCK, P - i0)
        for j0 in range(0, i0 + 1, BLOCK):
            bj = min(BLOCK, P - j0)
            # rows i0.. of X: A = X[i0:, i0:i0+bi] (k x bi, used transposed), B = X[i0:, j0:j0+bj]
            if ENGINE in ('cublas', 'trsm'):
                # K walked in segments whose partial products are added in fp32: a single fp32 accumulation over up to
                # P = 59 635 terms loses ~10x more than potri's blocked lauum (measured on |H S - I|)
                out = S[i0:i0 + bi, j0:j0 + bj]
                for k0 in range(i0, P, KSEG):
                    k1 = min(P, k0 + KSEG)
                    part = torch.mm(X[k0:k1, i0:i0 + bi].T, X[k0:k1, j0:j0 + bj])
                    if k0 == i0:
                        out.copy_(part)
                    else:
                        out.add_(part)
                continue
            _gemm(1, 0, bi, bj, P - i0, 1.0, _addr(X, i0, i0), ld, _addr(X, i0, j0), ld, 0.0, _addr(S, i0, j0), ld)
    check(lib.bnnp_symmetrize_lower(ptr(S), P, stream()), 'bnnp_symmetrize_lower')
    return S


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def spd_inverse_sharded(L):
    """The same inverse with the work SHARED by the ranks of the default process group (every rank holds the same L after
    the exchange of the accumulators, and would otherwise repeat the same 2.9 s): block columns are dealt out cyclically,

      1. X[:, J] = L^-1 E_J      one triangular solve against the trailing part of L per owned block column J;
      2. all ranks sum their column sets (one all-reduce of X: disjoint supports, so the sum is a gather);
      3. S[I, J] = sum_{K >= I} X[K, I]^T X[K, J]  for the owned block columns J, I >= J (the blocks above the diagonal of
         X are zero and never touched), K walked in 4096-row segments added in fp32;
      4. all ranks sum their block columns of S (again a gather), lower triangle mirrored.

    Cyclic dealing balances the (P - j0)^2-shaped cost of a column.  Every rank ends with identical bits."""
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    P = L.shape[0]
    L = L.contiguous()
    cols = [(j0, min(BLOCK, P - j0)) for j0 in range(0, P, BLOCK)]
    X = torch.zeros(P, P, dtype=torch.float32, device=L.device)
    for j, (j0, bj) in enumerate(cols):
        if j % world != rank:
            continue
        rhs = torch.zeros(P - j0, bj, dtype=torch.float32, device=L.device)
        rhs[:bj].diagonal().fill_(1.0)
        X[j0:, j0:j0 + bj] = torch.linalg.solve_triangular(L[j0:, j0:], rhs, upper=False)
        del rhs
    dist.all_reduce(X)
    S = torch.zeros(P, P, dtype=torch.float32, device=L.device)
    for j, (j0, bj) in enumerate(cols):
        if j % world != rank:
            continue
        for (i0, bi) in cols[j:]:
            out = S[i0:i0 + bi, j0:j0 + bj]
            for k0 in range(i0, P, KSEG):
                k1 = min(P, k0 + KSEG)
                part = torch.mm(X[k0:k1, i0:i0 + bi].T, X[k0:k1, j0:j0 + bj])
                if k0 == i0:
                    out.copy_(part)
                else:
                    out.add_(part)
    del X
    dist.all_reduce(S)
    check(lib.bnnp_symmetrize_lower(ptr(S), P, stream()), 'bnnp_symmetrize_lower')
    return S


def spd_inverse_from_chol(L, sharded=False):
    """Sigma = (L L^T)^-1 from the lower Cholesky factor; equals torch.cholesky_inverse(L) up to fp32 rounding.
    ``sharded``: every rank of the default process group calls this with the same L and the work is split between them
    (``spd_inverse_sharded``)."""
    if sharded and L.shape[0] >= MIN_P and _dist() is not None:
        return spd_inverse_sharded(L)
    if L.shape[0] < MIN_P:
        S = torch.cholesky_inverse(L, upper=False)
        return S.T if (not S.is_contiguous() and S.T.is_contiguous()) else S
    X = tri_inverse_lower(L)
    return gram_of_lower(X)
