"""The one-off factorisation chain of the full-covariance posterior on the library's own tensor-core GEMM
(preds/laplace.py:43-45: ``Chol = cholesky(H)``, ``Sigma = cholesky_inverse(Chol)``, ``Sigma_chol = cholesky(Sigma)``).

Every O(P^3) product runs on ``bnnp_gemm3_nt`` (csrc/gemm3.cu): three-term bf16 split, six MMAs per product, chunked main
accumulation -- the error class of an fp32 library GEMM at ~5x its rate.  Only the NB x NB diagonal blocks (a Cholesky and
a triangular inverse of a 1024 x 1024 matrix per panel, O(P NB^2) in total) go through the library (cuSOLVER / cuBLAS
through torch), as the north star allows for the factorisation.

  cholesky_lower(A)       right-looking blocked Cholesky, in place: panel L21 = A21 L11^-T as a GEMM with the inverted
                          diagonal block, trailing update A22 -= L21 L21^T on the lower tiles only.
  inverse_from_chol(L)    Y = L^-T is built block column by block column *as bf16 planes* (it only ever enters products):
                          Y[:i0, I] = -(Y[:i0, :i0] L[I, :i0]^T) L_II^-T, skipping the zero part of the triangular operand
                          per tile; then Sigma = Y Y^T on the lower tiles, K walked in 4096-wide segments.
"""
import torch

from ._cabi import lib, check, ptr, stream

NB = 1024            # panel width = K of every trailing update (one correction accumulation)
KSEG = 4096          # longest K of one bnnp_gemm3_nt call
MIN_P = 4096         # below this the library calls are used as they are
ENGINE = 'tc'        # 'library': cuSOLVER / cuBLAS through torch whatever the size (A/B timing, scripts/factor_check.py)



class _Planes:
    """Three bf16 planes (hi, mid, lo) of a [rows, cols] matrix, rows k-contiguous; the pitch is a multiple of 64 elements so
    that every row starts on a 128-byte line (with 8-element granularity the 64-byte rows of a TMA box straddle sectors and
    the L2 -> SM traffic, already at 72 % of its peak in this kernel, grows by a third: measured 850 vs 1270 TFLOP/s)."""

    def __init__(self, rows, cols, device, zero=False):
        self.rows, self.cols, self.pitch = rows, cols, (cols + 63) // 64 * 64
        self.pstride = rows * self.pitch
        alloc = torch.zeros if zero else torch.empty
        self.buf = alloc(3 * self.pstride, dtype=torch.bfloat16, device=device)

    def addr(self, r=0, c=0):
        return self.buf.data_ptr() + 2 * (r * self.pitch + c)

    def fill(self, src_addr, ld, rows, cols, r=0, c=0, transposed=False, scale=1.0):
        """planes[r:r+rows, c:c+cols] <- scale * src (transposed: planes[r:r+cols, c:c+rows] <- scale * src^T)."""
        check(lib.bnnp_split3_bf16(src_addr, ld, rows, cols, self.addr(r, c), self.pitch, self.pstride, int(transposed),
                                   scale, stream()), 'bnnp_split3_bf16')


def _gemm3(M, N, K, alpha, A, a_addr, B, b_addr, beta, c_addr, ldc, lower=0, ktri=0, koff=0, own=None):
    """own = (world, rank, dim, offset): only the tiles of C whose global column (dim 0) / row (dim 1) tile index is
    congruent to rank modulo world are computed (the chain shared between the ranks of a data-parallel job)."""
    if own is None or own[0] <= 1:
        check(lib.bnnp_gemm3_nt(M, N, K, alpha, a_addr, A.pitch, A.pstride, b_addr, B.pitch, B.pstride, beta, c_addr, ldc,
                                lower, ktri, koff, stream()), 'bnnp_gemm3_nt')
    else:
        check(lib.bnnp_gemm3_nt_owned(M, N, K, alpha, a_addr, A.pitch, A.pstride, b_addr, B.pitch, B.pstride, beta, c_addr,
                                      ldc, lower, ktri, koff, own[0], own[1], own[2], own[3], stream()),
              'bnnp_gemm3_nt_owned')


TILE = 128           # tile edge of bnnp_gemm3_nt: the unit of ownership when the chain is shared between ranks


def _group(shared):
    """(dist, rank, world) of the default process group when the chain is to be shared between its ranks."""
    if shared:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def _own_cols_mask(c0, ncols, rank, world, device):
    """1.0 for the columns c0 .. c0+ncols-1 (global indices) whose 128-wide tile column belongs to ``rank``."""
    t = (torch.arange(c0, c0 + ncols, device=device) // TILE) % world
    return (t == rank).to(torch.float32)


def _addr(t, r, c):
    return t.data_ptr() + 4 * (r * t.stride(0) + c)


def _tri_inv(Lii):
    eye = torch.eye(Lii.shape[0], dtype=torch.float32, device=Lii.device)
    return torch.linalg.solve_triangular(Lii, eye, upper=False).contiguous()      # (the library hands back column-major)


def cholesky_lower(A, shared=False):
    """In-place lower Cholesky factor of the symmetric positive definite A [P,P] (fp32, contiguous; only the lower triangle
    is read); the strict upper triangle of the result is zero.  Raises like torch.linalg.cholesky when A is not SPD.

    ``shared``: every rank of the default process group calls this with the same A.  The tiles of the trailing update are
    dealt out by global tile column (rank = column tile mod world), so a rank's copy is always complete on its own tiles;
    after each update the NEXT panel column -- the only part anyone else needs -- is gathered (an all-reduce of a buffer
    holding each rank's own 128-column strips and zeros elsewhere: 240 MB at P = 59 635).  Every tile is computed by the
    same kernel on the same inputs as in the single-process chain, so every rank ends with exactly its bits."""
    P = A.shape[0]
    assert A.is_contiguous() and A.dtype == torch.float32
    dev = A.device
    dist, rank, world = _group(shared)
    panel = _Planes(max(P - NB, 1), NB, dev)
    wpl = _Planes(NB, NB, dev)
    infos = []
    for j0 in range(0, P, NB):
        nb = min(NB, P - j0)
        D = A[j0:j0 + nb, j0:j0 + nb]
        L11, info = torch.linalg.cholesky_ex(D)
        infos.append(info)
        D.copy_(L11)
        if j0 + nb >= P:
            break
        m = P - j0 - nb
        W = _tri_inv(L11)
        # panel: L21[r, n] = sum_k A21[r, k] W[n, k]
        panel.fill(_addr(A, j0 + nb, j0), P, m, nb)
        wpl.fill(W.data_ptr(), nb, nb, nb)
        _gemm3(m, nb, nb, 1.0, panel, panel.addr(), wpl, wpl.addr(), 0.0, _addr(A, j0 + nb, j0), P)
        # trailing update, lower tiles: A22 -= L21 L21^T
        panel.fill(_addr(A, j0 + nb, j0), P, m, nb)
        j1 = j0 + nb
        _gemm3(m, m, nb, -1.0, panel, panel.addr(), panel, panel.addr(), 1.0, _addr(A, j1, j1), P, lower=1,
               own=(world, rank, 0, j1 // TILE))
        if dist is not None:
            nb1 = min(NB, P - j1)
            col = A[j1:, j1:j1 + nb1]
            stage = col * _own_cols_mask(j1, nb1, rank, world, dev)
            dist.all_reduce(stage)
            col.copy_(stage)
            del stage
    if bool(torch.stack(infos).ne(0).any()) or not bool(torch.isfinite(A.diagonal()).all()):      # the only sync
        raise torch.linalg.LinAlgError('cholesky_lower: the matrix is not positive definite')
    A.tril_()
    return A


def inverse_from_chol(L, shared=False):
    """Sigma = (L L^T)^-1 [P,P] (symmetric, both triangles filled) from the lower Cholesky factor L.

    ``shared`` (every rank holds the same L): the rows of Y = L^-T are independent of each other -- row n of a new block
    column needs row n of the earlier ones only -- so they are dealt out by 128-row tile and built without any
    communication; one all-reduce of the planes (disjoint supports: a gather) completes Y everywhere, then the tiles of
    Sigma = Y Y^T are dealt out by tile column and gathered the same way.  Same kernels on the same inputs per tile as the
    single-process chain: identical bits."""
    P = L.shape[0]
    assert L.is_contiguous() and L.dtype == torch.float32
    dev = L.device
    dist, rank, world = _group(shared)
    by_row, by_col = (world, rank, 1, 0), (world, rank, 0, 0)
    diag_blocks = []
    Y = _Planes(P, P, dev, zero=True)                      # Y = L^-T (upper triangular), only ever an operand
    lrow = _Planes(NB, max(P - 1, 1), dev)
    tpl = _Planes(max(P - 1, 1), NB, dev)
    xpl = _Planes(NB, NB, dev)
    T = torch.empty(max(P - 1, 1) * NB, dtype=torch.float32, device=dev)
    for i0 in range(0, P, NB):
        nb = min(NB, P - i0)
        Xii = _tri_inv(L[i0:i0 + nb, i0:i0 + nb])
        Y.fill(Xii.data_ptr(), nb, nb, nb, r=i0, c=i0, transposed=True)
        if dist is not None:
            diag_blocks.append((i0, nb, Xii))
        if i0 == 0:
            continue
        # Tt[n, m] = sum_{n <= k < i0} Y[n, k] L[i0 + m, k]
        lrow.fill(_addr(L, i0, 0), P, nb, i0)
        Tt = T[:i0 * nb].view(i0, nb)
        Tt.zero_()
        for k0 in range(0, i0, KSEG):
            kk = min(KSEG, i0 - k0)
            rows = min(i0, k0 + kk)                        # rows n >= k0 + kk have no entry in this segment
            _gemm3(rows, nb, kk, 1.0, Y, Y.addr(0, k0), lrow, lrow.addr(0, k0), 1.0, Tt.data_ptr(), nb, ktri=2, koff=-k0,
                   own=by_row)
        # Y[:i0, I] = -Tt Xii^T :  out[n, m'] = -sum_m Tt[n, m] Xii[m', m]
        tpl.fill(Tt.data_ptr(), nb, i0, nb)
        xpl.fill(Xii.data_ptr(), nb, nb, nb)
        # (rows of other ranks: Tt stays zero there, and so do their planes)
        _gemm3(i0, nb, nb, -1.0, tpl, tpl.addr(), xpl, xpl.addr(), 0.0, Tt.data_ptr(), nb, own=by_row)
        Y.fill(Tt.data_ptr(), nb, i0, nb, r=0, c=i0)
    del lrow, tpl, xpl, T
    if dist is not None:
        dist.all_reduce(Y.buf)                             # disjoint row supports: exact in bf16
        for i0, nb, Xii in diag_blocks:                    # (every rank had written every diagonal block)
            Y.fill(Xii.data_ptr(), nb, nb, nb, r=i0, c=i0, transposed=True)
        del diag_blocks
    S = torch.zeros(P, P, dtype=torch.float32, device=dev)
    for k0 in range(0, P, KSEG):
        kk = min(KSEG, P - k0)
        rows = min(P, k0 + kk)
        _gemm3(rows, rows, kk, 1.0, Y, Y.addr(0, k0), Y, Y.addr(0, k0), 1.0, S.data_ptr(), P, lower=1, ktri=2, koff=-k0,
               own=by_col)
    del Y
    if dist is not None:
        dist.all_reduce(S)                                 # disjoint tile supports: a gather
    check(lib.bnnp_symmetrize_lower(ptr(S), P, stream()), 'bnnp_symmetrize_lower')
    return S


def use_tc(P):
    return ENGINE == 'tc' and P >= MIN_P


def cholesky(A, shared=False):
    """Lower Cholesky factor of A as a new tensor (torch.linalg.cholesky semantics; A is left untouched)."""
    if use_tc(A.shape[0]):
        return cholesky_lower(A.clone(memory_format=torch.contiguous_format), shared=shared)
    return torch.linalg.cholesky(A)


def posterior_from_precision(H, shared=False):
    """(Sigma_chol, Sigma) of preds/laplace.py:43-45 from the precision H: chol(H) -> Sigma = H^-1 -> chol(Sigma).
    ``shared``: all ranks of the default process group hold the same H and split the work (identical bits everywhere)."""
    if not use_tc(H.shape[0]):
        from . import dense
        chol = torch.linalg.cholesky(H)
        Sigma = dense.spd_inverse_from_chol(chol, sharded=shared)
        del chol
        return torch.linalg.cholesky(Sigma), Sigma
    chol = cholesky(H, shared=shared)
    Sigma = inverse_from_chol(chol, shared=shared)
    del chol
    return cholesky(Sigma, shared=shared), Sigma
