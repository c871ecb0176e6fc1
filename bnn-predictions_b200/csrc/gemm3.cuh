// Three-term split-bf16 tensor-core GEMM of the one-off factorisation (gemm3.cu).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace bnnp {
// C[M,N] = alpha * sum_k A[m,k] B[n,k] + beta * C.  A / B: three bf16 planes (hi, mid, lo; `pstride` elements apart), rows
// K-major with pitch lda / ldb (multiples of 8), 16-byte aligned.  K <= 4096 per call (walk longer K in segments, beta = 1).
// lower: only tiles on/below the diagonal (M == N).  ktri = 1 / 2: the K loop of a tile starts at its column / row origin
// + koff (operands that are triangular matrices); tiles without work leave C untouched when beta == 1.
// own_mod > 1: only the tiles whose global column (own_dim 0) / row (own_dim 1) tile index (local + own_off) is congruent to
// own_rank modulo own_mod are computed; all others are left untouched (work shared between the GPUs of one job).
int gemm3_nt(int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* A, int64_t lda, int64_t a_pstride,
             const __nv_bfloat16* B, int64_t ldb, int64_t b_pstride, float beta, float* C, int64_t ldc, int lower, int ktri,
             int64_t koff, cudaStream_t st, int own_mod = 1, int own_rank = 0, int own_dim = 0, int64_t own_off = 0);
// planes <- scale * src (or its transpose), split into hi / mid / lo
int split3(const float* src, int64_t lds, int64_t rows, int64_t cols, __nv_bfloat16* planes, int64_t pitch, int64_t pstride,
           int transposed, float scale, cudaStream_t st);
}  // namespace bnnp
