// TMA-fed tensor-core GEMM on pre-split bf16 hi/lo operands (see gemm_tma.cu).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace bnnp {
// C[M,N] = alpha * sum_k A[m,k] B[n,k] + beta * C (+ bias[n], only without split-K partials).  A: [M, lda], B: [N, ldb] bf16
// hi / lo, k contiguous, lda / ldb multiples of 8, bases 16-byte aligned.  scratch: gemm_tma_scratch_floats() floats
// (split-K partials; may be NULL when K <= 4096).
int64_t gemm_tma_scratch_floats(int64_t M, int64_t N, int64_t K);
int gemm_tma_nt(int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* Ah, const __nv_bfloat16* Al, int64_t lda,
                const __nv_bfloat16* Bh, const __nv_bfloat16* Bl, int64_t ldb, float beta, float* C, int64_t ldc,
                float* scratch, cudaStream_t st, const float* bias = nullptr, int ct_L = 0, const float* had = nullptr,
                int64_t ldh = 0);
// conv backward-data product dUt[r, e, pos] = sum_co gout[r, co, pos] W[co, e] (gout row pitch gld, W [Cout, E]) with the
// gradients transposed + split once ([(r,pos), co] bf16 hi/lo) and the result stored transposed per row for the col2im gather
int64_t conv_bwd_tma_scratch_floats(int64_t R, int64_t Cout, int64_t L, int64_t E);
int conv_bwd_tma(const float* gout, int64_t gld, int64_t R, int Cout, int L, const float* W, int E, float* dUt,
                 float* scratch, cudaStream_t st);
// hi/lo[r, c] = split(src[r*lds + c]) (optionally squared first), row pitch `pitch` >= cols (multiple of 8), zero padded
int split_rows(const float* src, int64_t lds, int64_t rows, int64_t cols, int64_t pitch, __nv_bfloat16* hi,
               __nv_bfloat16* lo, bool square, cudaStream_t st);
// hi/lo[c, r] = split(src[r*lds + c]), row pitch `pitch` >= rows (multiple of 8), zero padded
int split_transposed(const float* src, int64_t lds, int64_t rows, int64_t cols, int64_t pitch, __nv_bfloat16* hi,
                     __nv_bfloat16* lo, bool square, cudaStream_t st);
// C[m,m] = alpha * S^T S + beta * C for S [R, m] (row pitch lds): transposed split into scratch, then one NT GEMM
int64_t gram_tma_scratch_floats(int64_t R, int64_t m);
int gram_tma(const float* S, int64_t lds, int64_t R, int64_t m, float alpha, float beta, float* C, int64_t ldc,
             float* scratch, cudaStream_t st);
// Y[R, n_out] = X[R, k] * op(W) (+ bias):  w_transposed == 0: W is [n_out, k] (row pitch ldw), Y = X W^T  (forward);
// w_transposed == 1: W is [k, n_out], Y = X W (backward-data).  scratch: linear_tma_scratch_floats().
int64_t linear_tma_scratch_floats(int64_t R, int64_t k, int64_t n_out);
// square_x != 0: X is squared element-wise before the split (the diagonal / Kronecker variance contractions).
int linear_tma(const float* X, int64_t ldx, int64_t R, int64_t k, const float* W, int64_t ldw, int64_t n_out,
               int w_transposed, const float* bias, float* Y, int64_t ldy, float* scratch, cudaStream_t st,
               int square_x = 0);
}  // namespace bnnp
