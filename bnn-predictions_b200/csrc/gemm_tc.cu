// Non-template parts of the generic tensor-core GEMM + the plain-matrix entry point.
#include "gemm_tc.cuh"

namespace bnnp {
namespace tcg {

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int ksplit, long long M, long long N,
                                     int batch, float alpha, float beta, float* __restrict__ C, long long ldc,
                                     long long bsc) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = M * N, tot = per * batch;
  if (i >= tot) return;
  float acc = 0.f;
  for (int k = 0; k < ksplit; ++k) acc += ws[(long long)k * tot + i];
  const long long b = i / per, r = (i % per) / N, c = i % N;
  float* q = C + b * bsc + r * ldc + c;
  *q = (beta == 0.f) ? alpha * acc : alpha * acc + beta * (*q);
}

}  // namespace tcg
}  // namespace bnnp

using namespace bnnp;
using namespace bnnp::tcg;

extern "C" int64_t bnnp_sgemm_tc_scratch_floats(int64_t M, int64_t N, int64_t K) {
  return (int64_t)pick_ksplit(M, N, K, 1) * M * N + 64;
}

extern "C" int bnnp_sgemm_tc(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                             int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
                             float* scratch, void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  const int ks = scratch ? pick_ksplit(M, N, K, 1) : 1;
  if (!ta && !tb) return gemm_tc_splitk(M, N, K, 1, TcRowMajor{A, lda, 0}, TcColMajor{B, ldb, 0}, alpha, beta, C, ldc, 0, scratch, ks, st);
  if (!ta && tb) return gemm_tc_splitk(M, N, K, 1, TcRowMajor{A, lda, 0}, TcRowMajor{B, ldb, 0}, alpha, beta, C, ldc, 0, scratch, ks, st);
  if (ta && !tb) return gemm_tc_splitk(M, N, K, 1, TcColMajor{A, lda, 0}, TcColMajor{B, ldb, 0}, alpha, beta, C, ldc, 0, scratch, ks, st);
  return gemm_tc_splitk(M, N, K, 1, TcColMajor{A, lda, 0}, TcRowMajor{B, ldb, 0}, alpha, beta, C, ldc, 0, scratch, ks, st);
}
