// Generic fp32 SIMT GEMM with functor operands:  acc(m,n) = sum_k A(m,k) * B(k,n).
//
// Operands are *generated* by functors (so Jacobian tiles g (x) a never touch HBM) and the
// result is consumed by an epilogue functor.  Used for every small / irregular contraction of
// the path; the large contractions go through the tcgen05 kernels (gram_tc.cu).
//
//   struct LA { static constexpr bool K_CONTIG; __device__ float operator()(int64_t m, int64_t k, int b) const; };
//   struct LB { static constexpr bool K_CONTIG; __device__ float operator()(int64_t k, int64_t n, int b) const; };
//   struct EP { __device__ bool skip(int64_t m0, int64_t n0, int bm, int bn) const;
//               __device__ void operator()(int64_t m, int64_t n, float acc, int b) const; };
#pragma once
#include "common.cuh"

namespace bnnp {

constexpr int G_BM = 64, G_BN = 64, G_BK = 16, G_TM = 4, G_TN = 4;
constexpr int G_THREADS = (G_BM / G_TM) * (G_BN / G_TN);   // 256

template <class LA, class LB, class EP>
__global__ void __launch_bounds__(G_THREADS)
gemm_simt_kernel(int64_t M, int64_t N, int64_t K, LA la, LB lb, EP ep) {
  __shared__ float As[G_BK][G_BM + 4];
  __shared__ float Bs[G_BK][G_BN + 4];
  const int b = blockIdx.z;
  const int64_t m0 = (int64_t)blockIdx.y * G_BM, n0 = (int64_t)blockIdx.x * G_BN;
  if (ep.skip(m0, n0, G_BM, G_BN)) return;
  const int t = threadIdx.x;
  const int tx = t % (G_BN / G_TN), ty = t / (G_BN / G_TN);
  float acc[G_TM][G_TN];
#pragma unroll
  for (int i = 0; i < G_TM; ++i)
#pragma unroll
    for (int j = 0; j < G_TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < K; k0 += G_BK) {
    // A tile: G_BM x G_BK
#pragma unroll
    for (int i = 0; i < (G_BM * G_BK) / G_THREADS; ++i) {
      int e = t + i * G_THREADS, mm, kk;
      if (LA::K_CONTIG) { kk = e % G_BK; mm = e / G_BK; } else { mm = e % G_BM; kk = e / G_BM; }
      int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? la(m, k, b) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (G_BN * G_BK) / G_THREADS; ++i) {
      int e = t + i * G_THREADS, nn, kk;
      if (LB::K_CONTIG) { kk = e % G_BK; nn = e / G_BK; } else { nn = e % G_BN; kk = e / G_BN; }
      int64_t n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < K) ? lb(k, n, b) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < G_BK; ++kk) {
      float a[G_TM], bb[G_TN];
#pragma unroll
      for (int i = 0; i < G_TM; ++i) a[i] = As[kk][ty * G_TM + i];
#pragma unroll
      for (int j = 0; j < G_TN; ++j) bb[j] = Bs[kk][tx * G_TN + j];
#pragma unroll
      for (int i = 0; i < G_TM; ++i)
#pragma unroll
        for (int j = 0; j < G_TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < G_TM; ++i)
#pragma unroll
    for (int j = 0; j < G_TN; ++j) {
      int64_t m = m0 + ty * G_TM + i, n = n0 + tx * G_TN + j;
      if (m < M && n < N) ep(m, n, acc[i][j], b);
    }
}

template <class LA, class LB, class EP>
inline int gemm_simt(int64_t M, int64_t N, int64_t K, int batch, LA la, LB lb, EP ep,
                     cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return BNNP_OK;
  dim3 grid((unsigned)ceil_div64(N, G_BN), (unsigned)ceil_div64(M, G_BM), (unsigned)batch);
  if (grid.y > 65535u || grid.z > 65535u) return BNNP_ERR_UNSUPPORTED;
  gemm_simt_kernel<LA, LB, EP><<<grid, G_THREADS, 0, st>>>(M, N, K, la, lb, ep);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

// ---- stock functors ---------------------------------------------------------------------
// element (r, c) of a strided matrix, optional batch stride
struct LoadRowMajorA {          // A(m,k) = p[b*bs + m*ld + k]
  static constexpr bool K_CONTIG = true;
  const float* p; int64_t ld, bs;
  __device__ float operator()(int64_t m, int64_t k, int b) const { return p[b * bs + m * ld + k]; }
};
struct LoadColMajorA {          // A(m,k) = p[b*bs + k*ld + m]   (transposed storage)
  static constexpr bool K_CONTIG = false;
  const float* p; int64_t ld, bs;
  __device__ float operator()(int64_t m, int64_t k, int b) const { return p[b * bs + k * ld + m]; }
};
struct LoadRowMajorB {          // B(k,n) = p[b*bs + k*ld + n]
  static constexpr bool K_CONTIG = false;
  const float* p; int64_t ld, bs;
  __device__ float operator()(int64_t k, int64_t n, int b) const { return p[b * bs + k * ld + n]; }
};
struct LoadColMajorB {          // B(k,n) = p[b*bs + n*ld + k]   (e.g. a Linear weight [out,in])
  static constexpr bool K_CONTIG = true;
  const float* p; int64_t ld, bs;
  __device__ float operator()(int64_t k, int64_t n, int b) const { return p[b * bs + n * ld + k]; }
};
struct StoreAxpby {             // C = alpha*acc + beta*C
  float* p; int64_t ld, bs; float alpha, beta;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t m, int64_t n, float acc, int b) const {
    float* q = p + b * bs + m * ld + n;
    *q = (beta == 0.f) ? alpha * acc : alpha * acc + beta * (*q);
  }
};

}  // namespace bnnp
