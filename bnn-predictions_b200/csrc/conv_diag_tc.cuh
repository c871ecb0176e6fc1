// Conv-weight diagonal GGN on TMA-fed tensor cores (conv_diag_tc.cu).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace bnnp {

// implicit im2col of a conv input: U(pos, e) with e = (ci, a, b), pos = (oh, ow); batch b -> example b / V
struct ConvUnfold {
  const float* in; int64_t ild; int V;
  int Cin, H, W, OW, kh, kw, sh, sw, pl, pt;
  __device__ float at(int64_t n, int pos, int e) const {
    int b = e % kw, a = (e / kw) % kh, ci = e / (kw * kh);
    int oh = pos / OW, ow = pos % OW;
    int ih = oh * sh - pt + a, iw = ow * sw - pl + b;
    if (ih < 0 || ih >= H || iw < 0 || iw >= W) return 0.f;
    return in[n * ild + ((int64_t)ci * H + ih) * W + iw];
  }
  // the same with the index arithmetic hoisted: (ci, a, b) of e and (oh, ow) of pos are decomposed once by the caller and
  // advanced incrementally -- `at` costs eight integer divisions per element, which made the unfold passes compute-bound
  // at 0.5 TB/s of stores
  __device__ __forceinline__ float at_dec(const float* __restrict__ xn, int ci, int a, int b, int oh, int ow) const {
    const int ih = oh * sh - pt + a, iw = ow * sw - pl + b;
    if (ih < 0 || ih >= H || iw < 0 || iw >= W) return 0.f;
    return __ldg(xn + (ci * H + ih) * W + iw);
  }
};

// d_w[co, e] += factor * sum_{r < N*V} (sum_pos G[r*gld + co*L + pos] * unfold(in[r / V])[e, pos])^2
bool conv_diag_tma_ok(int Cout, int L, long long N, int V);
int64_t conv_diag_tma_scratch_floats(int64_t N, int V, int Cout, int L, int E);
int conv_diag_tma(const ConvUnfold& u, const float* G, int64_t gld, int64_t N, int V, int Cout, int L, int E, float factor,
                  float* d_w, float* scratch, cudaStream_t st, const __nv_bfloat16* pre_hi = nullptr,
                  const __nv_bfloat16* pre_lo = nullptr);
// pre_hi / pre_lo: the gradient already as bf16 hi / lo planes [(r, co), pos] with row pitch L (L % 8 == 0); G is then unused

}  // namespace bnnp
