// Shared helpers for libbnnp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/bnnp_b200.h"

#define BNNP_CUDA_CHECK(expr)                                         \
  do {                                                                \
    cudaError_t e__ = (expr);                                         \
    if (e__ != cudaSuccess) {                                         \
      fprintf(stderr, "[bnnp_b200] CUDA error %s at %s:%d\n",         \
              cudaGetErrorString(e__), __FILE__, __LINE__);           \
      return BNNP_ERR_CUDA;                                           \
    }                                                                 \
  } while (0)

// every kernel launch of the library passes through here: counted for bench.py's gpu_launches
extern unsigned long long g_bnnp_launches;
#define BNNP_LAUNCH_CHECK()                 \
  do {                                      \
    ++g_bnnp_launches;                      \
    BNNP_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

// Engine-selection options (bnnp_set_option): process-wide, default 0, set explicitly by tests / A-B scripts -- never
// read from the environment, and none of them makes a kernel skip work.
struct BnnpOptions {
  int gram_kernel;     // full GGN on tensor cores: 0 auto (wide CTA-pair kernel), 1 single-CTA kernel, 2 two-unit CTA-pair kernel (BNNP_EXPERIMENTS builds)
  int glm_single;      // full-covariance GLM: 1 forces the single-CTA TMA kernel
  int glm_simt;        // kron / diag GLM contractions on the fp32 SIMT kernels with the unfused reductions
  int debug_sync;      // synchronise and report after every GLM TMA launch
  int net_simt;        // forward / backward / diag / KFAC contractions on the fp32 SIMT kernels whatever their size (parity bisection)
  int gemm_generated;  // 1: Linear-layer products / KFAC Grams on the generated-operand tensor-core GEMM instead of the TMA-fed one (A/B timing)
  int conv_bwd_explicit;  // 1: conv backward-data as product + col2im gather and conv forward as unfold + GEMM instead of the implicit GEMM, fp32 first-conv gradient + split pass in the diag path (A/B timing, parity)
};
extern BnnpOptions g_bnnp_opt;

namespace bnnp {
void glm_prof_set(bool on);
// wide-K variants (wide_k.cu): BNNP_MAX_K < K <= BNNP_MAX_K_WIDE
int lik_wide(int mode, const float* f, const void* y, int64_t N, int K, float* out, cudaStream_t st);
int link_moments_wide(const float* f, const float* fmu, const float* fvar, float sigma2, int64_t N, int K, float* mu_star,
                      float* var_f, float* var_noise, cudaStream_t st);
int glm_sample_wide(int lik, const float* fmu, const float* fvar, const float* eps, uint64_t seed, int64_t S, int64_t N,
                    int K, float* out, int32_t* status, int mean, int diag, cudaStream_t st);
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline bool is_param_op(const bnnp_op_t& o) {
  return o.kind == BNNP_OP_LINEAR || o.kind == BNNP_OP_CONV2D;
}
static inline bool is_elementwise_op(const bnnp_op_t& o) {
  return o.kind == BNNP_OP_RELU || o.kind == BNNP_OP_TANH || o.kind == BNNP_OP_SIGMOID;
}
// net.cu: first conv followed by ([activation,] max-pool) -- helpers shared with ggn.cu
int net_first_conv_possum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, int V, const float* ws,
                          float* S, cudaStream_t st);
bool net_first_conv_planes_structure_ok(const bnnp_op_t* ops, int n_ops);

static inline int64_t op_in_size(const bnnp_op_t& o) { return (int64_t)o.in_c * o.in_h * o.in_w; }
static inline int64_t op_out_size(const bnnp_op_t& o) { return (int64_t)o.out_c * o.out_h * o.out_w; }
