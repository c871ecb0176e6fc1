// Likelihood terms and the GLM sampling epilogue for WIDE outputs (BNNP_MAX_K < K <= BNNP_MAX_K_WIDE, e.g. the 100
// classes of CIFAR100Net, preds/models.py:182-236).  The narrow versions (lik.cu, glm.cu) keep one example's K x K
// quantities in the registers of one thread; here one WARP owns an example (class probabilities spread over the lanes,
// K x K outputs written coalesced) and one BLOCK owns an input of the sampler (its K x K Cholesky factor lives in
// shared memory).  Same arithmetic, same Philox stream (counter = (s*N + n, k/4), four normals per call).
#include "common.cuh"

namespace bnnp {
namespace wide {

constexpr float kClampEps = 1e-7f;     // preds/likelihoods.py:4
constexpr int PER_LANE = BNNP_MAX_K_WIDE / 32;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// softmax of row f[0..K) spread over the lanes: p[j] belongs to class lane + 32 j
__device__ __forceinline__ void softmax_lanes(const float* __restrict__ f, int K, int lane, float* p) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < PER_LANE; ++j) {
    const int k = lane + 32 * j;
    p[j] = k < K ? f[k] : -INFINITY;
    mx = fmaxf(mx, p[j]);
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PER_LANE; ++j) {
    const int k = lane + 32 * j;
    p[j] = k < K ? expf(p[j] - mx) : 0.f;
    s += p[j];
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
#pragma unroll
  for (int j = 0; j < PER_LANE; ++j) p[j] *= inv;
}

// mode 0: Hessian diag(p) - p p^T on clamped p   [N,K,K]
// mode 1: sqrt seed  seed[n, v, k] = (k == v) sqrt(p_v) - p_k sqrt(p_v)  (unclamped, BackPACK's CE Hessian)   [N,K,K]
// mode 2: residual onehot(y) - p   [N,K]        mode 3: inv_link = p   [N,K] (out may alias f)
__global__ void __launch_bounds__(128)
lik_wide_kernel(int mode, const float* f, const int64_t* __restrict__ y, int64_t N, int K, float* out) {
  __shared__ float ps[4][BNNP_MAX_K_WIDE];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 4 + wid;
  if (n >= N) return;
  float p[PER_LANE];
  softmax_lanes(f + n * K, K, lane, p);
#pragma unroll
  for (int j = 0; j < PER_LANE; ++j) {
    const int k = lane + 32 * j;
    if (k < K) ps[wid][k] = mode == 0 ? fminf(fmaxf(p[j], kClampEps), 1.f - kClampEps) : p[j];
  }
  __syncwarp();
  const float* pw = ps[wid];
  if (mode == 0) {
    float* o = out + n * K * K;
    for (int e = lane; e < K * K; e += 32) {
      const int k = e / K, l = e % K;
      o[e] = (k == l ? pw[k] : 0.f) - pw[k] * pw[l];
    }
  } else if (mode == 1) {
    float* o = out + n * K * K;
    for (int e = lane; e < K * K; e += 32) {
      const int v = e / K, k = e % K;
      const float sq = sqrtf(pw[v]);
      o[e] = (k == v ? sq : 0.f) - pw[k] * sq;
    }
  } else if (mode == 2) {
    const int64_t label = y[n];
    for (int k = lane; k < K; k += 32) out[n * K + k] = (k == label ? 1.f : 0.f) - pw[k];
  } else {
    for (int k = lane; k < K; k += 32) out[n * K + k] = pw[k];
  }
}

// linear_regression_predictive epilogue for wide K (categorical only): one block per input,
// shared memory: lam [K,K], t [K,K], p [K], d [K]
__global__ void __launch_bounds__(128)
link_moments_wide_kernel(const float* __restrict__ f, const float* __restrict__ fmu, const float* __restrict__ fvar,
                         float sigma2, int64_t N, int K, float* __restrict__ mu_star, float* __restrict__ var_f,
                         float* __restrict__ var_noise) {
  extern __shared__ float sh[];
  float* lam = sh;
  float* t = lam + K * K;
  float* p = t + K * K;
  float* d = p + K;
  const int64_t n = blockIdx.x;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    float pl[PER_LANE];
    softmax_lanes(f + n * K, K, lane, pl);
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j) {
      const int k = lane + 32 * j;
      if (k < K) { p[k] = pl[j]; d[k] = fmu[n * K + k] - f[n * K + k]; }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K * K; e += blockDim.x) {
    const int k = e / K, l = e % K;
    const float pk = fminf(fmaxf(p[k], kClampEps), 1.f - kClampEps), pl = fminf(fmaxf(p[l], kClampEps), 1.f - kClampEps);
    lam[e] = ((k == l ? pk : 0.f) - pk * pl) * sigma2;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < K; l += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(d[k], lam[k * K + l], acc);
    mu_star[n * K + l] = p[l] + acc;
  }
  const float* V = fvar + n * K * K;
  for (int e = threadIdx.x; e < K * K; e += blockDim.x) {       // t = f_var Lam
    const int k = e / K, l = e % K;
    float acc = 0.f;
    for (int c = 0; c < K; ++c) acc = fmaf(V[k * K + c], lam[c * K + l], acc);
    t[e] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K * K; e += blockDim.x) {       // var_f = Lam^T t
    const int k = e / K, l = e % K;
    float acc = 0.f;
    for (int c = 0; c < K; ++c) acc = fmaf(lam[c * K + k], t[c * K + l], acc);
    var_f[n * K * K + e] = acc;
    var_noise[n * K * K + e] = lam[e] * sigma2;
  }
}

// ---- sampling epilogue -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
  *hi = __umulhi(a, b);
  return a * b;
}
__device__ __forceinline__ uint4 philox4(uint4 ctr, uint2 key) {      // Philox4x32-10, as in glm.cu
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0, hi1;
    uint32_t lo0 = mulhilo32(0xD2511F53u, ctr.x, &hi0);
    uint32_t lo1 = mulhilo32(0xCD9E8D57u, ctr.z, &hi1);
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
  float u2 = (float)b * 2.3283064365386963e-10f;
  float r = sqrtf(-2.f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}
// the k-th standard normal of draw (s, n): injected, or the Philox stream of glm_sample_kernel
__device__ __forceinline__ float draw(const float* __restrict__ eps, uint64_t seed, int64_t sn, int K, int k) {
  if (eps) return eps[sn * K + k];
  uint4 r = philox4(make_uint4((uint32_t)sn, (uint32_t)((uint64_t)sn >> 32), (uint32_t)(k / 4), 0u),
                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
  const int c = k & 3;
  return c == 0 ? a.x : (c == 1 ? a.y : (c == 2 ? b.x : b.y));
}

// One block per input n: lower Cholesky factor of f_var[n] (from its lower triangle) in shared memory, then every warp
// walks its share of the block's sample slab: f = f_mu + L eps, inverse link, store (or sum for the fused mean).
// diag != 0: the reference's whole-batch fallback (preds/predictives.py:73-76) -- Normal(f_mu, diag(f_var).clamp(1e-5))
// with the variance used as the scale.
template <bool MEAN>
__global__ void __launch_bounds__(128)
glm_sample_wide_kernel(int lik, const float* __restrict__ fmu, const float* __restrict__ fvar,
                       const float* __restrict__ eps, uint64_t seed, int64_t S, int64_t N, int K,
                       float* __restrict__ out, int32_t* __restrict__ status, int s_per_block, int diag) {
  extern __shared__ float sh[];
  float* L = sh;                         // [K, K] row-major, lower triangle used
  float* mu = L + K * K;                 // [K]
  float* ev = mu + K;                    // [4 warps][K] draws of the warp's current sample
  float* fv = ev + 4 * K;                // [4 warps][K]
  float* accs = fv + 4 * K;              // [K] block-level sum of the samples (MEAN)
  __shared__ int bad;
  const int64_t n = blockIdx.x;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  if (tid == 0) bad = 0;
  for (int e = tid; e < K * K; e += blockDim.x) L[e] = fvar[n * K * K + e];
  for (int k = tid; k < K; k += blockDim.x) { mu[k] = fmu[n * K + k]; accs[k] = 0.f; }
  __syncthreads();
  if (!diag) {
    // right-looking Cholesky, column by column (row-major lower triangle)
    for (int j = 0; j < K; ++j) {
      if (tid == 0) {
        const float s = L[j * K + j];
        if (!(s > 0.f)) bad = 1;
        L[j * K + j] = sqrtf(s);
      }
      __syncthreads();
      const float dj = L[j * K + j];
      for (int i = j + 1 + tid; i < K; i += blockDim.x) L[i * K + j] /= dj;
      __syncthreads();
      // trailing update of the lower triangle: L[i, c] -= L[i, j] * L[c, j]  for j < c <= i
      const int rem = K - j - 1;
      for (int e = tid; e < rem * rem; e += blockDim.x) {
        const int i = j + 1 + e / rem, c = j + 1 + e % rem;
        if (c <= i) L[i * K + c] -= L[i * K + j] * L[c * K + j];
      }
      __syncthreads();
    }
    if (blockIdx.y == 0 && tid == 0 && status) status[n] = bad;
  }
  const int64_t s0 = (int64_t)blockIdx.y * s_per_block;
  const int64_t s1 = s0 + s_per_block < S ? s0 + s_per_block : S;
  float* e = ev + wid * K;
  float* fw = fv + wid * K;
  float acc[PER_LANE];
#pragma unroll
  for (int j = 0; j < PER_LANE; ++j) acc[j] = 0.f;
  for (int64_t s = s0 + wid; s < s1; s += 4) {
    const int64_t sn = s * N + n;
    for (int k = lane; k < K; k += 32) e[k] = draw(eps, seed, sn, K, k);
    __syncwarp();
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) {
      float v = mu[k];
      if (diag) {
        v = fmaf(fmaxf(fvar[(n * K + k) * K + k], 1e-5f), e[k], v);
      } else {
        for (int l = 0; l <= k; ++l) v = fmaf(L[k * K + l], e[l], v);
      }
      fw[k] = v;
      mx = fmaxf(mx, v);
    }
    if (lik == BNNP_LIK_CATEGORICAL) {
      mx = warp_max(mx);
      float sum = 0.f;
      for (int k = lane; k < K; k += 32) { const float t = __expf(fw[k] - mx); fw[k] = t; sum += t; }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      for (int k = lane; k < K; k += 32) fw[k] *= inv;
    } else if (lik == BNNP_LIK_BERNOULLI) {
      for (int k = lane; k < K; k += 32) fw[k] = 1.f / (1.f + __expf(-fw[k]));
    }
    if (MEAN) {
#pragma unroll
      for (int j = 0; j < PER_LANE; ++j) { const int k = lane + 32 * j; if (k < K) acc[j] += fw[k]; }
    } else {
      for (int k = lane; k < K; k += 32) out[sn * K + k] = fw[k];
    }
    __syncwarp();
  }
  if (MEAN) {
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j) { const int k = lane + 32 * j; if (k < K) atomicAdd(&accs[k], acc[j]); }
    __syncthreads();
    const float invS = 1.f / (float)S;
    for (int k = tid; k < K; k += blockDim.x) atomicAdd(out + n * K + k, accs[k] * invS);
  }
}

__global__ void zero_kernel(float* p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

}  // namespace wide

int lik_wide(int mode, const float* f, const void* y, int64_t N, int K, float* out, cudaStream_t st) {
  wide::lik_wide_kernel<<<(unsigned)ceil_div64(N, 4), 128, 0, st>>>(mode, f, reinterpret_cast<const int64_t*>(y), N, K, out);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

int link_moments_wide(const float* f, const float* fmu, const float* fvar, float sigma2, int64_t N, int K, float* mu_star,
                      float* var_f, float* var_noise, cudaStream_t st) {
  const size_t sh = (size_t)(2 * K * K + 2 * K) * sizeof(float);
  BNNP_CUDA_CHECK(cudaFuncSetAttribute(wide::link_moments_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  wide::link_moments_wide_kernel<<<(unsigned)N, 128, sh, st>>>(f, fmu, fvar, sigma2, N, K, mu_star, var_f, var_noise);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

// mean != 0: out is [N,K] (zeroed here), else [S,N,K]; diag != 0: the diagonal fallback (status may be NULL)
int glm_sample_wide(int lik, const float* fmu, const float* fvar, const float* eps, uint64_t seed, int64_t S, int64_t N,
                    int K, float* out, int32_t* status, int mean, int diag, cudaStream_t st) {
  // enough sample slabs to fill the machine a few times over, at least 8 samples per slab
  int64_t want = ceil_div64(148 * 8, N);
  int64_t slabs = want < 1 ? 1 : want;
  if (slabs > ceil_div64(S, 8)) slabs = ceil_div64(S, 8);
  if (slabs < 1) slabs = 1;
  const int spb = (int)ceil_div64(S, slabs);
  slabs = ceil_div64(S, spb);
  if (slabs > 65535 || N > 0x7fffffff) return BNNP_ERR_UNSUPPORTED;
  const size_t sh = (size_t)(K * K + K + 8 * K + K) * sizeof(float);
  if (mean) {
    wide::zero_kernel<<<(unsigned)ceil_div64(N * K, 256), 256, 0, st>>>(out, N * K);
    BNNP_LAUNCH_CHECK();
    BNNP_CUDA_CHECK(cudaFuncSetAttribute(wide::glm_sample_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    wide::glm_sample_wide_kernel<true><<<dim3((unsigned)N, (unsigned)slabs), 128, sh, st>>>(lik, fmu, fvar, eps, seed, S, N, K,
                                                                                         out, status, spb, diag);
  } else {
    BNNP_CUDA_CHECK(cudaFuncSetAttribute(wide::glm_sample_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    wide::glm_sample_wide_kernel<false><<<dim3((unsigned)N, (unsigned)slabs), 128, sh, st>>>(lik, fmu, fvar, eps, seed, S, N, K,
                                                                                          out, status, spb, diag);
  }
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

}  // namespace bnnp
