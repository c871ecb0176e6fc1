// Likelihood terms of preds/likelihoods.py as device kernels (one thread per example for K <= 32; wider outputs go
// to the warp-per-example kernels of wide_k.cu).
#include "common.cuh"

namespace bnnp {

constexpr float kClampEps = 1e-7f;     // preds/likelihoods.py:4

__device__ __forceinline__ void softmax_row(const float* f, int K, float* p) {
  float mx = f[0];
  for (int k = 1; k < K; ++k) mx = fmaxf(mx, f[k]);
  float s = 0.f;
  for (int k = 0; k < K; ++k) { p[k] = expf(f[k] - mx); s += p[k]; }
  float inv = 1.f / s;
  for (int k = 0; k < K; ++k) p[k] *= inv;
}

__global__ void lik_hessian_kernel(int lik, const float* __restrict__ f, int64_t N, int K,
                                   float sigma2, float* __restrict__ L) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float* out = L + n * K * K;
  if (lik == BNNP_LIK_CATEGORICAL) {
    float p[BNNP_MAX_K];
    softmax_row(f + n * K, K, p);
    for (int k = 0; k < K; ++k) p[k] = fminf(fmaxf(p[k], kClampEps), 1.f - kClampEps);
    for (int k = 0; k < K; ++k)
      for (int l = 0; l < K; ++l) out[k * K + l] = (k == l ? p[k] : 0.f) - p[k] * p[l];
  } else if (lik == BNNP_LIK_GAUSSIAN) {
    out[0] = 1.f / sigma2;
  } else {
    float p = 1.f / (1.f + expf(-f[n]));
    p = fminf(fmaxf(p, kClampEps), 1.f - kClampEps);
    out[0] = p * (1.f - p);
  }
}

// seed[n, v, k] = S_n[k, v];  CE: S = diag(sqrt p) - p sqrt(p)^T (unclamped), MSE: sqrt(2) I
__global__ void lik_sqrt_seed_kernel(int lik, const float* __restrict__ f, int64_t N, int K,
                                     float* __restrict__ seed) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float* out = seed + n * K * K;
  if (lik == BNNP_LIK_CATEGORICAL) {
    float p[BNNP_MAX_K];
    softmax_row(f + n * K, K, p);
    for (int v = 0; v < K; ++v) {
      float sq = sqrtf(p[v]);
      for (int k = 0; k < K; ++k) out[v * K + k] = (k == v ? sq : 0.f) - p[k] * sq;
    }
  } else {
    for (int v = 0; v < K; ++v)
      for (int k = 0; k < K; ++k) out[v * K + k] = (k == v) ? 1.41421356237309515f : 0.f;
  }
}

__global__ void identity_seed_kernel(int64_t N, int K, float* __restrict__ seed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K * K) return;
  int k = (int)(i % K), v = (int)((i / K) % K);
  seed[i] = (k == v) ? 1.f : 0.f;
}

__global__ void lik_residual_kernel(int lik, const float* __restrict__ f, const void* __restrict__ y,
                                    int64_t N, int K, float sigma2, float* __restrict__ rs) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (lik == BNNP_LIK_CATEGORICAL) {
    float p[BNNP_MAX_K];
    softmax_row(f + n * K, K, p);
    int64_t label = reinterpret_cast<const int64_t*>(y)[n];
    for (int k = 0; k < K; ++k) rs[n * K + k] = (k == label ? 1.f : 0.f) - p[k];
  } else if (lik == BNNP_LIK_GAUSSIAN) {
    const float* yy = reinterpret_cast<const float*>(y);
    for (int k = 0; k < K; ++k) rs[n * K + k] = (yy[n * K + k] - f[n * K + k]) / sigma2;
  } else {
    const float* yy = reinterpret_cast<const float*>(y);
    for (int k = 0; k < K; ++k) rs[n * K + k] = yy[n * K + k] - 1.f / (1.f + expf(-f[n * K + k]));
  }
}

__global__ void lik_inv_link_kernel(int lik, const float* f, int64_t rows, int K, float* out) {   // out may alias f
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= rows) return;
  if (lik == BNNP_LIK_CATEGORICAL) {
    float p[BNNP_MAX_K];
    softmax_row(f + n * K, K, p);
    for (int k = 0; k < K; ++k) out[n * K + k] = p[k];
  } else if (lik == BNNP_LIK_GAUSSIAN) {
    for (int k = 0; k < K; ++k) out[n * K + k] = f[n * K + k];
  } else {
    for (int k = 0; k < K; ++k) out[n * K + k] = 1.f / (1.f + expf(-f[n * K + k]));
  }
}

// linear_regression_predictive epilogue (preds/predictives.py:79-94): with Lam = Hessian(f) * sigma2 (K x K,
// symmetric), mu_star = inv_link(f) + Lam (f_mu - f), var_f = Lam f_var Lam, var_noise = Lam * sigma2.
__global__ void lik_link_moments_kernel(int lik, const float* __restrict__ f, const float* __restrict__ fmu,
                                        const float* __restrict__ fvar, float sigma2, int64_t N, int K,
                                        float* __restrict__ mu_star, float* __restrict__ var_f,
                                        float* __restrict__ var_noise) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float p[BNNP_MAX_K], lam[BNNP_MAX_K * BNNP_MAX_K], t[BNNP_MAX_K * BNNP_MAX_K];
  if (lik == BNNP_LIK_CATEGORICAL) {
    float pc[BNNP_MAX_K];
    softmax_row(f + n * K, K, p);
    for (int k = 0; k < K; ++k) pc[k] = fminf(fmaxf(p[k], kClampEps), 1.f - kClampEps);
    for (int k = 0; k < K; ++k)
      for (int l = 0; l < K; ++l) lam[k * K + l] = ((k == l ? pc[k] : 0.f) - pc[k] * pc[l]) * sigma2;
  } else if (lik == BNNP_LIK_GAUSSIAN) {
    p[0] = f[n];
    lam[0] = (1.f / sigma2) * sigma2;
  } else {
    p[0] = 1.f / (1.f + expf(-f[n]));
    const float pc = fminf(fmaxf(p[0], kClampEps), 1.f - kClampEps);
    lam[0] = pc * (1.f - pc) * sigma2;
  }
  for (int l = 0; l < K; ++l) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(fmu[n * K + k] - f[n * K + k], lam[k * K + l], acc);
    mu_star[n * K + l] = p[l] + acc;
  }
  const float* V = fvar + n * K * K;
  for (int k = 0; k < K; ++k)           // t = f_var Lam
    for (int l = 0; l < K; ++l) {
      float acc = 0.f;
      for (int c = 0; c < K; ++c) acc = fmaf(V[k * K + c], lam[c * K + l], acc);
      t[k * K + l] = acc;
    }
  for (int k = 0; k < K; ++k)           // var_f = Lam^T t
    for (int l = 0; l < K; ++l) {
      float acc = 0.f;
      for (int c = 0; c < K; ++c) acc = fmaf(lam[c * K + k], t[c * K + l], acc);
      var_f[(n * K + k) * K + l] = acc;
    }
  for (int e = 0; e < K * K; ++e) var_noise[n * K * K + e] = lam[e] * sigma2;
}

}  // namespace bnnp
using namespace bnnp;

static int check_lik(int lik, int K) {
  if (lik < 0 || lik > 2 || K <= 0 || K > BNNP_MAX_K_WIDE) return BNNP_ERR_INVALID;
  if (K > BNNP_MAX_K && lik != BNNP_LIK_CATEGORICAL) return BNNP_ERR_INVALID;      // wide outputs: classification only
  if (lik != BNNP_LIK_CATEGORICAL && K != 1 && lik == BNNP_LIK_BERNOULLI) return BNNP_ERR_INVALID;
  return BNNP_OK;
}

extern "C" int bnnp_lik_hessian(int lik, const float* f, int64_t N, int K, float sigma2,
                                float* Lambda, void* stream) {
  if (check_lik(lik, K) || !f || !Lambda || N <= 0) return BNNP_ERR_INVALID;
  if (lik != BNNP_LIK_CATEGORICAL && K != 1) return BNNP_ERR_INVALID;   // preds/likelihoods.py:51
  if (K > BNNP_MAX_K) return lik_wide(0, f, nullptr, N, K, Lambda, as_stream(stream));
  lik_hessian_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, as_stream(stream)>>>(lik, f, N, K, sigma2, Lambda);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" int bnnp_lik_sqrt_seed(int lik, const float* f, int64_t N, int K, float* seed, void* stream) {
  if (check_lik(lik, K) || !f || !seed || N <= 0) return BNNP_ERR_INVALID;
  if (lik == BNNP_LIK_BERNOULLI) return BNNP_ERR_INVALID;               // no nn_loss, :74-75
  if (K > BNNP_MAX_K) return lik_wide(1, f, nullptr, N, K, seed, as_stream(stream));
  lik_sqrt_seed_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, as_stream(stream)>>>(lik, f, N, K, seed);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" int bnnp_identity_seed(int64_t N, int K, float* seed, void* stream) {
  if (!seed || N <= 0 || K <= 0) return BNNP_ERR_INVALID;
  identity_seed_kernel<<<(unsigned)ceil_div64(N * K * K, 256), 256, 0, as_stream(stream)>>>(N, K, seed);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" int bnnp_lik_residual(int lik, const float* f, const void* y, int64_t N, int K,
                                 float sigma2, float* rs, void* stream) {
  if (check_lik(lik, K) || !f || !y || !rs || N <= 0) return BNNP_ERR_INVALID;
  if (K > BNNP_MAX_K) return lik_wide(2, f, y, N, K, rs, as_stream(stream));
  lik_residual_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, as_stream(stream)>>>(lik, f, y, N, K, sigma2, rs);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" int bnnp_lik_inv_link(int lik, const float* f, int64_t rows, int K, float* out, void* stream) {
  if (check_lik(lik, K) || !f || !out || rows <= 0) return BNNP_ERR_INVALID;
  if (K > BNNP_MAX_K) return lik_wide(3, f, nullptr, rows, K, out, as_stream(stream));
  lik_inv_link_kernel<<<(unsigned)ceil_div64(rows, 128), 128, 0, as_stream(stream)>>>(lik, f, rows, K, out);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" int bnnp_lik_link_moments(int lik, const float* f, const float* f_mu, const float* f_var, float sigma2,
                                     int64_t N, int K, float* mu_star, float* var_f, float* var_noise,
                                     void* stream) {
  if (check_lik(lik, K) || !f || !f_mu || !f_var || !mu_star || !var_f || !var_noise || N <= 0) return BNNP_ERR_INVALID;
  if (lik != BNNP_LIK_CATEGORICAL && K != 1) return BNNP_ERR_INVALID;
  if (K > BNNP_MAX_K) return link_moments_wide(f, f_mu, f_var, sigma2, N, K, mu_star, var_f, var_noise, as_stream(stream));
  lik_link_moments_kernel<<<(unsigned)ceil_div64(N, 64), 64, 0, as_stream(stream)>>>(lik, f, f_mu, f_var, sigma2, N, K,
                                                                                    mu_star, var_f, var_noise);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}
