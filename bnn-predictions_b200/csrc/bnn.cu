// BNN predictive (preds/predictives.py:12-28), Kron.sample (preds/kron.py:85-110) and the small
// dense helpers of the posterior code.
#include <string.h>
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace bnnp {

struct StoreThetaT {        // theta_s[s, p] = mu[p] + acc   for acc(p, s)
  float* out; const float* mu; int64_t P;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t p, int64_t s, float acc, int) const { out[s * P + p] = mu[p] + acc; }
};
struct LoadLowerA {         // A(p, q) = L[p, q] for q <= p else 0 (Sigma_chol is lower triangular)
  static constexpr bool K_CONTIG = true;
  const float* L; int64_t P;
  __device__ float operator()(int64_t p, int64_t q, int) const { return q <= p ? L[p * P + q] : 0.f; }
};

__global__ void sample_theta_diag_kernel(const float* __restrict__ sigma, const float* __restrict__ mu,
                                         const float* __restrict__ eps, int64_t P, int64_t S,
                                         float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * S) return;
  int64_t p = i % P, s = i / P;
  out[s * P + p] = mu[p] + sigma[p] * eps[p * S + s];
}

__global__ void axpy_kernel(int64_t n, float alpha, const float* __restrict__ x, float* __restrict__ y) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaf(alpha, x[i], y[i]);
}

__global__ void scale_by_kernel(float* __restrict__ z, const float* __restrict__ D, int64_t per, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) z[i] *= D[i % per];
}
__global__ void kron_d_kernel(const float* __restrict__ l1, const float* __restrict__ l2, int m, int n,
                              float delta, int dampen, float* __restrict__ D) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)m * n) return;
  float a = l1[i / n], b = l2[i % n], v;
  if (dampen) { float sd = sqrtf(delta); v = 1.f / ((a + sd) * (b + sd)); }
  else v = 1.f / (a * b + delta);
  D[i] = sqrtf(v);
}
struct StoreThetaGroup {    // theta_s[s, off + i*n + j] = mu[...] + acc, batch s, (i, j)
  float* out; const float* mu; int64_t P, off; int n;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t i, int64_t j, float acc, int s) const {
    int64_t p = off + i * n + j;
    out[(int64_t)s * P + p] = (mu ? mu[p] : 0.f) + acc;
  }
};
// 1-factor: t[i, s] = sqrt(1/(l[i]+delta)) * sum_k W[k, i] eps[k, s];  out[s, off+i] = mu + sum_k W[i,k] t[k,s]
struct ScaleRowsStore {
  float* t; int64_t S; const float* l; float delta;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t i, int64_t s, float acc, int) const { t[i * S + s] = acc * sqrtf(1.f / (l[i] + delta)); }
};
struct StoreThetaVec {
  float* out; const float* mu; int64_t P, off;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t i, int64_t s, float acc, int) const { out[s * P + off + i] = (mu ? mu[off + i] : 0.f) + acc; }
};

// theta_s = mu + (L eps)^T on the tensor cores (zeros above the diagonal of the factor are generated, not read).
// The factor is the wide (N) operand: 256 of its rows per tile stream 32 KB per K-chunk through the generators (16 KB
// when it was the 128-row A operand), eps -- the S <= 128 samples -- is the A operand, and the output is written in its
// own [S, P] layout.  p_off: a K-segment [k0, k1) only concerns the rows p >= k0 of a lower-triangular factor.
struct TcLowerTriOff {
  static constexpr bool K_CONTIG = true;
  struct Row { const float* p; long long r; };
  const float* L; long long P, p_off;
  __device__ __forceinline__ Row row(int64_t r, int) const { return Row{L + (r + p_off) * P, r + p_off}; }
  __device__ __forceinline__ void load8(const Row& rw, int64_t k0, int64_t, float* out) const {
    if (k0 + 7 <= rw.r) {
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = __ldg(rw.p + k0 + e);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = (k0 + e <= rw.r) ? __ldg(rw.p + k0 + e) : 0.f;
    }
  }
};
struct TcStoreTheta : tcg::TcNoState {     // theta_s[s*P + p] = (first ? mu[p] : theta_s[s*P + p]) + acc(s, p - p_off)
  float* out; const float* mu; long long P, p_off; int first;
  __host__ __device__ TcStoreTheta(float* o_, const float* m_, long long P_, long long off_, int f_)
      : out(o_), mu(m_), P(P_), p_off(off_), first(f_) {}
  __device__ __forceinline__ void row(State&, int64_t s_, int64_t n0, const float* v, int nvalid, int, int) const {
    float* q = out + s_ * P + p_off + n0;
    const float* m = mu + p_off + n0;
    if (first) {
#pragma unroll
      for (int e = 0; e < 32; ++e) if (e < nvalid) q[e] = __ldg(m + e) + v[e];
    } else {
      float h[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) h[e] = e < nvalid ? q[e] : 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e) if (e < nvalid) q[e] = h[e] + v[e];
    }
  }
};

// Kron weight sampling on the tensor cores (preds/kron.py:94-109): epilogues of the four contractions
struct TcStoreScaledBy : tcg::TcNoState {      // T2[s][i, j] = acc * D[i, j]
  float* out; long long ld, bs; const float* D;
  __host__ __device__ TcStoreScaledBy(float* o_, long long ld_, long long bs_, const float* D_) : out(o_), ld(ld_), bs(bs_), D(D_) {}
  __device__ __forceinline__ void row(State&, int64_t i, int64_t j0, const float* v, int nvalid, int s, int) const {
    float* q = out + (long long)s * bs + i * ld + j0;
    const float* d = D + i * ld + j0;
#pragma unroll
    for (int e = 0; e < 32; ++e) if (e < nvalid) q[e] = v[e] * __ldg(d + e);
  }
};
struct TcStoreThetaGroup : tcg::TcNoState {    // theta_s[s, off + i*n + j] = mu[...] + acc
  float* out; const float* mu; long long P, off; int n;
  __host__ __device__ TcStoreThetaGroup(float* o_, const float* m_, long long P_, long long off_, int n_)
      : out(o_), mu(m_), P(P_), off(off_), n(n_) {}
  __device__ __forceinline__ void row(State&, int64_t i, int64_t j0, const float* v, int nvalid, int s, int) const {
    const long long p = off + i * n + j0;
    float* q = out + (long long)s * P + p;
#pragma unroll
    for (int e = 0; e < 32; ++e) if (e < nvalid) q[e] = (mu ? __ldg(mu + p + e) : 0.f) + v[e];
  }
};

// batched-weights Linear layer of the BNN predictive: out[s, n, :] = act(in[s or shared, n, :] W_s^T + b_s)
struct BnnStore {                 // SIMT epilogue
  float* out; int64_t ld, bs; const float* theta; int64_t P, offb; int has_bias, act;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t n, int64_t u, float acc, int s) const {
    float v = acc + (has_bias ? theta[(int64_t)s * P + offb + u] : 0.f);
    if (act == BNNP_OP_RELU) v = v > 0.f ? v : 0.f;
    else if (act == BNNP_OP_TANH) v = tanhf(v);
    else if (act == BNNP_OP_SIGMOID) v = 1.f / (1.f + expf(-v));
    out[(int64_t)s * bs + n * ld + u] = v;
  }
};
struct TcBnnStore : tcg::TcNoState {     // tensor-core epilogue
  float* out; long long ld, bs; const float* theta; long long P, offb; int has_bias, act;
  __host__ __device__ TcBnnStore(float* o_, long long ld_, long long bs_, const float* t_, long long P_, long long ob_,
                                 int hb_, int act_) : out(o_), ld(ld_), bs(bs_), theta(t_), P(P_), offb(ob_), has_bias(hb_), act(act_) {}
  __device__ __forceinline__ void row(State&, int64_t n, int64_t u0, const float* v, int nvalid, int s, int) const {
    float* q = out + (long long)s * bs + n * ld + u0;
    const float* b = theta + (long long)s * P + offb + u0;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      if (e < nvalid) {
        float x = v[e] + (has_bias ? __ldg(b + e) : 0.f);
        if (act == BNNP_OP_RELU) x = x > 0.f ? x : 0.f;
        else if (act == BNNP_OP_TANH) x = tanhf(x);
        else if (act == BNNP_OP_SIGMOID) x = 1.f / (1.f + expf(-x));
        q[e] = x;
      }
    }
  }
};

}  // namespace bnnp
using namespace bnnp;

extern "C" int bnnp_sample_theta_full(const float* L, const float* mu, const float* eps, int64_t P,
                                      int64_t S, float* theta_s, void* stream) {
  if (!L || !mu || !eps || !theta_s || P <= 0 || S <= 0) return BNNP_ERR_INVALID;
  if ((long long)P * P * S >= (1LL << 28))
    // K = P is long: sequential <= 4096-element segments accumulate through theta_s (TMEM adds truncate)
  {
    const int all = (int)((P + tcg::BK - 1) / tcg::BK);
    int seg = 0;
    for (int c0 = 0; c0 < all; c0 += tcg::kMaxChunksPerAccum, ++seg) {
      const long long p_off = (long long)c0 * tcg::BK;          // rows above the segment's first column are all zero there
      int rc = tcg::gemm_tc(S, P - p_off, P, 1, tcg::TcColMajor{eps, S, 0}, TcLowerTriOff{L, P, p_off},
                            TcStoreTheta(theta_s, mu, P, p_off, seg == 0), 1, as_stream(stream), c0,
                            c0 + tcg::kMaxChunksPerAccum);
      if (rc) return rc;
    }
    return BNNP_OK;
  }
  return gemm_simt(P, S, P, 1, LoadLowerA{L, P}, LoadRowMajorB{eps, S, 0}, StoreThetaT{theta_s, mu, P},
                   as_stream(stream));
}

extern "C" int bnnp_sample_theta_diag(const float* sigma, const float* mu, const float* eps, int64_t P,
                                      int64_t S, float* theta_s, void* stream) {
  if (!sigma || !mu || !eps || !theta_s || P <= 0 || S <= 0) return BNNP_ERR_INVALID;
  sample_theta_diag_kernel<<<(unsigned)ceil_div64(P * S, 256), 256, 0, as_stream(stream)>>>(sigma, mu, eps, P, S, theta_s);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" int bnnp_kron_sample_group(const float* W1, const float* l1, const float* W2, const float* l2,
                                      int m, int n, float delta, int dampen, const float* eps, int64_t S,
                                      const float* mu, float* theta_s, int64_t P, int64_t off,
                                      float* scratch, void* stream) {
  if (!W1 || !l1 || !eps || !theta_s || !scratch || m <= 0 || S <= 0) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  if (!W2) {
    // 1-factor (bias-like): eps [m, S]   preds/kron.py:88-93
    float* t = scratch;
    int rc = gemm_simt(m, S, m, 1, LoadColMajorA{W1, m, 0}, LoadRowMajorB{eps, S, 0},
                       ScaleRowsStore{t, S, l1, delta}, st);
    if (rc) return rc;
    return gemm_simt(m, S, m, 1, LoadRowMajorA{W1, m, 0}, LoadRowMajorB{t, S, 0},
                     StoreThetaVec{theta_s, mu, P, off}, st);
  }
  if (!l2 || n <= 0 || S > 65535) return BNNP_ERR_INVALID;
  // 2-factor: eps [S, m, n]   preds/kron.py:94-109
  const int64_t per = (int64_t)m * n;
  float* T1 = scratch;                  // [S, m, n]
  float* T2 = scratch + S * per;        // [S, m, n]
  float* D = T2 + S * per;              // [m, n]
  kron_d_kernel<<<(unsigned)ceil_div64(per, 256), 256, 0, st>>>(l1, l2, m, n, delta, dampen, D);
  BNNP_LAUNCH_CHECK();
  // T1 = eps W2 ; T2 = W1^T T1 ; T2 *= D ; T1 = T2 W2^T ; out = W1 T1
  if ((long long)S * m * n * (m < n ? m : n) >= (1LL << 27)) {
    // tensor cores: the W2 products for all samples at once ([S*m, n] x [n, n]), the W1 products batched over samples;
    // the D scaling rides in the epilogue of the second one
    int rc = tcg::gemm_tc(S * m, n, n, 1, tcg::TcRowMajor{eps, n, 0}, tcg::TcColMajor{W2, n, 0},
                          tcg::TcStoreAxpby(T1, n, 0, 1.f, 0.f), 1, st);
    if (rc) return rc;
    rc = tcg::gemm_tc(m, n, m, (int)S, tcg::TcColMajor{W1, m, 0}, tcg::TcColMajor{T1, n, per},
                      TcStoreScaledBy(T2, n, per, D), 1, st);
    if (rc) return rc;
    rc = tcg::gemm_tc(S * m, n, n, 1, tcg::TcRowMajor{T2, n, 0}, tcg::TcRowMajor{W2, n, 0},
                      tcg::TcStoreAxpby(T1, n, 0, 1.f, 0.f), 1, st);
    if (rc) return rc;
    return tcg::gemm_tc(m, n, m, (int)S, tcg::TcRowMajor{W1, m, 0}, tcg::TcColMajor{T1, n, per},
                        TcStoreThetaGroup(theta_s, mu, P, off, n), 1, st);
  }
  int rc = gemm_simt(m, n, n, (int)S, LoadRowMajorA{eps, n, per}, LoadRowMajorB{W2, n, 0},
                     StoreAxpby{T1, n, per, 1.f, 0.f}, st);
  if (rc) return rc;
  rc = gemm_simt(m, n, m, (int)S, LoadColMajorA{W1, m, 0}, LoadRowMajorB{T1, n, per},
                 StoreAxpby{T2, n, per, 1.f, 0.f}, st);
  if (rc) return rc;
  scale_by_kernel<<<(unsigned)ceil_div64(S * per, 256), 256, 0, st>>>(T2, D, per, S * per);
  BNNP_LAUNCH_CHECK();
  rc = gemm_simt(m, n, n, (int)S, LoadRowMajorA{T2, n, per}, LoadColMajorB{W2, n, 0},
                 StoreAxpby{T1, n, per, 1.f, 0.f}, st);
  if (rc) return rc;
  return gemm_simt(m, n, m, (int)S, LoadRowMajorA{W1, m, 0}, LoadRowMajorB{T1, n, per},
                   StoreThetaGroup{theta_s, mu, P, off, n}, st);
}

// ------------------------------------------------------------------------------------------
// forward passes at S sampled weight vectors
// ------------------------------------------------------------------------------------------
static bool bnn_linear_only(const bnnp_op_t* ops, int n_ops) {
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind != BNNP_OP_LINEAR && !is_elementwise_op(ops[i]) && ops[i].kind != BNNP_OP_FLATTEN) return false;
  return true;
}
// scratch: Linear-only networks keep two ping-pong activation buffers [S, N, max width]; others the
// workspace of one forward pass (+ logits)
extern "C" int64_t bnnp_bnn_forward_scratch_floats(const bnnp_op_t* ops, int n_ops, int64_t N, int64_t S) {
  if (bnn_linear_only(ops, n_ops)) {
    int64_t w = 0;
    for (int i = 0; i < n_ops; ++i) if (ops[i].kind == BNNP_OP_LINEAR && ops[i].out_c > w) w = ops[i].out_c;
    return 2 * S * N * w + 128;
  }
  bnnp_plan_t plan;
  if (bnnp_net_plan(ops, n_ops, N, 1, &plan) != BNNP_OK) return -1;
  return plan.total_floats + N * plan.K + 64;
}

extern "C" int bnnp_bnn_forward(const bnnp_op_t* ops, int n_ops, int lik, const float* X, int64_t N,
                                const float* theta_s, int64_t S, int64_t P, float* out, float* scratch,
                                void* stream) {
  if (!ops || !X || !theta_s || !out || !scratch || N <= 0 || S <= 0 || n_ops > BNNP_MAX_OPS) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  if (bnn_linear_only(ops, n_ops) && S <= 65535) {
    // all S weight samples at once: every Linear is a batched-weights GEMM (batch index = sample), the
    // activation that follows it is fused into the epilogue.  The model's own parameters are not touched.
    int64_t w = 0;
    for (int i = 0; i < n_ops; ++i) if (ops[i].kind == BNNP_OP_LINEAR && ops[i].out_c > w) w = ops[i].out_c;
    float* buf[2] = {scratch, scratch + S * N * w + 64};
    const float* in = X;
    int64_t in_ld = ops[0].in_c, in_bs = 0;       // the input batch is shared by all samples
    int pp = 0;
    int K = 0;
    for (int i = 0; i < n_ops; ++i) {
      const bnnp_op_t& o = ops[i];
      if (o.kind != BNNP_OP_LINEAR) continue;
      int act = 0;
      if (i + 1 < n_ops && is_elementwise_op(ops[i + 1])) act = ops[i + 1].kind;
      float* dst = buf[pp];
      const int64_t out_ld = o.out_c, out_bs = N * (int64_t)o.out_c;
      int rc;
      if ((long long)N * o.out_c * o.in_c >= (1LL << 22) && (long long)S * N * o.out_c * o.in_c >= (1LL << 25))
        rc = tcg::gemm_tc(N, o.out_c, o.in_c, (int)S, tcg::TcRowMajor{in, in_ld, in_bs},
                          tcg::TcRowMajor{theta_s + o.theta_off_w, o.in_c, P},
                          TcBnnStore(dst, out_ld, out_bs, theta_s, P, o.theta_off_b, o.has_bias, act), 1, st);
      else
        rc = gemm_simt(N, o.out_c, o.in_c, (int)S, LoadRowMajorA{in, in_ld, in_bs},
                       LoadColMajorB{theta_s + o.theta_off_w, o.in_c, P},
                       BnnStore{dst, out_ld, out_bs, theta_s, P, o.theta_off_b, o.has_bias, act}, st);
      if (rc) return rc;
      in = dst; in_ld = out_ld; in_bs = out_bs; pp ^= 1;
      K = o.out_c;
    }
    return bnnp_lik_inv_link(lik, in, S * N, K, out, stream);
  }
  bnnp_plan_t plan;
  int rc = bnnp_net_plan(ops, n_ops, N, 1, &plan);
  if (rc) return rc;
  if (plan.P != P) return BNNP_ERR_INVALID;
  float* logits = scratch + plan.total_floats;
  bnnp_op_t local[BNNP_MAX_OPS];
  for (int64_t s = 0; s < S; ++s) {
    for (int i = 0; i < n_ops; ++i) {
      local[i] = ops[i];
      if (is_param_op(ops[i])) {
        local[i].weight = theta_s + s * P + ops[i].theta_off_w;
        local[i].bias = ops[i].has_bias ? theta_s + s * P + ops[i].theta_off_b : nullptr;
      }
    }
    rc = bnnp_net_forward(local, n_ops, &plan, X, N, scratch, logits, stream);
    if (rc) return rc;
    rc = bnnp_lik_inv_link(lik, logits, N, plan.K, out + s * N * plan.K, stream);
    if (rc) return rc;
  }
  return BNNP_OK;
}

// ------------------------------------------------------------------------------------------
// dense helpers
// ------------------------------------------------------------------------------------------
extern "C" int bnnp_sgemm(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                          int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
                          void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  StoreAxpby ep{C, ldc, 0, alpha, beta};
  if (!ta && !tb) return gemm_simt(M, N, K, 1, LoadRowMajorA{A, lda, 0}, LoadRowMajorB{B, ldb, 0}, ep, st);
  if (!ta && tb) return gemm_simt(M, N, K, 1, LoadRowMajorA{A, lda, 0}, LoadColMajorB{B, ldb, 0}, ep, st);
  if (ta && !tb) return gemm_simt(M, N, K, 1, LoadColMajorA{A, lda, 0}, LoadRowMajorB{B, ldb, 0}, ep, st);
  return gemm_simt(M, N, K, 1, LoadColMajorA{A, lda, 0}, LoadColMajorB{B, ldb, 0}, ep, st);
}

extern "C" int bnnp_axpy(int64_t n, float alpha, const float* x, float* y, void* stream) {
  if (!x || !y || n <= 0) return BNNP_ERR_INVALID;
  axpy_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(n, alpha, x, y);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" const char* bnnp_strerror(int code) {
  switch (code) {
    case BNNP_OK: return "ok";
    case BNNP_ERR_INVALID: return "invalid argument";
    case BNNP_ERR_UNSUPPORTED: return "unsupported layer type or size";
    case BNNP_ERR_CUDA: return "CUDA error";
    case BNNP_ERR_NO_DEVICE: return "no sm_100 device";
    default: return "unknown error";
  }
}
unsigned long long g_bnnp_launches = 0;
BnnpOptions g_bnnp_opt = {0, 0, 0, 0, 0, 0, 0};
static int* option_slot(const char* name) {
  if (!name) return nullptr;
  if (!strcmp(name, "gram_kernel")) return &g_bnnp_opt.gram_kernel;
  if (!strcmp(name, "glm_single")) return &g_bnnp_opt.glm_single;
  if (!strcmp(name, "glm_simt")) return &g_bnnp_opt.glm_simt;
  if (!strcmp(name, "debug_sync")) return &g_bnnp_opt.debug_sync;
  if (!strcmp(name, "net_simt")) return &g_bnnp_opt.net_simt;
  if (!strcmp(name, "gemm_generated")) return &g_bnnp_opt.gemm_generated;
  if (!strcmp(name, "conv_bwd_explicit")) return &g_bnnp_opt.conv_bwd_explicit;
  return nullptr;
}
extern "C" int bnnp_set_option(const char* name, int value) {
  int* slot = option_slot(name);
  if (!slot) return BNNP_ERR_INVALID;
  *slot = value;
  return BNNP_OK;
}
extern "C" int bnnp_get_option(const char* name) {
  int* slot = option_slot(name);
  return slot ? *slot : BNNP_ERR_INVALID;
}
extern "C" uint64_t bnnp_launch_count(void) { return g_bnnp_launches; }
extern "C" int bnnp_version(void) { return 100; }
extern "C" int bnnp_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return 0; }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { cudaGetLastError(); return 0; }
  return p.major == 10 ? 1 : 0;
}
