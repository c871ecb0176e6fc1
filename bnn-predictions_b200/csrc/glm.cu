// GLM predictive: f_mu = f + J (mu - theta*), f_var = J Sigma J^T, then the fused
// Cholesky + sampling + inverse-link epilogue.  preds/predictives.py:50-76, preds/kron.py:112-143.
#include <stdlib.h>
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_tma.cuh"

namespace bnnp {

struct Im2colG {   // same implicit im2col as ggn.cu (kept local to this TU)
  const float* in; int64_t ild; int V;
  int Cin, H, W, OW, kh, kw, sh, sw, pl, pt;
  __device__ float at(int64_t n, int pos, int e) const {
    int b = e % kw, a = (e / kw) % kh, ci = e / (kw * kh);
    int oh = pos / OW, ow = pos % OW;
    int ih = oh * sh - pt + a, iw = ow * sw - pl + b;
    if (ih < 0 || ih >= H || iw < 0 || iw >= W) return 0.f;
    return in[n * ild + ((int64_t)ci * H + ih) * W + iw];
  }
};
static Im2colG make_im2col_g(const bnnp_op_t& o, const float* in, int64_t ild, int V) {
  return Im2colG{in, ild, V, o.in_c, o.in_h, o.in_w, o.out_w, o.kh, o.kw, o.stride_h, o.stride_w,
                 o.pad_l, o.pad_t};
}
static const float* op_input_g(const bnnp_op_t* ops, const bnnp_plan_t* plan, int i, const float* X,
                               const float* ws, int64_t* ld) {
  if (i == 0) { *ld = op_in_size(ops[0]); return X; }
  *ld = plan->op[i - 1].act_ld;
  return ws + plan->op[i - 1].act_off;
}

__global__ void zero_kernel(float* p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}
static int zero(float* p, int64_t n, cudaStream_t st) {
  zero_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(p, n);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

// ---- per-row Jacobian block of a conv weight: Jb[r, co*E + e] -----------------------------
struct ConvGradA2 {
  static constexpr bool K_CONTIG = true;
  const float* g; int64_t gld; int L;
  __device__ float operator()(int64_t co, int64_t pos, int r) const { return g[(int64_t)r * gld + co * L + pos]; }
};
struct ConvUnfoldB2 {
  static constexpr bool K_CONTIG = false;
  Im2colG u;
  __device__ float operator()(int64_t pos, int64_t e, int r) const { return u.at(r / u.V, (int)pos, (int)e); }
};
struct StoreBlock2 {
  float* dst; int64_t ld; int E;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t co, int64_t e, float acc, int r) const { dst[(int64_t)r * ld + co * E + e] = acc; }
};
__global__ void conv_possum2_kernel(const float* __restrict__ g, int64_t gld, int64_t R, int Cout, int L,
                                    float* __restrict__ s) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * Cout) return;
  const float* p = g + (i / Cout) * gld + (i % Cout) * L;
  float acc = 0.f;
  for (int k = 0; k < L; ++k) acc += p[k];
  s[i] = acc;
}

// ---- f_mu ---------------------------------------------------------------------------------
// f_mu[n,k] += sum_u G[(n,k),u] * (sum_j a[n,j] dW[u,j] + db[u])
__global__ void fmu_direct_kernel(const float* __restrict__ G, int64_t gld, const float* __restrict__ A,
                                  int64_t ald, const float* __restrict__ d, int64_t offw, int64_t offb,
                                  int m, int nin, int has_bias, int K, int64_t N, float* __restrict__ fmu) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  int64_t n = i / K;
  float acc = 0.f;
  for (int u = 0; u < m; ++u) {
    float t = has_bias ? d[offb + u] : 0.f;
    for (int j = 0; j < nin; ++j) t = fmaf(A[n * ald + j], d[offw + (int64_t)u * nin + j], t);
    acc = fmaf(G[i * gld + u], t, acc);
  }
  fmu[i] += acc;
}

// ---- f_var, full covariance (Linear-only) ---------------------------------------------------
struct JacRowA {        // A((n,k), p) = G[(n,k), u(p)] * a[n, c(p)]
  static constexpr bool K_CONTIG = true;
  const float* G; const float* A; const int32_t* pu; const int32_t* pc; int64_t gld, ald; int K;
  __device__ float operator()(int64_t r, int64_t p, int) const {
    return G[r * gld + pu[p]] * A[(r / K) * ald + pc[p]];
  }
};
// f_var[n,k,l] = sum_q T[(n,k), q] * J[(n,l), q]      one block per (n,k); K accumulators in registers
template <int KK>
__global__ void __launch_bounds__(256)
fvar_from_T_kernel(const float* T, JacRowA J, int64_t P, float* __restrict__ fvar) {
  __shared__ float red[KK][8];
  const int64_t r = blockIdx.x;              // (n,k)
  const int64_t n = r / KK;
  const int k = (int)(r % KK);
  float acc[KK];
#pragma unroll
  for (int l = 0; l < KK; ++l) acc[l] = 0.f;
  const float* Trow = T + r * P;
  const float* Arow = J.A + n * J.ald;
  const float* Grow = J.G + n * KK * J.gld;
  for (int64_t q = threadIdx.x; q < P; q += 256) {
    const float ta = Trow[q] * Arow[J.pc[q]];
    const int u = J.pu[q];
#pragma unroll
    for (int l = 0; l < KK; ++l) acc[l] = fmaf(ta, Grow[l * J.gld + u], acc[l]);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int l = 0; l < KK; ++l) {
    const float v = warp_sum(acc[l]);
    if (lane == 0) red[l][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < KK) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[threadIdx.x][i];
    fvar[(n * KK + k) * KK + threadIdx.x] = v;
  }
}
template <int KK>
static int launch_fvar_from_T(const float* T, const JacRowA& J, int64_t P, int64_t N, float* fvar, cudaStream_t st) {
  fvar_from_T_kernel<KK><<<(unsigned)(N * KK), 256, 0, st>>>(T, J, P, fvar);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}
__global__ void param_maps2_kernel(int32_t* __restrict__ pu, int32_t* __restrict__ pc, int64_t off,
                                   int m, int nin, int unit0, int col0, int is_bias) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t cnt = is_bias ? m : (int64_t)m * nin;
  if (e >= cnt) return;
  if (is_bias) { pu[off + e] = unit0 + (int)e; pc[off + e] = col0 + nin; }
  else { pu[off + e] = unit0 + (int)(e / nin); pc[off + e] = col0 + (int)(e % nin); }
}


// ---- f_var, full covariance on tensor cores (structured) --------------------------------------
//   T[n,u,u'] = sum_{p in unit u} sum_{q in unit u'} a~[n,j(p)] Sigma[p,q] a~[n,j(q)]
//   f_var[n,k,l] = sum_{u,u'} G[n,k,u] T[n,u,u'] G[n,l,u']
// One gemm_tc launch per layer l' (batch = its units u'):  Y_u'[n,p] = sum_q a~[n,j(q)] Sigma[p,q]
// with the row-dot  T[n,u(p),u'] += a~[n,j(p)] * Y_u'[n,p]  fused into the epilogue.
struct SigmaUnitB {       // B(k, p) for batch u':  Sigma[p, theta(u',k)],  k < nin: weight (u',k), k == nin: bias u'
  static constexpr bool K_CONTIG = true;
  struct Row { const float* w; const float* bias; };
  const float* Sigma; long long P, offw, offb; int nin, has_bias;
  __device__ __forceinline__ Row row(int64_t p, int up) const {
    const float* r = Sigma + p * P;
    return Row{r + offw + (long long)up * nin, r + offb + up};
  }
  __device__ __forceinline__ void load8(const Row& rw, int64_t k0, int64_t, float* out) const {
    const float* w = rw.w + k0;
    if (k0 + 8 <= nin) {
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = __ldg(w + e);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t k = k0 + e;
        out[e] = (k < nin) ? __ldg(w + e) : ((k == nin && has_bias) ? __ldg(rw.bias) : 0.f);
      }
    }
  }
};
struct RowDotT {          // epilogue: T[n, u(p), u'] += a~[n, c(p)] * acc(n, p), flushed once per run of equal u
  struct State { float acc; int u; };
  float* T; const float* A; const int32_t* pu; const int32_t* pc; long long ald; int Mtot, unit0p;
  __device__ __forceinline__ void begin(State& s) const { s.acc = 0.f; s.u = -1; }
  __device__ __forceinline__ void flush(State& s, int64_t n, int up) const {
    if (s.u >= 0) atomicAdd(T + ((long long)n * Mtot + s.u) * Mtot + unit0p + up, s.acc);
  }
  __device__ __forceinline__ void row(State& s, int64_t n, int64_t p0, const float* v, int nvalid, int up, int) const {
    const float* __restrict__ a = A + n * ald;
    float prod[32];
    int uu[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {          // all loads first (no atomic in between)
      const bool ok = e < nvalid;
      const int64_t p = p0 + (ok ? e : 0);
      uu[e] = ok ? __ldg(pu + p) : -2;
      prod[e] = __ldg(a + __ldg(pc + p)) * v[e];
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      if (uu[e] != -2) {
        if (uu[e] != s.u) { flush(s, n, up); s.u = uu[e]; s.acc = 0.f; }
        s.acc += prod[e];
      }
    }
  }
  __device__ __forceinline__ void end(State& s, int64_t n, int up, int) const { flush(s, n, up); s.u = -1; }
};
// f_var[n,k,l] = sum_{u,u'} G[(n,k),u] T[n,u,u'] G[(n,l),u']     one block per input
__global__ void fvar_from_units_kernel(const float* __restrict__ G, int64_t gld, const float* __restrict__ T,
                                       int K, int M, float* __restrict__ fvar) {
  extern __shared__ float shm[];          // R[K][M]
  const int64_t n = blockIdx.x;
  const float* Tn = T + n * M * M;
  for (int i = threadIdx.x; i < K * M; i += blockDim.x) {
    const int k = i / M, up = i % M;
    float acc = 0.f;
    for (int u = 0; u < M; ++u) acc = fmaf(G[(n * K + k) * gld + u], Tn[u * M + up], acc);
    shm[i] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) {
    const int k = i / K, l = i % K;
    float acc = 0.f;
    for (int up = 0; up < M; ++up) acc = fmaf(shm[k * M + up], G[(n * K + l) * gld + up], acc);
    fvar[n * K * K + i] = acc;
  }
}

static int glm_fvar_full_tc(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, const float* ws,
                            const float* Sigma, float* f_var, float* scratch, cudaStream_t st) {
  const int K = plan->K, M = plan->Mtot;
  const int64_t P = plan->P, Ppad = (P + 31) / 32 * 32;
  int32_t* pu = reinterpret_cast<int32_t*>(scratch);
  int32_t* pc = pu + Ppad;
  float* T = scratch + 2 * Ppad;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    param_maps2_kernel<<<(unsigned)ceil_div64((int64_t)o.out_c * o.in_c, 256), 256, 0, st>>>(
        pu, pc, o.theta_off_w, o.out_c, o.in_c, q.lin_unit, q.lin_col, 0);
    BNNP_LAUNCH_CHECK();
    if (o.has_bias) {
      param_maps2_kernel<<<(unsigned)ceil_div64(o.out_c, 256), 256, 0, st>>>(pu, pc, o.theta_off_b, o.out_c, o.in_c,
                                                                             q.lin_unit, q.lin_col, 1);
      BNNP_LAUNCH_CHECK();
    }
  }
  int rc = zero(T, N * M * M, st);
  if (rc) return rc;
  const float* A = ws + plan->acat_off;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    const int kdim = o.in_c + (o.has_bias ? 1 : 0);
    rc = tcg::gemm_tc(N, P, kdim, o.out_c, tcg::TcRowMajor{A + q.lin_col, plan->Dtot, 0},
                      SigmaUnitB{Sigma, P, o.theta_off_w, o.theta_off_b, o.in_c, o.has_bias},
                      RowDotT{T, A, pu, pc, plan->Dtot, M, q.lin_unit}, 1, st);
    if (rc) return rc;
  }
  fvar_from_units_kernel<<<(unsigned)N, 256, (size_t)K * M * sizeof(float), st>>>(ws + plan->gcat_off, M, T, K, M, f_var);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

// ---- f_var, diagonal ------------------------------------------------------------------------
struct LoadSqA {        // A(n, j) = a[n, col0 + j]^2   (ones column included: j == nin)
  static constexpr bool K_CONTIG = true;
  const float* A; int64_t ld; int col0;
  __device__ float operator()(int64_t n, int64_t j, int) const { float a = A[n * ld + col0 + j]; return a * a; }
};
struct LoadSigma2B {    // B(j, i) = sigma2[offw + i*nin + j], j == nin -> sigma2[offb + i]
  static constexpr bool K_CONTIG = true;
  const float* s; int64_t offw, offb; int nin, has_bias;
  __device__ float operator()(int64_t j, int64_t i, int) const {
    if (j < nin) return s[offw + i * nin + j];
    return has_bias ? s[offb + i] : 0.f;
  }
};
// fvar[n,k,l] += sum_u G[(n,k),u] G[(n,l),u] T[n,u]
__global__ void fvar_weighted_gg_kernel(const float* __restrict__ G, int64_t gld, const float* __restrict__ T,
                                        int64_t tld, int64_t N, int K, int m, float* __restrict__ fvar) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K * K) return;
  int l = (int)(i % K), k = (int)((i / K) % K); int64_t n = i / (K * K);
  const float* gk = G + (n * K + k) * gld;
  const float* gl = G + (n * K + l) * gld;
  const float* t = T + n * tld;
  float acc = 0.f;
  for (int u = 0; u < m; ++u) acc = fmaf(gk[u] * gl[u], t[u], acc);
  fvar[i] += acc;
}
// fvar[n,k,l] += sum_e w[e] * B[(n,k), e] * B[(n,l), e]        (w may be NULL -> 1)
__global__ void fvar_weighted_block_kernel(const float* __restrict__ B, int64_t bld, const float* __restrict__ w,
                                           int64_t E, int K, float* __restrict__ fvar) {
  __shared__ float red[8];
  int64_t i = blockIdx.x;            // (n,k,l)
  int l = (int)(i % K), k = (int)((i / K) % K); int64_t n = i / (K * K);
  const float* bk = B + (n * K + k) * bld;
  const float* bl = B + (n * K + l) * bld;
  float acc = 0.f;
  for (int64_t e = threadIdx.x; e < E; e += blockDim.x) acc = fmaf(bk[e] * bl[e], w ? w[e] : 1.f, acc);
  acc = warp_sum(acc);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int j = 0; j < (int)(blockDim.x / 32); ++j) v += red[j];
    fvar[i] += v;
  }
}

// ---- f_var, Kronecker ---------------------------------------------------------------------
// Dinv[i, j] = 1/(l1[i] l2[j] + delta)   or dampened 1/((l1[i]+sqrt d)(l2[j]+sqrt d))   kron.py:130-134
__global__ void kron_dinv_kernel(const float* __restrict__ l1, const float* __restrict__ l2, int m, int n,
                                 float delta, int dampen, int take_sqrt, float* __restrict__ D) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)m * n) return;
  float a = l1[i / n], b = l2[i % n], v;
  if (dampen) { float sd = sqrtf(delta); v = 1.f / ((a + sd) * (b + sd)); }
  else v = 1.f / (a * b + delta);
  D[i] = take_sqrt ? sqrtf(v) : v;
}
__global__ void inv_shift_kernel(const float* __restrict__ l, int m, float delta, int take_sqrt,
                                 float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  float v = 1.f / (l[i] + delta);
  out[i] = take_sqrt ? sqrtf(v) : v;
}
struct LoadSqRows {     // A(n, j) = Ahat[n, j]^2
  static constexpr bool K_CONTIG = true;
  const float* A; int64_t ld;
  __device__ float operator()(int64_t n, int64_t j, int) const { float a = A[n * ld + j]; return a * a; }
};
struct LoadDinvT {      // B(j, i) = Dinv[i, j]
  static constexpr bool K_CONTIG = true;
  const float* D; int n;
  __device__ float operator()(int64_t j, int64_t i, int) const { return D[i * n + j]; }
};
// fvar[n,k,l] += sum_u w[u] Gh[(n,k),u] Gh[(n,l),u]
__global__ void fvar_gg_const_kernel(const float* __restrict__ Gh, int64_t gld, const float* __restrict__ w,
                                     int64_t N, int K, int m, float* __restrict__ fvar) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K * K) return;
  int l = (int)(i % K), k = (int)((i / K) % K); int64_t n = i / (K * K);
  const float* gk = Gh + (n * K + k) * gld;
  const float* gl = Gh + (n * K + l) * gld;
  float acc = 0.f;
  for (int u = 0; u < m; ++u) acc = fmaf(gk[u] * gl[u], w[u], acc);
  fvar[i] += acc;
}

// fvar[n,k,l] += sum_u Gh[(n,k),u] Gh[(n,l),u] (T[n,u] + wb[u])   -- weight and bias terms of a Linear layer in ONE pass.
// One warp per input: the K rows of Gh[n] are read once (coalesced along u), the K(K+1)/2 products live in registers
// (the thread-per-(n,k,l) kernels above re-read every row K times with stride-m access).
template <int KK>
__global__ void __launch_bounds__(128)
fvar_gg_fused_kernel(const float* __restrict__ Gh, int64_t gld, const float* __restrict__ T, int64_t tld,
                     const float* __restrict__ wb, int64_t N, int m, float* __restrict__ fvar) {
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float acc[KK * (KK + 1) / 2];
#pragma unroll
  for (int i = 0; i < KK * (KK + 1) / 2; ++i) acc[i] = 0.f;
  const float* g = Gh + n * KK * gld;
  const float* t = T + n * tld;
  for (int u = lane; u < m; u += 32) {
    const float w = t[u] + (wb ? wb[u] : 0.f);
    float gv[KK];
#pragma unroll
    for (int k = 0; k < KK; ++k) gv[k] = g[k * gld + u];
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      const float gw = gv[k] * w;
#pragma unroll
      for (int l = 0; l <= k; ++l) acc[k * (k + 1) / 2 + l] = fmaf(gw, gv[l], acc[k * (k + 1) / 2 + l]);
    }
  }
#pragma unroll
  for (int k = 0; k < KK; ++k)
#pragma unroll
    for (int l = 0; l <= k; ++l) {
      const float v = warp_sum(acc[k * (k + 1) / 2 + l]);
      if (lane == 0) {
        fvar[(n * KK + k) * KK + l] += v;
        if (l != k) fvar[(n * KK + l) * KK + k] += v;
      }
    }
}
template <int KK>
static int launch_gg_fused(const float* Gh, int64_t gld, const float* T, int64_t tld, const float* wb, int64_t N, int m,
                           float* fvar, cudaStream_t st) {
  fvar_gg_fused_kernel<KK><<<(unsigned)ceil_div64(N * 32, 128), 128, 0, st>>>(Gh, gld, T, tld, wb, N, m, fvar);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}
struct TcRowMajorSq {      // X(r, k) = p[r*ld + k]^2
  static constexpr bool K_CONTIG = true;
  using Row = const float*;
  const float* p; long long ld;
  __device__ __forceinline__ Row row(int64_t r, int) const { return p + r * ld; }
  __device__ __forceinline__ void load8(const Row& rw, int64_t k0, int64_t K, float* out) const {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float v = (k0 + e < K) ? __ldg(rw + k0 + e) : 0.f;
      out[e] = v * v;
    }
  }
};
constexpr long long kTcMinMacsK = 1LL << 25;
// option glm_simt (tests): keep the kron / diag GLM contractions on the fp32 SIMT kernels and the unfused reductions
static bool glm_force_simt() { return g_bnnp_opt.glm_simt != 0; }

// ---- sampling epilogue ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
  *hi = __umulhi(a, b);
  return a * b;
}
// Philox4x32-10
__device__ __forceinline__ uint4 philox4(uint4 ctr, uint2 key) {
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

// One thread owns one input n (its K x K Cholesky factor lives in registers) and walks a slab of
// samples; consecutive threads own consecutive n so the [S,N,K] stores of a warp are contiguous.
// MEAN: instead of storing the [S,N,K] samples, each thread sums its slab of samples in registers and adds
// slab_sum / S to out[n, :] ([N,K], zeroed by the caller): the `.mean(0)` every driver applies
// (experiments/imginference.py:391, classification.py:36) without the [S,N,K] round trip through HBM.
template <int K, bool MEAN>
__global__ void __launch_bounds__(128)
glm_sample_kernel(int lik, const float* __restrict__ fmu, const float* __restrict__ fvar,
                  const float* __restrict__ eps, uint64_t seed, int64_t S, int64_t N,
                  float* __restrict__ out, int32_t* __restrict__ status, int s_per_block) {
  const int64_t n_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = n_raw < N;
  const int64_t n = valid ? n_raw : N - 1;         // out-of-range lanes shadow the last input and never store
  // samples of a warp's 32 inputs are staged in shared memory and stored as one contiguous run per sample: a thread
  // owns K consecutive floats, so direct stores touch every 32-byte sector K times (measured 0.15 of HBM bandwidth)
  __shared__ float stg[4][32 * (K + 1)];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nw0 = (int64_t)blockIdx.x * blockDim.x + wid * 32;
  const int cnt = (int)(N - nw0 < 32 ? (N - nw0 < 0 ? 0 : N - nw0) : 32) * K;     // floats of this warp per sample
  float L[K * (K + 1) / 2];
  float mu[K];
  bool ok = true;
  // lower Cholesky from the lower triangle of f_var[n]
#pragma unroll
  for (int i = 0; i < K; ++i) {
    mu[i] = fmu[n * K + i];
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      float s = fvar[(n * K + i) * K + j];
#pragma unroll
      for (int t = 0; t < j; ++t) s -= L[i * (i + 1) / 2 + t] * L[j * (j + 1) / 2 + t];
      if (i == j) {
        if (!(s > 0.f)) ok = false;
        L[i * (i + 1) / 2 + j] = sqrtf(s);
      } else {
        L[i * (i + 1) / 2 + j] = s / L[j * (j + 1) / 2 + j];
      }
    }
  }
  if (blockIdx.y == 0 && valid) status[n] = ok ? 0 : 1;
  int64_t s0 = (int64_t)blockIdx.y * s_per_block;
  int64_t s1 = s0 + s_per_block < S ? s0 + s_per_block : S;
  float acc[K];
#pragma unroll
  for (int i = 0; i < K; ++i) acc[i] = 0.f;
  for (int64_t s = s0; s < s1; ++s) {
    float e[K], f[K];
    if (eps) {
#pragma unroll
      for (int i = 0; i < K; ++i) e[i] = eps[(s * N + n) * K + i];
    } else {
#pragma unroll
      for (int i = 0; i < K; i += 4) {
        uint64_t c = (uint64_t)(s * N + n);
        uint4 r = philox4(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)(i / 4), 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
        e[i] = a.x;
        if (i + 1 < K) e[i + 1] = a.y;
        if (i + 2 < K) e[i + 2] = b.x;
        if (i + 3 < K) e[i + 3] = b.y;
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      float v = mu[i];
#pragma unroll
      for (int j = 0; j <= i; ++j) v = fmaf(L[i * (i + 1) / 2 + j], e[j], v);
      f[i] = v;
      mx = fmaxf(mx, v);
    }
    if (lik == BNNP_LIK_CATEGORICAL) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i) { f[i] = __expf(f[i] - mx); sum += f[i]; }
      float inv = 1.f / sum;
#pragma unroll
      for (int i = 0; i < K; ++i) f[i] *= inv;
    } else if (lik == BNNP_LIK_BERNOULLI) {
#pragma unroll
      for (int i = 0; i < K; ++i) f[i] = 1.f / (1.f + __expf(-f[i]));
    }
    if (MEAN) {
#pragma unroll
      for (int i = 0; i < K; ++i) acc[i] += f[i];
    } else {
      float* st_ = stg[wid];
#pragma unroll
      for (int i = 0; i < K; ++i) st_[lane * (K + 1) + i] = f[i];
      __syncwarp();
      float* o = out + (s * N + nw0) * K;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int e = lane + 32 * j;
        if (e < cnt) o[e] = st_[(e / K) * (K + 1) + e % K];
      }
      __syncwarp();
    }
  }
  if (MEAN) {
    const float invS = 1.f / (float)S;
#pragma unroll
    for (int i = 0; i < K; ++i)
      if (valid) atomicAdd(out + n * K + i, acc[i] * invS);
  }
}


// Reference fallback when the batched Cholesky fails (preds/predictives.py:73-76):
// Normal(f_mu, diag(f_var).clamp(1e-5)) -- the variance is used as the *scale* (reference quirk).
__global__ void glm_sample_diag_kernel(int lik, const float* __restrict__ fmu, const float* __restrict__ fvar,
                                       const float* __restrict__ eps, uint64_t seed, int64_t S, int64_t N,
                                       int K, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // (s, n)
  if (i >= S * N) return;
  int64_t n = i % N;
  float f[BNNP_MAX_K];
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) {
    float e;
    if (eps) e = eps[i * K + k];
    else {
      uint4 r = philox4(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)k, 1u),
                        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      e = box_muller(r.x, r.y).x;
    }
    float sc = fmaxf(fvar[(n * K + k) * K + k], 1e-5f);
    f[k] = fmaf(sc, e, fmu[n * K + k]);
    mx = fmaxf(mx, f[k]);
  }
  float* o = out + i * K;
  if (lik == BNNP_LIK_CATEGORICAL) {
    float sum = 0.f;
    for (int k = 0; k < K; ++k) { f[k] = __expf(f[k] - mx); sum += f[k]; }
    for (int k = 0; k < K; ++k) o[k] = f[k] / sum;
  } else if (lik == BNNP_LIK_BERNOULLI) {
    for (int k = 0; k < K; ++k) o[k] = 1.f / (1.f + __expf(-f[k]));
  } else {
    for (int k = 0; k < K; ++k) o[k] = f[k];
  }
}

template <int K>
static int launch_sample(int lik, const float* fmu, const float* fvar, const float* eps, uint64_t seed,
                         int64_t S, int64_t N, float* out, int32_t* status, bool mean, cudaStream_t st) {
  unsigned gx = (unsigned)ceil_div64(N, 128);
  // enough y-slabs to fill the machine (148 SMs x several CTAs), at least 8 samples per slab
  int64_t want = (148 * 8 + gx - 1) / gx;
  int64_t slabs = want < 1 ? 1 : want;
  if (slabs > S) slabs = S;
  int spb = (int)ceil_div64(S, slabs);
  slabs = ceil_div64(S, spb);
  if (slabs > 65535) return BNNP_ERR_UNSUPPORTED;
  if (mean) {
    int rc = zero(out, N * K, st);
    if (rc) return rc;
    glm_sample_kernel<K, true><<<dim3(gx, (unsigned)slabs), 128, 0, st>>>(lik, fmu, fvar, eps, seed, S, N, out, status, spb);
  } else {
    glm_sample_kernel<K, false><<<dim3(gx, (unsigned)slabs), 128, 0, st>>>(lik, fmu, fvar, eps, seed, S, N, out, status, spb);
  }
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

}  // namespace bnnp
using namespace bnnp;

static bool linear_only_g(const bnnp_op_t* ops, int n_ops) {
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind == BNNP_OP_CONV2D || ops[i].kind == BNNP_OP_MAXPOOL2D || ops[i].kind == BNNP_OP_AVGPOOL2D)
      return false;
  return true;
}

// ------------------------------------------------------------------------------------ f_mu
extern "C" int bnnp_glm_fmu(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                            const float* X, const float* ws, const float* f, const float* delta,
                            float* f_mu, void* stream) {
  if (!ops || !plan || !ws || !f || !f_mu || N <= 0) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  (void)X;
  BNNP_CUDA_CHECK(cudaMemcpyAsync(f_mu, f, sizeof(float) * N * plan->K, cudaMemcpyDeviceToDevice, st));
  if (!delta) return BNNP_OK;
  if (!linear_only_g(ops, n_ops)) return BNNP_ERR_UNSUPPORTED;
  const int K = plan->K;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    fmu_direct_kernel<<<(unsigned)ceil_div64(N * K, 128), 128, 0, st>>>(
        ws + q.grad_off, q.grad_ld, ws + plan->acat_off + q.lin_col, plan->Dtot, delta, o.theta_off_w,
        o.theta_off_b, o.out_c, o.in_c, o.has_bias, K, N, f_mu);
    BNNP_LAUNCH_CHECK();
  }
  return BNNP_OK;
}

// ------------------------------------------------------------------------------------ full
static bool fvar_full_use_tc(const bnnp_plan_t* plan, int engine) {
  if (engine == 1) return false;
  if (engine == 2) return true;
  return plan->P >= 4096;                 // tiny posteriors: the SIMT path is launch-bound anyway
}
extern "C" int64_t bnnp_glm_fvar_full_scratch_floats(const bnnp_plan_t* plan, int64_t N, int engine) {
  int64_t Ppad = (plan->P + 31) / 32 * 32;
  if (fvar_full_use_tc(plan, engine)) return 2 * Ppad + N * (int64_t)plan->Mtot * plan->Mtot + 64;
  return 2 * Ppad + N * plan->K * plan->P + 64;
}

extern "C" int bnnp_glm_fvar_full(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                                  const float* ws, const float* Sigma, float* f_var, float* scratch,
                                  int engine, void* stream) {
  if (!ops || !plan || !ws || !Sigma || !f_var || !scratch || N <= 0) return BNNP_ERR_INVALID;
  if (!linear_only_g(ops, n_ops)) return BNNP_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  if (fvar_full_use_tc(plan, engine)) return glm_fvar_full_tc(ops, n_ops, plan, N, ws, Sigma, f_var, scratch, st);
  const int K = plan->K;
  const int64_t P = plan->P, Ppad = (P + 31) / 32 * 32;
  int32_t* pu = reinterpret_cast<int32_t*>(scratch);
  int32_t* pc = pu + Ppad;
  float* T = scratch + 2 * Ppad;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    param_maps2_kernel<<<(unsigned)ceil_div64((int64_t)o.out_c * o.in_c, 256), 256, 0, st>>>(
        pu, pc, o.theta_off_w, o.out_c, o.in_c, q.lin_unit, q.lin_col, 0);
    BNNP_LAUNCH_CHECK();
    if (o.has_bias) {
      param_maps2_kernel<<<(unsigned)ceil_div64(o.out_c, 256), 256, 0, st>>>(pu, pc, o.theta_off_b, o.out_c, o.in_c,
                                                                             q.lin_unit, q.lin_col, 1);
      BNNP_LAUNCH_CHECK();
    }
  }
  JacRowA J{ws + plan->gcat_off, ws + plan->acat_off, pu, pc, plan->Mtot, plan->Dtot, K};
  int rc = gemm_simt(N * K, P, P, 1, J, LoadRowMajorB{Sigma, P, 0}, StoreAxpby{T, P, 0, 1.f, 0.f}, st);
  if (rc) return rc;
#define BNNP_CASE(KK) case KK: return launch_fvar_from_T<KK>(T, J, P, N, f_var, st);
  switch (K) {
    BNNP_CASE(1) BNNP_CASE(2) BNNP_CASE(3) BNNP_CASE(4) BNNP_CASE(5) BNNP_CASE(6) BNNP_CASE(7) BNNP_CASE(8)
    BNNP_CASE(9) BNNP_CASE(10) BNNP_CASE(11) BNNP_CASE(12) BNNP_CASE(13) BNNP_CASE(14) BNNP_CASE(15) BNNP_CASE(16)
    default: return BNNP_ERR_UNSUPPORTED;
  }
#undef BNNP_CASE
}

// ------------------------------------------------------------------------------------ diag
static int64_t conv_block_floats(const bnnp_op_t* ops, int n_ops, int64_t N, int K) {
  int64_t mx = 0;
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind == BNNP_OP_CONV2D) {
      int64_t E = (int64_t)ops[i].in_c * ops[i].kh * ops[i].kw;
      int64_t v = N * K * (ops[i].out_c * E + ops[i].out_c);
      mx = v > mx ? v : mx;
    }
  return mx;
}

extern "C" int64_t bnnp_glm_fvar_diag_scratch_floats(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                                     int64_t N) {
  int64_t lin = N * (int64_t)plan->Mtot;
  int64_t tma = 0;                    // bf16 hi/lo operand copies of the TMA-fed contraction T = (a^2) sigma2^T
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind == BNNP_OP_LINEAR) {
      int64_t t = linear_tma_scratch_floats(N, ops[i].in_c, ops[i].out_c);
      tma = tma > t ? tma : t;
    }
  lin += tma + 64;
  int64_t cv = conv_block_floats(ops, n_ops, N, plan->K);
  return (lin > cv ? lin : cv) + 64;
}

extern "C" int bnnp_glm_fvar_diag(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                                  const float* X, const float* ws, const float* sigma2, float* f_var,
                                  float* scratch, void* stream) {
  if (!ops || !plan || !ws || !sigma2 || !f_var || !scratch || N <= 0) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  const int K = plan->K;
  int rc = zero(f_var, N * K * K, st);
  if (rc) return rc;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (!is_param_op(o)) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    const float* G = ws + q.grad_off;
    if (o.kind == BNNP_OP_LINEAR) {
      // T[n,u] = sum_j sigma2[u,j] a[n,j]^2 (+ sigma2_b[u])
      // (the bias variance is the constant term wb[u] of the fused reduction when the tensor-core path is taken)
      const bool simt = glm_force_simt();
      const bool tcpath = !simt && (long long)N * o.out_c * o.in_c >= kTcMinMacsK && K <= 16;
      if (tcpath && g_bnnp_opt.gemm_generated == 0 && o.in_c <= 4096)
        rc = linear_tma(ws + plan->acat_off + q.lin_col, plan->Dtot, N, o.in_c, sigma2 + o.theta_off_w, o.in_c, o.out_c, 0,
                        nullptr, scratch, o.out_c, scratch + N * (int64_t)plan->Mtot + 32, st, 1);
      else if (tcpath)
        rc = tcg::gemm_tc(N, o.out_c, o.in_c, 1, TcRowMajorSq{ws + plan->acat_off + q.lin_col, plan->Dtot},
                          tcg::TcRowMajor{sigma2 + o.theta_off_w, o.in_c, 0}, tcg::TcStoreAxpby(scratch, o.out_c, 0, 1.f, 0.f),
                          1, st);
      else
        rc = gemm_simt(N, o.out_c, o.in_c + 1, 1, LoadSqA{ws + plan->acat_off, plan->Dtot, q.lin_col},
                       LoadSigma2B{sigma2, o.theta_off_w, o.theta_off_b, o.in_c, o.has_bias},
                       StoreAxpby{scratch, o.out_c, 0, 1.f, 0.f}, st);
      if (rc) return rc;
      {
        const float* wb = (tcpath && o.has_bias) ? sigma2 + o.theta_off_b : nullptr;
#define BNNP_CASE(KK) case KK: rc = launch_gg_fused<KK>(G, q.grad_ld, scratch, o.out_c, wb, N, o.out_c, f_var, st); break;
        switch (simt ? 0 : K) {
          BNNP_CASE(1) BNNP_CASE(2) BNNP_CASE(3) BNNP_CASE(4) BNNP_CASE(5) BNNP_CASE(6) BNNP_CASE(7) BNNP_CASE(8)
          BNNP_CASE(9) BNNP_CASE(10) BNNP_CASE(11) BNNP_CASE(12) BNNP_CASE(13) BNNP_CASE(14) BNNP_CASE(15) BNNP_CASE(16)
          default:
            fvar_weighted_gg_kernel<<<(unsigned)ceil_div64(N * K * K, 128), 128, 0, st>>>(G, q.grad_ld, scratch, o.out_c,
                                                                                        N, K, o.out_c, f_var);
            BNNP_LAUNCH_CHECK();
            rc = BNNP_OK;
        }
#undef BNNP_CASE
        if (rc) return rc;
      }
    } else {
      int64_t ild; const float* in = op_input_g(ops, plan, i, X, ws, &ild);
      if (!in) return BNNP_ERR_INVALID;
      int L = o.out_h * o.out_w, E = o.in_c * o.kh * o.kw;
      int64_t R = N * K, bw = (int64_t)o.out_c * E;
      if (R > 65535) return BNNP_ERR_UNSUPPORTED;
      float* Jb = scratch;
      float* sb = scratch + R * bw;
      rc = gemm_simt(o.out_c, E, L, (int)R, ConvGradA2{G, q.grad_ld, L}, ConvUnfoldB2{make_im2col_g(o, in, ild, K)},
                     StoreBlock2{Jb, bw, E}, st);
      if (rc) return rc;
      fvar_weighted_block_kernel<<<(unsigned)(N * K * K), 256, 0, st>>>(Jb, bw, sigma2 + o.theta_off_w, bw, K, f_var);
      BNNP_LAUNCH_CHECK();
      if (o.has_bias) {
        conv_possum2_kernel<<<(unsigned)ceil_div64(R * o.out_c, 128), 128, 0, st>>>(G, q.grad_ld, R, o.out_c, L, sb);
        BNNP_LAUNCH_CHECK();
        fvar_weighted_block_kernel<<<(unsigned)(N * K * K), 64, 0, st>>>(sb, o.out_c, sigma2 + o.theta_off_b,
                                                                         o.out_c, K, f_var);
        BNNP_LAUNCH_CHECK();
      }
    }
  }
  return BNNP_OK;
}


// ------------------------------------------------------------------------------------ kron block
// One parameter group of Kron.functional_variance on a materialised Jacobian block
// Jb[R, m*n] (row r = (b,k), row stride ldj):  f_var[b,k,l] += sum_ij M_k M_l Dinv,  M = W1^T J W2
// (equal to bmm(J, (W1 ((W1^T J W2) * Dinv) W2^T)^T) of preds/kron.py:135-141).
// ---- conv layers on the tensor cores: the per-row weight-Jacobian block and its two rotations as gemm_tc launches ----
// U[(n,pos), e] = unfold(in[n])[e, pos]   (row pitch Epad = roundup4(E), zero padded)
__global__ void im2col_g_kernel(Im2colG u, int64_t N, int L, int E, int Epad, float* __restrict__ U) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * L * Epad) return;
  const int e = (int)(i % Epad);
  const int64_t rp = i / Epad;
  U[i] = e < E ? u.at(rp / L, (int)(rp % L), e) : 0.f;
}
struct TcUnfoldTG {        // X(e, k = pos) of batch r: U[((r / V) * L + pos) * Epad + e]
  static constexpr bool K_CONTIG = false;
  using Row = const float*;
  const float* U; long long Epad; int L, V;
  __device__ __forceinline__ Row row(int64_t e, int b) const { return U + (long long)(b / V) * L * Epad + e; }
  __device__ __forceinline__ void load8(const Row& rw, int64_t k0, int64_t K, float* out) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = (k0 + j < K) ? __ldg(rw + (k0 + j) * Epad) : 0.f;
  }
};
// Jb[r, co, e] = sum_pos gout[r, co, pos] * U[n(r), pos, e]     (one [m x L]·[L x E] product per row r = (n,k))
static int conv_jac_blocks_tc(const float* G, int64_t gld, const float* U, int Epad, int64_t R, int V, int m, int E, int L,
                              float* Jb, cudaStream_t st) {
  return tcg::gemm_tc(m, E, L, (int)R, tcg::TcRowMajor{G, L, gld}, TcUnfoldTG{U, Epad, L, V},
                      tcg::TcStoreAxpby(Jb, E, (long long)m * E, 1.f, 0.f), 1, st);
}

static int kron_fvar_block(const float* Jb, int64_t ldj, int64_t R, int K, int m, int n, const float* W1,
                           const float* l1, const float* W2, const float* l2, float delta, int dampen,
                           float* f_var, float* scratch, cudaStream_t st) {
  const int64_t N = R / K;
  int rc;
  if (!W2) {   // 1-factor (bias-like): ghat = J W,  f_var += ghat diag(1/(l+delta)) ghat^T
    float* gh = scratch;                 // [R, m]
    float* db = gh + R * m;              // [m]
    rc = gemm_simt(R, m, m, 1, LoadRowMajorA{Jb, ldj, 0}, LoadRowMajorB{W1, m, 0}, StoreAxpby{gh, m, 0, 1.f, 0.f}, st);
    if (rc) return rc;
    inv_shift_kernel<<<(unsigned)ceil_div64(m, 128), 128, 0, st>>>(l1, m, delta, 0, db);
    BNNP_LAUNCH_CHECK();
    fvar_gg_const_kernel<<<(unsigned)ceil_div64(N * K * K, 128), 128, 0, st>>>(gh, m, db, N, K, m, f_var);
    BNNP_LAUNCH_CHECK();
    return BNNP_OK;
  }
  if (R > 65535) return BNNP_ERR_UNSUPPORTED;
  const int64_t bw = (int64_t)m * n;
  float* T1 = scratch;                   // [R, m, n]
  float* T2 = T1 + R * bw;               // [R, m, n]
  float* D = T2 + R * bw;                // [m, n]
  if (!glm_force_simt() && ldj == bw && (long long)R * m * n * n >= kTcMinMacsK) {
    // T1 = Jb W2 for all rows at once ([R*m, n] x [n, n]); T2[r] = W1^T T1[r] batched over the rows
    rc = tcg::gemm_tc(R * m, n, n, 1, tcg::TcRowMajor{Jb, n, 0}, tcg::TcColMajor{W2, n, 0},
                      tcg::TcStoreAxpby(T1, n, 0, 1.f, 0.f), 1, st);
    if (rc) return rc;
    rc = tcg::gemm_tc(m, n, m, (int)R, tcg::TcColMajor{W1, m, 0}, tcg::TcColMajor{T1, n, bw},
                      tcg::TcStoreAxpby(T2, n, bw, 1.f, 0.f), 1, st);
    if (rc) return rc;
  } else {
    rc = gemm_simt(m, n, n, (int)R, LoadRowMajorA{Jb, n, ldj}, LoadRowMajorB{W2, n, 0},
                   StoreAxpby{T1, n, bw, 1.f, 0.f}, st);
    if (rc) return rc;
    rc = gemm_simt(m, n, m, (int)R, LoadColMajorA{W1, m, 0}, LoadRowMajorB{T1, n, bw},
                   StoreAxpby{T2, n, bw, 1.f, 0.f}, st);
    if (rc) return rc;
  }
  kron_dinv_kernel<<<(unsigned)ceil_div64(bw, 256), 256, 0, st>>>(l1, l2, m, n, delta, dampen, 0, D);
  BNNP_LAUNCH_CHECK();
  fvar_weighted_block_kernel<<<(unsigned)(N * K * K), 256, 0, st>>>(T2, bw, D, bw, K, f_var);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" int bnnp_kron_fvar_block(const float* Jb, int64_t ldj, int K, int m, int n, const float* W1,
                                    const float* l1, const float* W2, const float* l2, float delta,
                                    int dampen, float* f_var, float* scratch, int64_t R, void* stream) {
  if (!Jb || !W1 || !l1 || !f_var || !scratch || R <= 0 || K <= 0 || m <= 0 || R % K) return BNNP_ERR_INVALID;
  if (W2 && (!l2 || n <= 0)) return BNNP_ERR_INVALID;
  return kron_fvar_block(Jb, ldj, R, K, m, n, W1, l1, W2, l2, delta, dampen, f_var, scratch, as_stream(stream));
}

// ------------------------------------------------------------------------------------ kron
extern "C" int64_t bnnp_glm_fvar_kron_scratch_floats(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                                     int64_t N) {
  const int K = plan->K;
  int64_t mx = 0;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (!is_param_op(o)) continue;
    int64_t m = o.out_c, n = (int64_t)o.in_c * o.kh * o.kw, v;
    if (o.kind == BNNP_OP_LINEAR) {
      v = N * n + N * K * m + N * m + m * n + m;
      // bf16 hi/lo operand copies of the TMA-fed rotations
      int64_t t = linear_tma_scratch_floats(N * K, m, m), t2 = linear_tma_scratch_floats(N, n, n > m ? n : m);
      v += (t > t2 ? t : t2) + 64;
    }
    else v = 3 * N * K * m * n + 2 * N * K * m + m * n + m + 128 +
             N * (int64_t)o.out_h * o.out_w * ((n + 3) / 4 * 4);      // + unfolded input of the tensor-core path
    mx = v > mx ? v : mx;
  }
  return mx + 256;
}

extern "C" int bnnp_glm_fvar_kron(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                                  const float* X, const float* ws, const float* const* eig,
                                  const float* delta_w, const float* delta_b, int dampen, float* f_var,
                                  float* scratch, void* stream) {
  if (!ops || !plan || !ws || !eig || !delta_w || !delta_b || !f_var || !scratch || N <= 0) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  const int K = plan->K;
  const int64_t R = N * K;
  int rc = zero(f_var, N * K * K, st);
  if (rc) return rc;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (!is_param_op(o)) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    const float* W1 = eig[4 * i], *l1 = eig[4 * i + 1], *W2 = eig[4 * i + 2], *l2 = eig[4 * i + 3];
    if (!W1 || !l1 || !W2 || !l2) return BNNP_ERR_INVALID;
    const float* G = ws + q.grad_off;
    const int m = o.out_c, n = o.in_c * o.kh * o.kw;
    if (o.kind == BNNP_OP_LINEAR) {
      // J_n,k = g (x) a  =>  W1^T J W2 = (W1^T g)(W2^T a)^T : rotate the factors, never the matrix
      float* Ah = scratch;                 // [N, n]    a W2
      float* Gh = Ah + N * n;              // [N*K, m]  g W1
      float* T = Gh + R * m;               // [N, m]    Dinv (Ah^2)
      float* D = T + N * m;                // [m, n]
      float* db = D + (int64_t)m * n;      // [m]
      float* tsc = db + m + 32;            // operand copies of the TMA-fed products
      // the three contractions go to the tensor cores (split-bf16, fp32-grade) once they are large enough: TMA-fed
      // on bf16 hi/lo copies of the operands (gemm_tma.cu), or with generated operands (option gemm_generated)
      const float* Ain = ws + plan->acat_off + q.lin_col;
      const bool simt = glm_force_simt();
      const bool tma = g_bnnp_opt.gemm_generated == 0 && n <= 4096 && m <= 4096;
      if (!simt && tma && (long long)N * n * n >= kTcMinMacsK)
        rc = linear_tma(Ain, plan->Dtot, N, n, W2, n, n, 1, nullptr, Ah, n, tsc, st);
      else if (!simt && (long long)N * n * n >= kTcMinMacsK)
        rc = tcg::gemm_tc(N, n, n, 1, tcg::TcRowMajor{Ain, plan->Dtot, 0}, tcg::TcColMajor{W2, n, 0},
                          tcg::TcStoreAxpby(Ah, n, 0, 1.f, 0.f), 1, st);
      else
        rc = gemm_simt(N, n, n, 1, LoadRowMajorA{Ain, plan->Dtot, 0}, LoadRowMajorB{W2, n, 0}, StoreAxpby{Ah, n, 0, 1.f, 0.f}, st);
      if (rc) return rc;
      if (!simt && tma && (long long)R * m * m >= kTcMinMacsK)
        rc = linear_tma(G, q.grad_ld, R, m, W1, m, m, 1, nullptr, Gh, m, tsc, st);
      else if (!simt && (long long)R * m * m >= kTcMinMacsK)
        rc = tcg::gemm_tc(R, m, m, 1, tcg::TcRowMajor{G, q.grad_ld, 0}, tcg::TcColMajor{W1, m, 0},
                          tcg::TcStoreAxpby(Gh, m, 0, 1.f, 0.f), 1, st);
      else
        rc = gemm_simt(R, m, m, 1, LoadRowMajorA{G, q.grad_ld, 0}, LoadRowMajorB{W1, m, 0}, StoreAxpby{Gh, m, 0, 1.f, 0.f}, st);
      if (rc) return rc;
      kron_dinv_kernel<<<(unsigned)ceil_div64((int64_t)m * n, 256), 256, 0, st>>>(l1, l2, m, n, delta_w[i], dampen, 0, D);
      BNNP_LAUNCH_CHECK();
      if (!simt && tma && (long long)N * m * n >= kTcMinMacsK)
        rc = linear_tma(Ah, n, N, n, D, n, m, 0, nullptr, T, m, tsc, st, 1);           // T = (Ah^2) D^T
      else if (!simt && (long long)N * m * n >= kTcMinMacsK)
        rc = tcg::gemm_tc(N, m, n, 1, TcRowMajorSq{Ah, n}, tcg::TcRowMajor{D, n, 0}, tcg::TcStoreAxpby(T, m, 0, 1.f, 0.f), 1, st);
      else
        rc = gemm_simt(N, m, n, 1, LoadSqRows{Ah, n}, LoadDinvT{D, n}, StoreAxpby{T, m, 0, 1.f, 0.f}, st);
      if (rc) return rc;
      if (o.has_bias) {
        inv_shift_kernel<<<(unsigned)ceil_div64(m, 128), 128, 0, st>>>(l1, m, delta_b[i], 0, db);
        BNNP_LAUNCH_CHECK();
      }
      {
        const float* wb = o.has_bias ? db : nullptr;
#define BNNP_CASE(KK) case KK: rc = launch_gg_fused<KK>(Gh, m, T, m, wb, N, m, f_var, st); break;
        switch (simt ? 0 : K) {
          BNNP_CASE(1) BNNP_CASE(2) BNNP_CASE(3) BNNP_CASE(4) BNNP_CASE(5) BNNP_CASE(6) BNNP_CASE(7) BNNP_CASE(8)
          BNNP_CASE(9) BNNP_CASE(10) BNNP_CASE(11) BNNP_CASE(12) BNNP_CASE(13) BNNP_CASE(14) BNNP_CASE(15) BNNP_CASE(16)
          default:
            fvar_weighted_gg_kernel<<<(unsigned)ceil_div64(N * K * K, 128), 128, 0, st>>>(Gh, m, T, m, N, K, m, f_var);
            BNNP_LAUNCH_CHECK();
            if (o.has_bias) {
              fvar_gg_const_kernel<<<(unsigned)ceil_div64(N * K * K, 128), 128, 0, st>>>(Gh, m, db, N, K, m, f_var);
              BNNP_LAUNCH_CHECK();
            }
            rc = BNNP_OK;
        }
#undef BNNP_CASE
        if (rc) return rc;
      }
    } else {
      int64_t ild; const float* in = op_input_g(ops, plan, i, X, ws, &ild);
      if (!in) return BNNP_ERR_INVALID;
      if (R > 65535) return BNNP_ERR_UNSUPPORTED;
      int L = o.out_h * o.out_w;
      int64_t bw = (int64_t)m * n;
      float* Jb = scratch;                 // [R, m, n]
      float* sb = Jb + R * bw;             // [R, m]
      float* rest = sb + R * m;
      if (!glm_force_simt() && (long long)R * m * n * L >= kTcMinMacsK) {
        const int Epad = (n + 3) / 4 * 4;
        float* U = rest + 2 * R * bw + bw + 64;            // behind kron_fvar_block's T1, T2, D
        im2col_g_kernel<<<(unsigned)ceil_div64(N * L * (int64_t)Epad, 256), 256, 0, st>>>(make_im2col_g(o, in, ild, 1), N, L, n,
                                                                                       Epad, U);
        BNNP_LAUNCH_CHECK();
        rc = conv_jac_blocks_tc(G, q.grad_ld, U, Epad, R, K, m, n, L, Jb, st);
      } else {
        rc = gemm_simt(m, n, L, (int)R, ConvGradA2{G, q.grad_ld, L}, ConvUnfoldB2{make_im2col_g(o, in, ild, K)},
                       StoreBlock2{Jb, bw, n}, st);
      }
      if (rc) return rc;
      rc = kron_fvar_block(Jb, bw, R, K, m, n, W1, l1, W2, l2, delta_w[i], dampen, f_var, rest, st);
      if (rc) return rc;
      if (o.has_bias) {
        conv_possum2_kernel<<<(unsigned)ceil_div64(R * m, 128), 128, 0, st>>>(G, q.grad_ld, R, m, L, sb);
        BNNP_LAUNCH_CHECK();
        rc = kron_fvar_block(sb, m, R, K, m, 0, W1, l1, nullptr, nullptr, delta_b[i], dampen, f_var, rest, st);
        if (rc) return rc;
      }
    }
  }
  return BNNP_OK;
}

// ------------------------------------------------------------------------------------ sample
extern "C" int bnnp_glm_sample(int lik, const float* f_mu, const float* f_var, const float* eps,
                               uint64_t seed, int64_t S, int64_t N, int K, float* out, int32_t* status,
                               void* stream) {
  if (!f_mu || !f_var || !out || !status || S <= 0 || N <= 0 || K <= 0 || K > BNNP_MAX_K_WIDE) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  if (K > 16) return glm_sample_wide(lik, f_mu, f_var, eps, seed, S, N, K, out, status, 0, 0, st);
#define BNNP_CASE(KK) case KK: return launch_sample<KK>(lik, f_mu, f_var, eps, seed, S, N, out, status, false, st);
  switch (K) {
    BNNP_CASE(1) BNNP_CASE(2) BNNP_CASE(3) BNNP_CASE(4) BNNP_CASE(5) BNNP_CASE(6) BNNP_CASE(7) BNNP_CASE(8)
    BNNP_CASE(9) BNNP_CASE(10) BNNP_CASE(11) BNNP_CASE(12) BNNP_CASE(13) BNNP_CASE(14) BNNP_CASE(15) BNNP_CASE(16)
    default: return BNNP_ERR_UNSUPPORTED;
  }
#undef BNNP_CASE
}

extern "C" int bnnp_glm_sample_mean(int lik, const float* f_mu, const float* f_var, const float* eps,
                                    uint64_t seed, int64_t S, int64_t N, int K, float* mean, int32_t* status,
                                    void* stream) {
  if (!f_mu || !f_var || !mean || !status || S <= 0 || N <= 0 || K <= 0 || K > BNNP_MAX_K_WIDE) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  if (K > 16) return glm_sample_wide(lik, f_mu, f_var, eps, seed, S, N, K, mean, status, 1, 0, st);
#define BNNP_CASE(KK) case KK: return launch_sample<KK>(lik, f_mu, f_var, eps, seed, S, N, mean, status, true, st);
  switch (K) {
    BNNP_CASE(1) BNNP_CASE(2) BNNP_CASE(3) BNNP_CASE(4) BNNP_CASE(5) BNNP_CASE(6) BNNP_CASE(7) BNNP_CASE(8)
    BNNP_CASE(9) BNNP_CASE(10) BNNP_CASE(11) BNNP_CASE(12) BNNP_CASE(13) BNNP_CASE(14) BNNP_CASE(15) BNNP_CASE(16)
    default: return BNNP_ERR_UNSUPPORTED;
  }
#undef BNNP_CASE
}

extern "C" int bnnp_glm_sample_diag_fallback(int lik, const float* f_mu, const float* f_var, const float* eps,
                                             uint64_t seed, int64_t S, int64_t N, int K, float* out,
                                             void* stream) {
  if (!f_mu || !f_var || !out || S <= 0 || N <= 0 || K <= 0 || K > BNNP_MAX_K_WIDE) return BNNP_ERR_INVALID;
  if (K > BNNP_MAX_K) return glm_sample_wide(lik, f_mu, f_var, eps, seed, S, N, K, out, nullptr, 0, 1, as_stream(stream));
  glm_sample_diag_kernel<<<(unsigned)ceil_div64(S * N, 128), 128, 0, as_stream(stream)>>>(lik, f_mu, f_var, eps, seed,
                                                                                       S, N, K, out);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}
