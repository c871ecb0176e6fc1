// Callers either side of the GGN kernel, Linear-only networks:
//   * bnnp_jtr_accum:          grad[p] += sum_{n,k} J[(n,k),p] r[n,k]      preds/optimizers.py:96
//   * bnnp_glm_linear_samples: out[s,n,k] = f[n,k] + sum_p J[(n,k),p] dtheta[s,p]
//                                                                         preds/predictives.py:31-47
// Both consume the layer factors (A_cat, G_cat) that bnnp_net_forward + the identity-seeded
// bnnp_net_backward leave in the workspace; J is never materialised.  With J[(n,k),(u,j)] =
// G[(n,k),u] a[n,j] the first is one [m x N]·[N x nin] product per layer on the k-reduced gradient,
// the second one batched (over samples) layer product followed by a contraction with G over units --
// K times fewer flops than the reference's `Js @ samples[i]`.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace bnnp {

static bool linear_only_l(const bnnp_op_t* ops, int n_ops) {
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind == BNNP_OP_CONV2D || ops[i].kind == BNNP_OP_MAXPOOL2D || ops[i].kind == BNNP_OP_AVGPOOL2D)
      return false;
  return true;
}
constexpr long long kTcMinMacsL = 1LL << 25;

// gr[n,u] = sum_k G_cat[(n*K+k), u] * r[n,k]
__global__ void jtr_reduce_kernel(const float* __restrict__ G, int64_t gld, const float* __restrict__ r, int K,
                                  int64_t N, int Mtot, float* __restrict__ gr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Mtot) return;
  const int64_t n = i / Mtot;
  const int u = (int)(i - n * Mtot);
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(G[(n * K + k) * gld + u], r[n * K + k], acc);
  gr[i] = acc;
}
// c[n,u] = sum_{k,l} G_cat[(n,k),u] * Lambda[n,k,l] * G_cat[(n,l),u]      (diagonal of J^T Lambda J, unit part)
__global__ void diag_lambda_reduce_kernel(const float* __restrict__ G, int64_t gld, const float* __restrict__ Lam,
                                          int K, int64_t N, int Mtot, float* __restrict__ c) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Mtot) return;
  const int64_t n = i / Mtot;
  const int u = (int)(i - n * Mtot);
  const float* L = Lam + n * K * K;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float gk = G[(n * K + k) * gld + u];
    float t = 0.f;
    for (int l = 0; l < K; ++l) t = fmaf(L[k * K + l], G[(n * K + l) * gld + u], t);
    acc = fmaf(gk, t, acc);
  }
  c[i] = acc;
}
struct LoadRowMajorBSq {       // B(k,n) = p[k*ld + n]^2
  static constexpr bool K_CONTIG = false;
  const float* p; int64_t ld;
  __device__ float operator()(int64_t k, int64_t n, int) const { const float v = p[k * ld + n]; return v * v; }
};
struct TcColMajorSqL {         // X(j, k = example) = p[k*ld + j]^2
  static constexpr bool K_CONTIG = false;
  using Row = const float*;
  const float* p; long long ld;
  __device__ __forceinline__ Row row(int64_t r, int) const { return p + r; }
  __device__ __forceinline__ void load8(const Row& rw, int64_t k0, int64_t K, float* out) const {
    const float* q = rw + k0 * ld;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float v = (k0 + e < K) ? __ldg(q + (long long)e * ld) : 0.f;
      out[e] = v * v;
    }
  }
};

// bias gradient: grad[offb + u] += sum_n gr[n, unit0 + u]      one warp per unit
__global__ void jtr_bias_kernel(const float* __restrict__ gr, int Mtot, int unit0, int m, int64_t N,
                                float* __restrict__ grad_b) {
  const int w = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= m) return;
  float acc = 0.f;
  for (int64_t n = lane; n < N; n += 32) acc += gr[n * Mtot + unit0 + w];
  acc = warp_sum(acc);
  if (lane == 0) grad_b[w] += acc;
}

// U[s, n, unit0+i] = acc + dtheta[s, offb + i]
struct StoreUnitSimt {
  float* U; int64_t bs, ld; const float* db; int64_t P; int has_bias;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t n, int64_t i, float acc, int s) const {
    U[(int64_t)s * bs + n * ld + i] = acc + (has_bias ? db[(int64_t)s * P + i] : 0.f);
  }
};
struct StoreUnitTc : tcg::TcNoState {
  float* U; long long bs, ld; const float* db; long long P; int has_bias;
  __host__ __device__ StoreUnitTc(float* U_, long long bs_, long long ld_, const float* db_, long long P_, int hb_)
      : U(U_), bs(bs_), ld(ld_), db(db_), P(P_), has_bias(hb_) {}
  __device__ __forceinline__ void row(State&, int64_t n, int64_t i0, const float* v, int nvalid, int s, int) const {
    float* q = U + (long long)s * bs + n * ld + i0;
    const float* b = db + (long long)s * P + i0;
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e < nvalid) q[e] = v[e] + (has_bias ? __ldg(b + e) : 0.f);
  }
};
// out[s, n, k] = f[n,k] + sum_u G[(n,k),u] * U[s,n,u]       one warp per (s, n, k)
__global__ void __launch_bounds__(256)
linpred_contract_kernel(const float* __restrict__ G, int64_t gld, const float* __restrict__ U, int Mtot,
                        const float* __restrict__ f, int64_t N, int K, int64_t S, float* __restrict__ out) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= S * N * K) return;
  const int64_t r = w % (N * K);            // (n,k)
  const int64_t s = w / (N * K);
  const float* g = G + r * gld;
  const float* u = U + (s * N + r / K) * Mtot;
  float acc = 0.f;
  for (int c = lane; c < Mtot; c += 32) acc = fmaf(g[c], u[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[w] = f[r] + acc;
}

}  // namespace bnnp
using namespace bnnp;

extern "C" int64_t bnnp_jtr_scratch_floats(const bnnp_plan_t* plan, int64_t N) {
  return plan ? N * (int64_t)plan->Mtot + 64 : 0;
}

extern "C" int bnnp_jtr_accum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                              const float* ws, const float* rs, float* grad, float* scratch, void* stream) {
  if (!ops || !plan || !ws || !rs || !grad || !scratch || N <= 0) return BNNP_ERR_INVALID;
  if (!linear_only_l(ops, n_ops)) return BNNP_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int K = plan->K, Mtot = plan->Mtot;
  float* gr = scratch;
  jtr_reduce_kernel<<<(unsigned)ceil_div64(N * Mtot, 256), 256, 0, st>>>(ws + plan->gcat_off, Mtot, rs, K, N, Mtot, gr);
  BNNP_LAUNCH_CHECK();
  const float* A = ws + plan->acat_off;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    // grad_W[u, j] += sum_n gr[n, unit0+u] * a[n, col0+j]; bounded accumulation through grad for long N
    int rc;
    if ((long long)o.out_c * o.in_c * N >= kTcMinMacsL)
      rc = tcg::gemm_tc_ksegments(o.out_c, o.in_c, N, 1, tcg::TcColMajor{gr + q.lin_unit, Mtot, 0},
                                  tcg::TcColMajor{A + q.lin_col, plan->Dtot, 0},
                                  [=](int) { return tcg::TcStoreAxpby(grad + o.theta_off_w, o.in_c, 0, 1.f, 1.f); }, st);
    else
      rc = gemm_simt(o.out_c, o.in_c, N, 1, LoadColMajorA{gr + q.lin_unit, Mtot, 0},
                     LoadRowMajorB{A + q.lin_col, plan->Dtot, 0}, StoreAxpby{grad + o.theta_off_w, o.in_c, 0, 1.f, 1.f}, st);
    if (rc) return rc;
    if (o.has_bias) {
      jtr_bias_kernel<<<(unsigned)ceil_div64((int64_t)o.out_c * 32, 256), 256, 0, st>>>(gr, Mtot, q.lin_unit, o.out_c, N,
                                                                                       grad + o.theta_off_b);
      BNNP_LAUNCH_CHECK();
    }
  }
  return BNNP_OK;
}

/* diag[p] += diag(sum_n J_n^T Lambda_n J_n)[p] for an arbitrary per-example Lambda [N,K,K]: the diagonal GGN of the
 * GLM refinement loops (preds/refine.py:93-121), where Lambda is re-evaluated every iteration on fixed factors. */
extern "C" int bnnp_ggn_diag_lambda(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                                    const float* ws, const float* Lambda, float* diag, float* scratch, void* stream) {
  if (!ops || !plan || !ws || !Lambda || !diag || !scratch || N <= 0) return BNNP_ERR_INVALID;
  if (!linear_only_l(ops, n_ops)) return BNNP_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int K = plan->K, Mtot = plan->Mtot;
  float* c = scratch;                                  // [N, Mtot], same size as bnnp_jtr_scratch_floats()
  diag_lambda_reduce_kernel<<<(unsigned)ceil_div64(N * Mtot, 256), 256, 0, st>>>(ws + plan->gcat_off, Mtot, Lambda, K, N,
                                                                                 Mtot, c);
  BNNP_LAUNCH_CHECK();
  const float* A = ws + plan->acat_off;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    int rc;
    if ((long long)o.out_c * o.in_c * N >= kTcMinMacsL)
      rc = tcg::gemm_tc_ksegments(o.out_c, o.in_c, N, 1, tcg::TcColMajor{c + q.lin_unit, Mtot, 0},
                                  TcColMajorSqL{A + q.lin_col, plan->Dtot},
                                  [=](int) { return tcg::TcStoreAxpby(diag + o.theta_off_w, o.in_c, 0, 1.f, 1.f); }, st);
    else
      rc = gemm_simt(o.out_c, o.in_c, N, 1, LoadColMajorA{c + q.lin_unit, Mtot, 0},
                     LoadRowMajorBSq{A + q.lin_col, plan->Dtot}, StoreAxpby{diag + o.theta_off_w, o.in_c, 0, 1.f, 1.f}, st);
    if (rc) return rc;
    if (o.has_bias) {
      jtr_bias_kernel<<<(unsigned)ceil_div64((int64_t)o.out_c * 32, 256), 256, 0, st>>>(c, Mtot, q.lin_unit, o.out_c, N,
                                                                                       diag + o.theta_off_b);
      BNNP_LAUNCH_CHECK();
    }
  }
  return BNNP_OK;
}

extern "C" int64_t bnnp_glm_linear_samples_scratch_floats(const bnnp_plan_t* plan, int64_t N, int64_t S) {
  return plan ? S * N * (int64_t)plan->Mtot + 64 : 0;
}

extern "C" int bnnp_glm_linear_samples(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                                       const float* ws, const float* f, const float* dtheta, int64_t S,
                                       float* out, float* scratch, void* stream) {
  if (!ops || !plan || !ws || !f || !dtheta || !out || !scratch || N <= 0 || S <= 0) return BNNP_ERR_INVALID;
  if (!linear_only_l(ops, n_ops)) return BNNP_ERR_UNSUPPORTED;
  if (S > 65535) return BNNP_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int K = plan->K, Mtot = plan->Mtot;
  const int64_t P = plan->P;
  float* U = scratch;                                    // [S, N, Mtot]
  const float* A = ws + plan->acat_off;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    const float* dW = dtheta + o.theta_off_w;
    const float* db = o.has_bias ? dtheta + o.theta_off_b : dtheta;
    int rc;
    if ((long long)N * o.out_c * o.in_c * S >= kTcMinMacsL && o.in_c <= 4096)
      rc = tcg::gemm_tc(N, o.out_c, o.in_c, (int)S, tcg::TcRowMajor{A + q.lin_col, plan->Dtot, 0},
                        tcg::TcRowMajor{dW, o.in_c, P},
                        StoreUnitTc(U + q.lin_unit, N * (long long)Mtot, Mtot, db, P, o.has_bias), 1, st);
    else
      rc = gemm_simt(N, o.out_c, o.in_c, (int)S, LoadRowMajorA{A + q.lin_col, plan->Dtot, 0},
                     LoadColMajorB{dW, o.in_c, P},
                     StoreUnitSimt{U + q.lin_unit, N * (int64_t)Mtot, Mtot, db, P, o.has_bias}, st);
    if (rc) return rc;
  }
  const int64_t warps = S * N * K;
  linpred_contract_kernel<<<(unsigned)ceil_div64(warps * 32, 256), 256, 0, st>>>(ws + plan->gcat_off, Mtot, U, Mtot, f,
                                                                                 N, K, S, out);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}
