// GGN accumulation from the saved layer factors (a_l, g_l) -- the Jacobian J_n is never read
// from HBM; its tiles are synthesised inside the contraction kernels.
//   full : H += sum_n J_n^T Lambda_n J_n            preds/laplace.py:42, preds/optimizers.py:94
//   diag : d += factor * diag(...)                   preds/laplace.py:52-59 (BackPACK DiagGGNExact)
//   kron : G_l += factor * sum s s^T, A_l += bf/N * sum a a^T   preds/laplace.py:73-78, kron.py:58-69
#include "common.cuh"
#include "gemm_simt.cuh"
#include "tc_common.cuh"
#include "gemm_tc.cuh"
#include "gemm_tma.cuh"
#include "gram_tc.cuh"
#include "conv_diag_tc.cuh"

namespace bnnp {

// ------------------------------------------------------------------------------------------
// shared functors
// ------------------------------------------------------------------------------------------
// implicit im2col of a conv input: U(pos, e) with e = (ci, a, b), pos = (oh, ow); batch b -> example b / V
using Im2col = ConvUnfold;      // (conv_diag_tc.cuh)
static Im2col make_im2col(const bnnp_op_t& o, const float* in, int64_t ild, int V) {
  return Im2col{in, ild, V, o.in_c, o.in_h, o.in_w, o.out_w, o.kh, o.kw, o.stride_h, o.stride_w,
                o.pad_l, o.pad_t};
}

static const float* op_input(const bnnp_op_t* ops, const bnnp_plan_t* plan, int i, const float* X,
                             const float* ws, int64_t* ld) {
  if (i == 0) { *ld = op_in_size(ops[0]); return X; }
  *ld = plan->op[i - 1].act_ld;
  return ws + plan->op[i - 1].act_off;
}

// ------------------------------------------------------------------------------------------
// Jacobian materialisation  Js[n, p, v]
// ------------------------------------------------------------------------------------------
// grid (ceil(per / 256), N): the example is the block row and the parameter index a 32-bit quantity (a flat 64-bit element
// index costs three 64-bit divisions per element and made this pass instruction-bound)
__global__ void jac_linear_kernel(const float* __restrict__ G, int64_t gld, const float* __restrict__ A,
                                  int64_t ald, int64_t N, int V, int m, int nin, int has_bias,
                                  int64_t offw, int64_t offb, int64_t P, float* __restrict__ Js) {
  const int per = m * (nin + (has_bias ? 1 : 0));
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per) return;
  const int64_t n = blockIdx.y;
  int u, j = -1;
  int64_t p;
  if (e < m * nin) { u = e / nin; j = e - u * nin; p = offw + e; }
  else { u = e - m * nin; p = offb + u; }
  const float a = j >= 0 ? A[n * ald + j] : 1.f;
  const float* g = G + n * V * gld + u;
  float* out = Js + (n * P + p) * V;
  for (int v = 0; v < V; ++v) out[v] = g[v * gld] * a;
}

struct ConvGradA {     // A(co, pos) = gout[r, co, pos]
  static constexpr bool K_CONTIG = true;
  const float* g; int64_t gld; int L;
  __device__ float operator()(int64_t co, int64_t pos, int r) const { return g[(int64_t)r * gld + co * L + pos]; }
};
struct ConvUnfoldB {   // B(pos, e) = unfold(in[n])[e, pos]
  static constexpr bool K_CONTIG = false;
  Im2col u;
  __device__ float operator()(int64_t pos, int64_t e, int r) const { return u.at(r / u.V, (int)pos, (int)e); }
};
struct StoreJacConv {  // Js[n, offw + co*E + e, v]
  float* Js; int64_t P, offw; int V, E;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t co, int64_t e, float acc, int r) const {
    int64_t n = r / V; int v = r % V;
    Js[(n * P + offw + co * E + e) * V + v] = acc;
  }
};
// s[r, co] = sum_pos gout[r, co, pos]
__global__ void conv_possum_kernel(const float* __restrict__ g, int64_t gld, int64_t R, int Cout,
                                   int L, float* __restrict__ s, int64_t sld) {
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (w >= R * Cout) return;
  int64_t r = w / Cout; int co = (int)(w % Cout);
  const float* p = g + r * gld + (int64_t)co * L;
  float acc = 0.f;
  for (int i = lane; i < L; i += 32) acc += p[i];
  acc = warp_sum(acc);
  if (lane == 0) s[r * sld + co] = acc;
}

// the same for short rows (L <= 128: late conv layers, e.g. 6 x 6 maps): eight lanes per row, four rows per warp -- a warp per
// row keeps one 4-byte load per lane in flight and the launch is latency-bound (95 us for 94 MB on CIFAR10Net's conv3)
__global__ void conv_possum_short_kernel(const float* __restrict__ g, int64_t gld, int64_t R, int Cout,
                                         int L, float* __restrict__ s, int64_t sld) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  float acc = 0.f;
  const bool on = w < R * Cout;
  const int64_t r = on ? w / Cout : 0; const int co = on ? (int)(w % Cout) : 0;
  if (on) {
    const float* p = g + r * gld + (int64_t)co * L;
#pragma unroll 4
    for (int i = sub; i < L; i += 8) acc += __ldg(p + i);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (on && sub == 0) s[r * sld + co] = acc;
}
static void launch_conv_possum(const float* g, int64_t gld, int64_t R, int Cout, int L, float* s, int64_t sld, cudaStream_t st) {
  if (L <= 128)
    conv_possum_short_kernel<<<(unsigned)ceil_div64(R * Cout * 8, 256), 256, 0, st>>>(g, gld, R, Cout, L, s, sld);
  else
    conv_possum_kernel<<<(unsigned)ceil_div64(R * Cout * 32, 256), 256, 0, st>>>(g, gld, R, Cout, L, s, sld);
}

// Js[n, offb + co, v] = sum_pos gout[(n,v), co, pos]
__global__ void jac_conv_bias_direct(const float* __restrict__ g, int64_t gld, int64_t N, int V, int Cout,
                                     int L, int64_t offb, int64_t P, float* __restrict__ Js) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * V * Cout) return;
  int co = (int)(i % Cout); int64_t r = i / Cout; int64_t n = r / V; int v = (int)(r % V);
  const float* p = g + r * gld + (int64_t)co * L;
  float acc = 0.f;
  for (int k = 0; k < L; ++k) acc += p[k];
  Js[(n * P + offb + co) * V + v] = acc;
}

// ------------------------------------------------------------------------------------------
// full GGN, SIMT K-form:  H[p,q] += sum_{n,k} (Lambda G)[n,k,u(p)] a[n,c(p)] * G[n,k,u(q)] a[n,c(q)]
// ------------------------------------------------------------------------------------------
// per-parameter maps for Linear-only networks: unit column in G_cat, input column in A_cat
__global__ void param_maps_kernel(int32_t* __restrict__ pu, int32_t* __restrict__ pc, int64_t off,
                                  int m, int nin, int unit0, int col0, int is_bias) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t cnt = is_bias ? m : (int64_t)m * nin;
  if (e >= cnt) return;
  if (is_bias) { pu[off + e] = unit0 + (int)e; pc[off + e] = col0 + nin; }
  else { pu[off + e] = unit0 + (int)(e / nin); pc[off + e] = col0 + (int)(e % nin); }
}

// GL[(n,k), u] = sum_l Lambda[n,k,l] G[(n,l), u]
__global__ void lambda_g_kernel(const float* __restrict__ Lam, const float* __restrict__ G,
                                int64_t N, int K, int M, float* __restrict__ GL) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K * M) return;
  int u = (int)(i % M); int64_t r = i / M; int64_t n = r / K; int k = (int)(r % K);
  float acc = 0.f;
  for (int l = 0; l < K; ++l) acc = fmaf(Lam[(n * K + k) * K + l], G[(n * K + l) * M + u], acc);
  GL[i] = acc;
}

struct JacOperand {       // J-like operand: X(p, r) = Gx[r, u(p)] * A[r / K, c(p)]
  static constexpr bool K_CONTIG = false;
  const float* Gx; const float* A; const int32_t* pu; const int32_t* pc;
  int64_t gld, ald; int K;
  __device__ float operator()(int64_t p, int64_t r, int) const {
    return Gx[r * gld + pu[p]] * A[(r / K) * ald + pc[p]];
  }
};
struct JacOperandB {      // same, (k, n) argument order
  static constexpr bool K_CONTIG = false;
  JacOperand j;
  __device__ float operator()(int64_t r, int64_t q, int) const { return j(q, r, 0); }
};
struct AccumLower {
  float* H; int64_t P;
  __device__ bool skip(int64_t m0, int64_t n0, int bm, int) const { return m0 + bm - 1 < n0; }
  __device__ void operator()(int64_t m, int64_t n, float acc, int) const { H[m * P + n] += acc; }
};

// mirror of source rows [row0, row1): H[c, r] = H[r, c] for r in [row0, row1), c < r.
// grid: x = 32-column tiles 0 .. (row1-1)/32, y = 32-row tiles of the range (source tile = rows by, cols bx)
__global__ void symmetrize_kernel(float* __restrict__ H, int64_t P, int64_t row0, int64_t row1) {
  __shared__ float tile[32][33];
  const int64_t bx = blockIdx.x, by = row0 / 32 + blockIdx.y;     // source tile: rows by*32.., cols bx*32..
  if (bx > by) return;
  int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += 8) {
    int64_t sr = by * 32 + r, sc = bx * 32 + tx;
    tile[r][tx] = (sr >= row0 && sr < row1 && sc < P) ? H[sr * P + sc] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int64_t dr = bx * 32 + r, dc = by * 32 + tx;                  // destination: row = source col, col = source row
    if (dc >= row0 && dc < row1 && dr < dc) H[dr * P + dc] = tile[tx][r];
  }
}

// ------------------------------------------------------------------------------------------
// diag
// ------------------------------------------------------------------------------------------
// Q[n, u] = sum_v G[(n,v), u]^2
__global__ void sqsum_v_kernel(const float* __restrict__ G, int64_t gld, int64_t N, int V, int M,
                               float* __restrict__ Q) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * M) return;
  int u = (int)(i % M); int64_t n = i / M;
  float acc = 0.f;
  for (int v = 0; v < V; ++v) { float g = G[(n * V + v) * gld + u]; acc = fmaf(g, g, acc); }
  Q[i] = acc;
}
struct LoadQT {           // A(i, n) = Q[n, unit0 + i]
  static constexpr bool K_CONTIG = false;
  const float* Q; int64_t ld; int unit0;
  __device__ float operator()(int64_t i, int64_t n, int) const { return Q[n * ld + unit0 + i]; }
};
struct LoadSq {           // B(n, j) = A[n, col0 + j]^2
  static constexpr bool K_CONTIG = false;
  const float* A; int64_t ld; int col0;
  __device__ float operator()(int64_t n, int64_t j, int) const { float a = A[n * ld + col0 + j]; return a * a; }
};
struct AccumLinearParam { // (i, j) -> weight [m, nin] at offw, column nin -> bias at offb
  float* d; int64_t offw, offb; int nin; int has_bias; float factor;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t i, int64_t j, float acc, int) const {
    if (j < nin) d[offw + i * nin + j] += factor * acc;
    else if (has_bias) d[offb + i] += factor * acc;
  }
};

struct TcColMajorSq {      // X(j, k = example n) = A[n*ld + j]^2   (diag path: squared layer inputs)
  static constexpr bool K_CONTIG = false;
  using Row = const float*;
  const float* p; long long ld;
  __device__ __forceinline__ Row row(int64_t r, int) const { return p + r; }
  __device__ __forceinline__ void load8(const Row& rw, int64_t k0, int64_t K, float* out) const {
    const float* q = rw + k0 * ld;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = (k0 + e < K) ? __ldg(q + (long long)e * ld) : 0.f;
      out[e] = v * v;
    }
  }
};
// d[offw + i*nin + j] += factor * T[i, j] (j < nin),  d[offb + i] += factor * T[i, nin]
__global__ void diag_scatter_kernel(const float* __restrict__ T, int m, int nin, int has_bias, int64_t offw,
                                    int64_t offb, float factor, float* __restrict__ d) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)m * (nin + 1)) return;
  int i = (int)(e / (nin + 1)), j = (int)(e % (nin + 1));
  if (j < nin) d[offw + (int64_t)i * nin + j] += factor * T[e];
  else if (has_bias) d[offb + i] += factor * T[e];
}
constexpr long long kTcMinMacsG = 1LL << 25;

// U[(n,pos), e] = unfold(in[n])[e, pos]   (row pitch Epad = roundup4(E), zero padded)
__global__ void im2col_kernel(Im2col u, int64_t N, int L, int E, int Epad, float* __restrict__ U) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * L * Epad) return;
  const int e = (int)(i % Epad);
  const int64_t rp = i / Epad;
  U[i] = e < E ? u.at(rp / L, (int)(rp % L), e) : 0.f;
}

// Ut_hi/lo[e, rp] = bf16 split of unfold(in[n])[e, pos], rp = n*L + pos (k = rp contiguous, row pitch NLp, zero padded):
// the transposed, pre-split operand of the TMA-fed A-factor Gram  A = U^T U  (BackPACK KFLR, conv layers), written in one
// pass over the layer input instead of an fp32 im2col buffer that a generated-operand GEMM then re-reads.
// One thread per 8 consecutive rp (16-byte stores).
__global__ void im2col_t_split_kernel(Im2col u, int64_t N, int L, int E, int64_t NLp, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ lo) {
  // grid (ceil(NLp / 8 / 256), E): e is uniform per block, (n, oh, ow) of the first of a thread's 8 positions come from two
  // 32-bit divisions and are advanced incrementally (u.at() per element: eight divisions, three of them 64-bit)
  const int e = blockIdx.y;
  const int64_t rp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) << 3;
  if (rp0 >= NLp) return;
  const int b = e % u.kw, t = e / u.kw, a = t % u.kh, ci = t / u.kh;
  int64_t n = rp0 / L;
  int pos = (int)(rp0 - n * L);
  int oh = pos / u.OW, ow = pos - oh * u.OW;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = n < N ? u.at_dec(u.in + n * u.ild, ci, a, b, oh, ow) : 0.f;
    if (++ow == u.OW) { ow = 0; ++oh; }
    if (++pos == L) { pos = 0; oh = 0; ow = 0; ++n; }
  }
  uint4 h, l;
  tcp::split8(v, h, l);
  *reinterpret_cast<uint4*>(hi + (int64_t)e * NLp + rp0) = h;
  *reinterpret_cast<uint4*>(lo + (int64_t)e * NLp + rp0) = l;
}

// ---- conv weight diag on the tensor cores: one [Cout x L]·[L x E] product per row r = (n,v) as a batched gemm_tc whose
// epilogue squares the accumulator and adds it to the partial sum of the CTA that ran the tile (tiles are assigned
// statically, so every partial has a fixed summation order; the partials are reduced afterwards).
struct TcUnfoldT {         // X(e, k = pos) of batch r: U[((r / V) * L + pos) * Epad + e]
  static constexpr bool K_CONTIG = false;
  using Row = const float*;
  const float* U; long long Epad; int L, V;
  __device__ __forceinline__ Row row(int64_t e, int b) const { return U + (long long)(b / V) * L * Epad + e; }
  __device__ __forceinline__ void load8(const Row& rw, int64_t k0, int64_t K, float* out) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = (k0 + j < K) ? __ldg(rw + (k0 + j) * Epad) : 0.f;
  }
};
struct TcSqAccumPart : tcg::TcNoState {     // part[cta][e][co] += acc^2
  // The epilogue thread owns one accumulator row (co) and 32 consecutive columns (e): with the partials stored
  // column-major (co contiguous) the 32 lanes of a warp touch 128 contiguous bytes per access; row-major partials
  // cost 32 L1 wavefronts per instruction (ncu: generators idle on `empty`, tensor pipe 1.4 %, 52 us per tile).
  float* part; long long stride; int MP;
  __host__ __device__ TcSqAccumPart(float* p_, long long s_, int MP_) : part(p_), stride(s_), MP(MP_) {}
  __device__ __forceinline__ void row(State&, int64_t m, int64_t n0, const float* v, int nvalid, int, int) const {
    float* __restrict__ q = part + (long long)blockIdx.x * stride + n0 * MP + m;
    float h[32];                             // all 32 loads in flight before the first dependent store
#pragma unroll
    for (int e = 0; e < 32; ++e) h[e] = e < nvalid ? q[(long long)e * MP] : 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e < nvalid) q[(long long)e * MP] = fmaf(v[e], v[e], h[e]);
  }
};
// d[co*E + e] += factor * sum_cta part[cta][e*MP + co]   (fixed summation order)
__global__ void reduce_partials_t_kernel(const float* __restrict__ part, int ncta, long long stride, int Cout, int E,
                                         int MP, float factor, float* __restrict__ d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over [E][Cout], co fastest (coalesced reads)
  if (i >= (int64_t)E * Cout) return;
  const int co = (int)(i % Cout), e = (int)(i / Cout);
  float acc = 0.f;
  for (int z = 0; z < ncta; ++z) acc += part[(long long)z * stride + (long long)e * MP + co];
  d[(long long)co * E + e] += factor * acc;
}

// conv weight diag (SIMT, small problems): d[co, e] += factor * sum_r ( sum_pos gout[r,co,pos] U[r/V, e, pos] )^2
// one block per 64x64 output tile, looping over all rows r (the per-sample Jacobian tile is
// formed in registers, squared and accumulated -- never stored).
__global__ void __launch_bounds__(G_THREADS)
conv_diag_kernel(ConvGradA la, ConvUnfoldB lb, int64_t R, int Cout, int E, int L,
                 float* __restrict__ partial) {
  __shared__ float As[G_BK][G_BM + 4];
  __shared__ float Bs[G_BK][G_BN + 4];
  const int64_t m0 = (int64_t)blockIdx.y * G_BM, n0 = (int64_t)blockIdx.x * G_BN;
  const int t = threadIdx.x;
  const int tx = t % (G_BN / G_TN), ty = t / (G_BN / G_TN);
  float tot[G_TM][G_TN];
#pragma unroll
  for (int i = 0; i < G_TM; ++i)
#pragma unroll
    for (int j = 0; j < G_TN; ++j) tot[i][j] = 0.f;
  const int64_t rchunk = (R + gridDim.z - 1) / gridDim.z;
  const int64_t rbeg = blockIdx.z * rchunk, rend = rbeg + rchunk < R ? rbeg + rchunk : R;
  for (int64_t r = rbeg; r < rend; ++r) {
    float acc[G_TM][G_TN];
#pragma unroll
    for (int i = 0; i < G_TM; ++i)
#pragma unroll
      for (int j = 0; j < G_TN; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < L; k0 += G_BK) {
#pragma unroll
      for (int i = 0; i < (G_BM * G_BK) / G_THREADS; ++i) {
        int e = t + i * G_THREADS, kk = e % G_BK, mm = e / G_BK;
        int64_t m = m0 + mm; int k = k0 + kk;
        As[kk][mm] = (m < Cout && k < L) ? la(m, k, (int)r) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < (G_BN * G_BK) / G_THREADS; ++i) {
        int e = t + i * G_THREADS, nn = e % G_BN, kk = e / G_BN;
        int64_t n = n0 + nn; int k = k0 + kk;
        Bs[kk][nn] = (n < E && k < L) ? lb(k, n, (int)r) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < G_BK; ++kk) {
        float a[G_TM], b[G_TN];
#pragma unroll
        for (int i = 0; i < G_TM; ++i) a[i] = As[kk][ty * G_TM + i];
#pragma unroll
        for (int j = 0; j < G_TN; ++j) b[j] = Bs[kk][tx * G_TN + j];
#pragma unroll
        for (int i = 0; i < G_TM; ++i)
#pragma unroll
          for (int j = 0; j < G_TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < G_TM; ++i)
#pragma unroll
      for (int j = 0; j < G_TN; ++j) tot[i][j] = fmaf(acc[i][j], acc[i][j], tot[i][j]);
  }
#pragma unroll
  for (int i = 0; i < G_TM; ++i)
#pragma unroll
    for (int j = 0; j < G_TN; ++j) {
      int64_t m = m0 + ty * G_TM + i, n = n0 + tx * G_TN + j;
      if (m < Cout && n < E) partial[((int64_t)blockIdx.z * Cout + m) * E + n] = tot[i][j];
    }
}

// d[i] += factor * sum_z partial[z, i]   (fixed summation order: deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nz, int64_t n, float factor,
                                       float* __restrict__ d) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int z = 0; z < nz; ++z) acc += partial[(int64_t)z * n + i];
  d[i] += factor * acc;
}

// d[offb + co] += factor * sum_r s[r, co]^2
// block (32, 32): 32 columns x 32 row phases, combined in a fixed order (one thread per column walking all R rows took
// 0.29 ms per conv layer at R = 5120: 9 % of the cfg4 diag pass for a 1.3 MB input)
__global__ void colsq_accum_kernel(const float* __restrict__ s, int64_t R, int C, float factor,
                                   float* __restrict__ d) {
  __shared__ float red[32][33];
  const int co = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (co < C)
    for (int64_t r = threadIdx.y; r < R; r += 32) { const float v = s[r * C + co]; acc = fmaf(v, v, acc); }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && co < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) t += red[j][threadIdx.x];
    d[co] += factor * t;
  }
}

// ------------------------------------------------------------------------------------------
// KFAC
// ------------------------------------------------------------------------------------------
struct LoadTA {            // A(i, r) = p[r*ld + c0 + i]
  static constexpr bool K_CONTIG = false;
  const float* p; int64_t ld; int c0;
  __device__ float operator()(int64_t i, int64_t r, int) const { return p[r * ld + c0 + i]; }
};
struct LoadNB {            // B(r, j) = p[r*ld + c0 + j]
  static constexpr bool K_CONTIG = false;
  const float* p; int64_t ld; int c0;
  __device__ float operator()(int64_t r, int64_t j, int) const { return p[r * ld + c0 + j]; }
};
struct AccumScaled {
  float* C; int64_t ld; float alpha;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t m, int64_t n, float acc, int) const { C[m * ld + n] += alpha * acc; }
};
struct UnfoldTA {          // A(e, (n,pos)) = unfold(in[n])[e, pos]
  static constexpr bool K_CONTIG = false;
  Im2col u; int L;
  __device__ float operator()(int64_t e, int64_t r, int) const { return u.at(r / L, (int)(r % L), (int)e); }
};
struct UnfoldNB {
  static constexpr bool K_CONTIG = false;
  Im2col u; int L;
  __device__ float operator()(int64_t r, int64_t e, int) const { return u.at(r / L, (int)(r % L), (int)e); }
};

}  // namespace bnnp
using namespace bnnp;

static bool linear_only(const bnnp_op_t* ops, int n_ops) {
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind == BNNP_OP_CONV2D || ops[i].kind == BNNP_OP_MAXPOOL2D || ops[i].kind == BNNP_OP_AVGPOOL2D)
      return false;
  return true;
}

static int build_param_maps(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int32_t* pu,
                            int32_t* pc, cudaStream_t st) {
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    int64_t cnt = (int64_t)o.out_c * o.in_c;
    param_maps_kernel<<<(unsigned)ceil_div64(cnt, 256), 256, 0, st>>>(pu, pc, o.theta_off_w, o.out_c, o.in_c,
                                                                       q.lin_unit, q.lin_col, 0);
    BNNP_LAUNCH_CHECK();
    if (o.has_bias) {
      param_maps_kernel<<<(unsigned)ceil_div64(o.out_c, 256), 256, 0, st>>>(pu, pc, o.theta_off_b, o.out_c,
                                                                             o.in_c, q.lin_unit, q.lin_col, 1);
      BNNP_LAUNCH_CHECK();
    }
  }
  return BNNP_OK;
}

extern "C" int bnnp_jacobians(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                              const float* X, const float* ws, float* Js, void* stream) {
  if (!plan) return BNNP_ERR_INVALID;
  return bnnp_jacobians_cols(ops, n_ops, plan, N, plan->K, X, ws, Js, stream);
}

extern "C" int bnnp_jacobians_cols(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, int V,
                                   const float* X, const float* ws, float* Js, void* stream) {
  if (!ops || !plan || !ws || !Js || N <= 0 || V <= 0) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (!is_param_op(o)) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    const float* G = ws + q.grad_off;
    if (o.kind == BNNP_OP_LINEAR) {
      int64_t per = (int64_t)o.out_c * (o.in_c + (o.has_bias ? 1 : 0));
      if (N > 65535 || per >= (1LL << 31)) return BNNP_ERR_UNSUPPORTED;
      jac_linear_kernel<<<dim3((unsigned)ceil_div64(per, 256), (unsigned)N), 256, 0, st>>>(
          G, q.grad_ld, ws + plan->acat_off + q.lin_col, plan->Dtot, N, V, o.out_c, o.in_c, o.has_bias,
          o.theta_off_w, o.theta_off_b, plan->P, Js);
      BNNP_LAUNCH_CHECK();
    } else {
      int64_t ild; const float* in = op_input(ops, plan, i, X, ws, &ild);
      if (!in) return BNNP_ERR_INVALID;
      int L = o.out_h * o.out_w, E = o.in_c * o.kh * o.kw;
      if (N * V > 65535) return BNNP_ERR_UNSUPPORTED;
      int rc = gemm_simt(o.out_c, E, L, (int)(N * V), ConvGradA{G, q.grad_ld, L},
                         ConvUnfoldB{make_im2col(o, in, ild, V)},
                         StoreJacConv{Js, plan->P, o.theta_off_w, V, E}, st);
      if (rc) return rc;
      if (o.has_bias) {
        jac_conv_bias_direct<<<(unsigned)ceil_div64(N * V * o.out_c, 128), 128, 0, st>>>(
            G, q.grad_ld, N, V, o.out_c, L, o.theta_off_b, plan->P, Js);
        BNNP_LAUNCH_CHECK();
      }
    }
  }
  return BNNP_OK;
}

// ---------------------------------------------------------------------------- full
extern "C" int64_t bnnp_ggn_full_scratch_floats(const bnnp_plan_t* plan, int64_t N, int engine) {
  if (!plan) return 0;
  int64_t simt = 2 * ((plan->P + 31) / 32 * 32) + N * plan->K * (int64_t)plan->Mtot + 64;
  int64_t tc = gram_tc_scratch_floats(plan, N);
  (void)engine;
  return simt > tc ? simt : tc;
}

extern "C" int bnnp_ggn_full_accum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                                   const float* ws, const float* Lambda, float* H, float* scratch,
                                   int engine, void* stream) {
  return bnnp_ggn_full_accum_rows(ops, n_ops, plan, N, ws, Lambda, H, scratch, engine, 0, -1, 0, stream);
}

extern "C" int bnnp_ggn_full_row_chunks(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, int engine,
                                        int n_chunks, int64_t* bounds) {
  if (!ops || !plan || !bounds || n_chunks < 1 || N <= 0) return BNNP_ERR_INVALID;
  if (engine == 0) engine = gram_tc_preferred(ops, n_ops, plan, N) ? 2 : 1;
  if (engine != 2) { bounds[0] = 0; bounds[1] = plan->P; return 1; }       // the SIMT engine is not row-chunked
  return gram_row_chunks(ops, n_ops, plan, n_chunks, bounds);
}

extern "C" int bnnp_ggn_full_accum_rows(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                                        const float* ws, const float* Lambda, float* H, float* scratch, int engine,
                                        int64_t row_begin, int64_t row_end, int prepared, void* stream) {
  if (!ops || !plan || !ws || !Lambda || !H || !scratch || N <= 0) return BNNP_ERR_INVALID;
  if (!linear_only(ops, n_ops)) return BNNP_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  if (engine == 0) engine = gram_tc_preferred(ops, n_ops, plan, N) ? 2 : 1;
  if (engine == 2) return gram_tc_full_accum(ops, n_ops, plan, N, ws, Lambda, H, scratch, row_begin, row_end, prepared, st);
  if (!(row_end < 0 || (row_begin == 0 && row_end == plan->P))) return BNNP_ERR_UNSUPPORTED;   // SIMT engine: whole matrix only
  const int K = plan->K, M = plan->Mtot;
  const int64_t P = plan->P, Ppad = (P + 31) / 32 * 32;
  int32_t* pu = reinterpret_cast<int32_t*>(scratch);
  int32_t* pc = pu + Ppad;
  float* GL = scratch + 2 * Ppad;
  int rc = build_param_maps(ops, n_ops, plan, pu, pc, st);
  if (rc) return rc;
  const float* G = ws + plan->gcat_off;
  const float* A = ws + plan->acat_off;
  lambda_g_kernel<<<(unsigned)ceil_div64(N * K * M, 256), 256, 0, st>>>(Lambda, G, N, K, M, GL);
  BNNP_LAUNCH_CHECK();
  // one fp32 accumulator per output never sums more than 4096 examples in sequence; the segments
  // are combined through H (hierarchical summation keeps the long sum at fp32-reference quality)
  for (int64_t n0 = 0; n0 < N; n0 += 4096) {
    const int64_t nn = N - n0 < 4096 ? N - n0 : 4096;
    JacOperand ja{GL + n0 * K * M, A + n0 * plan->Dtot, pu, pc, M, plan->Dtot, K};
    JacOperand jb{G + n0 * K * M, A + n0 * plan->Dtot, pu, pc, M, plan->Dtot, K};
    rc = gemm_simt(P, P, nn * K, 1, ja, JacOperandB{jb}, AccumLower{H, P}, st);
    if (rc) return rc;
  }
  return BNNP_OK;
}

extern "C" int bnnp_symmetrize_lower_rows(float* H, int64_t P, int64_t row0, int64_t row1, void* stream) {
  if (!H || P <= 0 || row0 < 0 || row1 > P || row0 > row1) return BNNP_ERR_INVALID;
  if (row0 == row1) return BNNP_OK;
  const unsigned nx = (unsigned)ceil_div64(row1, 32), ny = (unsigned)(ceil_div64(row1, 32) - row0 / 32);
  if (ny > 65535u) return BNNP_ERR_UNSUPPORTED;
  symmetrize_kernel<<<dim3(nx, ny), dim3(32, 8), 0, as_stream(stream)>>>(H, P, row0, row1);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}
extern "C" int bnnp_symmetrize_lower(float* H, int64_t P, void* stream) {
  return bnnp_symmetrize_lower_rows(H, P, 0, P, stream);
}

// ---------------------------------------------------------------------------- diag
// row splits of the conv-diag kernel: enough CTAs to fill the machine a few times over
static int conv_diag_splits(const bnnp_op_t& o, int64_t R) {
  int64_t tiles = ceil_div64((int64_t)o.in_c * o.kh * o.kw, G_BN) * ceil_div64(o.out_c, G_BM);
  int64_t nz = ceil_div64(148 * 4, tiles);
  if (nz > R) nz = R;
  if (nz > 64) nz = 64;
  return (int)(nz < 1 ? 1 : nz);
}

extern "C" int64_t bnnp_ggn_diag_scratch_floats(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                                int64_t N, int V) {
  int64_t need = N * (int64_t)plan->Mtot + 64;
  int64_t lin = 0;
  for (int i = 0; i < n_ops; ++i) {
    if (ops[i].kind == BNNP_OP_CONV2D) {
      // position sums + per-CTA partials (<= 256 CTAs) + the unfolded input of the tensor-core path
      int64_t E = (int64_t)ops[i].in_c * ops[i].kh * ops[i].kw, L = (int64_t)ops[i].out_h * ops[i].out_w;
      int64_t v = N * V * ops[i].out_c + 64 + 256LL * ((ops[i].out_c + 31) / 32 * 32) * E + 64 + N * L * ((E + 3) / 4 * 4) + 64;
      need = need > v ? need : v;
      // TMA-fed variant: bf16 planes of the gradients and of the unfolded input + per-CTA sums
      v = N * V * ops[i].out_c + 64 + conv_diag_tma_scratch_floats(N, V, ops[i].out_c, (int)L, (int)E) + 64;
      need = need > v ? need : v;
    }
    if (ops[i].kind == BNNP_OP_LINEAR) {
      int64_t mm = ops[i].out_c, nn = ops[i].in_c + 1;
      int64_t v = (1 + tcg::pick_ksplit(mm, nn, N, 1)) * mm * nn + 128;
      lin = lin > v ? lin : v;
      // TMA-fed variant: transposed bf16 hi/lo copies of Q and a^2 + the split-K partials
      const int64_t Np = (N + 7) / 8 * 8;
      v = mm * nn + 64 + (mm * Np + 63) / 64 * 64 + (nn * Np + 63) / 64 * 64 + gemm_tma_scratch_floats(mm, nn, N) + 128;
      lin = lin > v ? lin : v;
    }
  }
  return need + lin + 64;
}

// whether the first conv's weight diagonal takes the TMA-fed kernel (the only consumer of pre-split gradient planes)
static bool conv_diag_on_tma(const bnnp_op_t& o, int64_t N, int V) {
  const int L = o.out_h * o.out_w, E = o.in_c * o.kh * o.kw;
  return !g_bnnp_opt.net_simt && g_bnnp_opt.gemm_generated == 0 && (long long)N * V * o.out_c * L * E >= kTcMinMacsG &&
         conv_diag_tma_ok(o.out_c, L, N, V);
}
extern "C" int bnnp_net_first_conv_planes_ok(const bnnp_op_t* ops, int n_ops, int64_t N, int V) {
  if (!ops || N <= 0 || V <= 0 || g_bnnp_opt.conv_bwd_explicit || !net_first_conv_planes_structure_ok(ops, n_ops)) return 0;
  for (int i = 0; i < n_ops; ++i)
    if (is_param_op(ops[i])) return conv_diag_on_tma(ops[i], N, V) ? 1 : 0;
  return 0;
}
extern "C" int bnnp_ggn_diag_accum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                                   int V, const float* X, const float* ws, float factor, float* diag,
                                   float* scratch, void* stream) {
  return bnnp_ggn_diag_accum_ex(ops, n_ops, plan, N, V, X, ws, factor, diag, scratch, 0, stream);
}
extern "C" int bnnp_ggn_diag_accum_ex(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                                      int V, const float* X, const float* ws, float factor, float* diag,
                                      float* scratch, int flags, void* stream) {
  if (!ops || !plan || !ws || !diag || !scratch || N <= 0 || V <= 0) return BNNP_ERR_INVALID;
  int first_param = -1;
  for (int i = 0; i < n_ops; ++i) if (is_param_op(ops[i])) { first_param = i; break; }
  if ((flags & BNNP_BWD_FIRST_CONV_PLANES) && !bnnp_net_first_conv_planes_ok(ops, n_ops, N, V)) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  const int M = plan->Mtot;
  if (M > 0) {
    sqsum_v_kernel<<<(unsigned)ceil_div64(N * M, 256), 256, 0, st>>>(ws + plan->gcat_off, M, N, V, M, scratch);
    BNNP_LAUNCH_CHECK();
  }
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (!is_param_op(o)) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    if (o.kind == BNNP_OP_LINEAR) {
      int rc;
      if (!g_bnnp_opt.net_simt && ((long long)o.out_c * (o.in_c + 1) * N >= kTcMinMacsG || N >= 2048)) {
        // T[i, j] = sum_n Q[n, i] a[n, j]^2 on the tensor cores (split-K partials -> T), then scatter into d
        const long long mm = o.out_c, nn = o.in_c + 1;
        float* T = scratch + N * (int64_t)M + 64;
        float* wsk = T + mm * nn + 64;
        if (g_bnnp_opt.gemm_generated == 0) {
          // transposed bf16 hi/lo copies of Q [N, m] and of the squared inputs [N, n + 1] (the constant-1 column carries
          // the bias row), then one TMA-fed product over the examples
          const int64_t Np = (N + 7) / 8 * 8;
          float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(wsk) + 15) & ~uintptr_t(15));
          __nv_bfloat16* qh = reinterpret_cast<__nv_bfloat16*>(base);
          __nv_bfloat16* ql = qh + mm * Np;
          float* base2 = base + (mm * Np + 63) / 64 * 64;
          __nv_bfloat16* ah = reinterpret_cast<__nv_bfloat16*>(base2);
          __nv_bfloat16* al = ah + nn * Np;
          float* part = base2 + (nn * Np + 63) / 64 * 64;
          rc = split_transposed(scratch + q.lin_unit, M, N, mm, Np, qh, ql, false, st);
          if (rc) return rc;
          rc = split_transposed(ws + plan->acat_off + q.lin_col, plan->Dtot, N, nn, Np, ah, al, true, st);
          if (rc) return rc;
          rc = gemm_tma_nt(mm, nn, N, 1.f, qh, ql, Np, ah, al, Np, 0.f, T, nn, part, st);
        } else
        rc = tcg::gemm_tc_splitk(mm, nn, N, 1, tcg::TcColMajor{scratch + q.lin_unit, M, 0},
                                 TcColMajorSq{ws + plan->acat_off + q.lin_col, plan->Dtot}, 1.f, 0.f, T, nn, 0, wsk,
                                 tcg::pick_ksplit(mm, nn, N, 1), st);
        if (rc) return rc;
        diag_scatter_kernel<<<(unsigned)ceil_div64(mm * nn, 256), 256, 0, st>>>(T, o.out_c, o.in_c, o.has_bias,
                                                                               o.theta_off_w, o.theta_off_b, factor, diag);
        BNNP_LAUNCH_CHECK();
      } else {
        rc = gemm_simt(o.out_c, o.in_c + 1, N, 1, LoadQT{scratch, M, q.lin_unit},
                       LoadSq{ws + plan->acat_off, plan->Dtot, q.lin_col},
                       AccumLinearParam{diag, o.theta_off_w, o.theta_off_b, o.in_c, o.has_bias, factor}, st);
        if (rc) return rc;
      }
    }
  }
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_CONV2D) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    int64_t ild; const float* in = op_input(ops, plan, i, X, ws, &ild);
    if (!in) return BNNP_ERR_INVALID;
    int L = o.out_h * o.out_w, E = o.in_c * o.kh * o.kw;
    const float* G = ws + q.grad_off;
    float* partial = scratch + N * V * (int64_t)o.out_c + 64;
    if (!g_bnnp_opt.net_simt && g_bnnp_opt.gemm_generated == 0 && (long long)N * V * o.out_c * L * E >= kTcMinMacsG &&
        conv_diag_tma_ok(o.out_c, L, N, V)) {
      // flags & BNNP_BWD_FIRST_CONV_PLANES: the pool backward left this layer's gradient as the kernel's bf16 planes
      const bool planes = (flags & BNNP_BWD_FIRST_CONV_PLANES) && i == first_param;
      const __nv_bfloat16* ph = planes ? reinterpret_cast<const __nv_bfloat16*>(G) : nullptr;
      int rc = conv_diag_tma(make_im2col(o, in, ild, 1), G, q.grad_ld, N, V, o.out_c, L, E, factor, diag + o.theta_off_w,
                             partial, st, ph, planes ? ph + N * V * q.grad_ld : nullptr);
      if (rc) return rc;
    } else if (!g_bnnp_opt.net_simt && (long long)N * V * o.out_c * L * E >= kTcMinMacsG && N * V <= 0x7fffffff) {
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const int Epad = (E + 3) / 4 * 4;
      const int MP = (o.out_c + 31) / 32 * 32;
      const long long stride = (long long)MP * E;
      float* U = partial + sms * stride + 64;
      BNNP_CUDA_CHECK(cudaMemsetAsync(partial, 0, sizeof(float) * sms * stride, st));
      im2col_kernel<<<(unsigned)ceil_div64(N * L * (int64_t)Epad, 256), 256, 0, st>>>(make_im2col(o, in, ild, 1), N, L, E,
                                                                                     Epad, U);
      BNNP_LAUNCH_CHECK();
      int rc = tcg::gemm_tc(o.out_c, E, L, (int)(N * V), tcg::TcRowMajor{G, L, q.grad_ld}, TcUnfoldT{U, Epad, L, V},
                            TcSqAccumPart(partial, stride, MP), 1, st);
      if (rc) return rc;
      reduce_partials_t_kernel<<<(unsigned)ceil_div64((int64_t)o.out_c * E, 256), 256, 0, st>>>(
          partial, sms, stride, o.out_c, E, MP, factor, diag + o.theta_off_w);
      BNNP_LAUNCH_CHECK();
    } else {
      const int nz = conv_diag_splits(o, N * V);
      dim3 grid((unsigned)ceil_div64(E, G_BN), (unsigned)ceil_div64(o.out_c, G_BM), (unsigned)nz);
      conv_diag_kernel<<<grid, G_THREADS, 0, st>>>(ConvGradA{G, q.grad_ld, L},
                                                   ConvUnfoldB{make_im2col(o, in, ild, V)}, N * V, o.out_c, E,
                                                   L, partial);
      BNNP_LAUNCH_CHECK();
      reduce_partials_kernel<<<(unsigned)ceil_div64((int64_t)o.out_c * E, 256), 256, 0, st>>>(
          partial, nz, (int64_t)o.out_c * E, factor, diag + o.theta_off_w);
      BNNP_LAUNCH_CHECK();
    }
    if (o.has_bias) {
      int64_t R = N * V;
      if ((flags & BNNP_BWD_FIRST_CONV_PLANES) && i == first_param) {      // no fp32 gradient: position sums from the pooled one
        int rc = net_first_conv_possum(ops, n_ops, plan, N, V, ws, scratch, st);
        if (rc) return rc;
      } else {
        launch_conv_possum(G, q.grad_ld, R, o.out_c, L, scratch, o.out_c, st);
      }
      BNNP_LAUNCH_CHECK();
      colsq_accum_kernel<<<(unsigned)ceil_div64(o.out_c, 32), dim3(32, 32), 0, st>>>(scratch, R, o.out_c, factor,
                                                                                    diag + o.theta_off_b);
      BNNP_LAUNCH_CHECK();
    }
  }
  return BNNP_OK;
}

// ---------------------------------------------------------------------------- kfac
// scratch layout of the KFAC pass: [conv position-sums][split-K partials of the tensor-core Grams]
static int64_t kfac_conv_floats(const bnnp_op_t* ops, int n_ops, int64_t N, int V) {
  int64_t need = 64;
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind == BNNP_OP_CONV2D) need = need > N * V * ops[i].out_c + 64 ? need : N * V * ops[i].out_c + 64;
  return need;
}
extern "C" int64_t bnnp_kfac_scratch_floats(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                            int64_t N, int V) {
  (void)plan;
  int64_t part = 0;
  for (int i = 0; i < n_ops; ++i) {
    if (ops[i].kind == BNNP_OP_LINEAR) {
      int64_t m = ops[i].out_c, n = ops[i].in_c;
      int64_t a = (int64_t)tcg::pick_ksplit(m, m, N * V, 1) * m * m, b = (int64_t)tcg::pick_ksplit(n, n, N, 1) * n * n;
      part = part > a ? part : a;
      part = part > b ? part : b;
      a = gram_tma_scratch_floats(N * V, m);
      b = gram_tma_scratch_floats(N, n);
      part = part > a ? part : a;
      part = part > b ? part : b;
    }
    if (ops[i].kind == BNNP_OP_CONV2D) {
      int64_t E = (int64_t)ops[i].in_c * ops[i].kh * ops[i].kw, L = (int64_t)ops[i].out_h * ops[i].out_w;
      int64_t v = N * L * ((E + 3) / 4 * 4) + 64 + (int64_t)tcg::pick_ksplit(E, E, N * L, 1) * E * E;
      part = part > v ? part : v;
      v = (E * ((N * L + 7) / 8 * 8) + 63) / 64 * 64 + gemm_tma_scratch_floats(E, E, N * L) + 128;     // TMA-fed A-factor Gram
      part = part > v ? part : v;
      int64_t m = ops[i].out_c;
      v = gram_tma_scratch_floats(N * V, m) + 64;
      part = part > v ? part : v;
      int64_t gsplit = (int64_t)(tcg::pick_ksplit(m, m, N * V, 1) + (N * V + 4095) / 4096 + 1) * m * m;
      part = part > gsplit ? part : gsplit;
    }
  }
  return kfac_conv_floats(ops, n_ops, N, V) + part + 128;
}

extern "C" int bnnp_kfac_accum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, int V,
                               const float* X, const float* ws, float factor, float batch_factor,
                               float* const* factors, float* scratch, void* stream) {
  return bnnp_kfac_accum_ex(ops, n_ops, plan, N, V, X, ws, factor, batch_factor, factors, scratch, 0, stream);
}

// flags & BNNP_BWD_FIRST_CONV_POSSUM: the backward ran with the same flag, i.e. the gradient buffer of the first
// parameterised (conv) layer holds its position sums [N*V, Cout] instead of the full [N*V, Cout, L] gradient
extern "C" int bnnp_kfac_accum_ex(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, int V,
                                  const float* X, const float* ws, float factor, float batch_factor,
                                  float* const* factors, float* scratch, int flags, void* stream) {
  if (!ops || !plan || !ws || !factors || !scratch || N <= 0 || V <= 0) return BNNP_ERR_INVALID;
  int first_param = -1;
  for (int i = 0; i < n_ops; ++i) if (is_param_op(ops[i])) { first_param = i; break; }
  cudaStream_t st = as_stream(stream);
  const int64_t R = N * V;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (!is_param_op(o)) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    float* Gf = factors[2 * i];
    float* Af = factors[2 * i + 1];
    if (!Gf || !Af) return BNNP_ERR_INVALID;
    const float* G = ws + q.grad_off;
    if (o.kind == BNNP_OP_LINEAR) {
      int rc;
      float* wsk = scratch + kfac_conv_floats(ops, n_ops, N, V);
      // (a narrow Gram over many rows -- the 10 x 10 factor of the output layer over N*K rows -- is ONE SIMT block
      // walking all rows: 3.5 ms of a 6.2 ms pass; split-K on the tensor cores spreads it over the machine)
      const bool tma = g_bnnp_opt.gemm_generated == 0;
      if (!g_bnnp_opt.net_simt && tma && ((long long)o.out_c * o.out_c * R >= kTcMinMacsG || R >= 2048))
        // G += factor * g^T g: transposed bf16 hi/lo split of the gradients once, then a TMA-fed tensor-core Gram
        rc = gram_tma(G, q.grad_ld, R, o.out_c, factor, 1.f, Gf, o.out_c, wsk, st);
      else if (!g_bnnp_opt.net_simt && ((long long)o.out_c * o.out_c * R >= kTcMinMacsG || R >= 2048))
        rc = tcg::gemm_tc_splitk(o.out_c, o.out_c, R, 1, tcg::TcColMajor{G, q.grad_ld, 0}, tcg::TcColMajor{G, q.grad_ld, 0},
                                 factor, 1.f, Gf, o.out_c, 0, wsk, tcg::pick_ksplit(o.out_c, o.out_c, R, 1), st);
      else
        rc = gemm_simt(o.out_c, o.out_c, R, 1, LoadTA{G, q.grad_ld, 0}, LoadNB{G, q.grad_ld, 0},
                       AccumScaled{Gf, o.out_c, factor}, st);
      if (rc) return rc;
      const float* A = ws + plan->acat_off;
      if (!g_bnnp_opt.net_simt && tma && ((long long)o.in_c * o.in_c * N >= kTcMinMacsG || N >= 2048))
        rc = gram_tma(A + q.lin_col, plan->Dtot, N, o.in_c, batch_factor / (float)N, 1.f, Af, o.in_c, wsk, st);
      else if (!g_bnnp_opt.net_simt && ((long long)o.in_c * o.in_c * N >= kTcMinMacsG || N >= 2048))
        rc = tcg::gemm_tc_splitk(o.in_c, o.in_c, N, 1, tcg::TcColMajor{A + q.lin_col, plan->Dtot, 0},
                                 tcg::TcColMajor{A + q.lin_col, plan->Dtot, 0}, batch_factor / (float)N, 1.f, Af, o.in_c,
                                 0, wsk, tcg::pick_ksplit(o.in_c, o.in_c, N, 1), st);
      else
        rc = gemm_simt(o.in_c, o.in_c, N, 1, LoadTA{A, plan->Dtot, q.lin_col}, LoadNB{A, plan->Dtot, q.lin_col},
                       AccumScaled{Af, o.in_c, batch_factor / (float)N}, st);
      if (rc) return rc;
    } else {
      int64_t ild; const float* in = op_input(ops, plan, i, X, ws, &ild);
      if (!in) return BNNP_ERR_INVALID;
      int L = o.out_h * o.out_w, E = o.in_c * o.kh * o.kw;
      if ((flags & BNNP_BWD_FIRST_CONV_POSSUM) && i == first_param) {
        BNNP_CUDA_CHECK(cudaMemcpyAsync(scratch, G, sizeof(float) * R * o.out_c, cudaMemcpyDeviceToDevice, st));
      } else {
        launch_conv_possum(G, q.grad_ld, R, o.out_c, L, scratch, o.out_c, st);
        BNNP_LAUNCH_CHECK();
      }
      int rc;
      // G += factor * S^T S over the R position-summed rows: a 64..128-wide Gram is one to four SIMT blocks walking
      // all R rows -- the split-K tensor-core Gram spreads it over the machine
      const bool tma = g_bnnp_opt.gemm_generated == 0;
      if (!g_bnnp_opt.net_simt && tma && (long long)o.out_c * o.out_c * R >= (1LL << 22))
        rc = gram_tma(scratch, o.out_c, R, o.out_c, factor, 1.f, Gf, o.out_c, scratch + kfac_conv_floats(ops, n_ops, N, V), st);
      else if (!g_bnnp_opt.net_simt && (long long)o.out_c * o.out_c * R >= (1LL << 22))
        rc = tcg::gemm_tc_splitk(o.out_c, o.out_c, R, 1, tcg::TcColMajor{scratch, o.out_c, 0}, tcg::TcColMajor{scratch, o.out_c, 0},
                                 factor, 1.f, Gf, o.out_c, 0, scratch + kfac_conv_floats(ops, n_ops, N, V),
                                 tcg::pick_ksplit(o.out_c, o.out_c, R, 1), st);
      else
        rc = gemm_simt(o.out_c, o.out_c, R, 1, LoadTA{scratch, o.out_c, 0}, LoadNB{scratch, o.out_c, 0},
                       AccumScaled{Gf, o.out_c, factor}, st);
      if (rc) return rc;
      Im2col u = make_im2col(o, in, ild, 1);
      if (!g_bnnp_opt.net_simt && tma && (long long)E * E * N * L >= kTcMinMacsG) {
        // the unfolded input written once, transposed and pre-split (bf16 hi/lo), then A += (bf/N) U^T U as a TMA-fed Gram
        const int64_t NL = N * L, NLp = (NL + 7) / 8 * 8;
        float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch + kfac_conv_floats(ops, n_ops, N, V)) + 15) & ~uintptr_t(15));
        __nv_bfloat16* uh = reinterpret_cast<__nv_bfloat16*>(base);
        __nv_bfloat16* ul = uh + (int64_t)E * NLp;
        float* part = base + ((int64_t)E * NLp + 63) / 64 * 64;
        if (E > 65535) return BNNP_ERR_UNSUPPORTED;
        im2col_t_split_kernel<<<dim3((unsigned)ceil_div64(NLp >> 3, 256), (unsigned)E), 256, 0, st>>>(u, N, L, E, NLp, uh, ul);
        BNNP_LAUNCH_CHECK();
        rc = gemm_tma_nt(E, E, NL, batch_factor / (float)N, uh, ul, NLp, uh, ul, NLp, 1.f, Af, E, part, st);
      } else if (!g_bnnp_opt.net_simt && (long long)E * E * N * L >= kTcMinMacsG) {
        // materialise the unfolded input once, then A += (bf/N) U^T U as a tensor-core Gram (split-K)
        const int Epad = (E + 3) / 4 * 4;
        float* U = scratch + kfac_conv_floats(ops, n_ops, N, V);
        float* wsk = U + N * L * (int64_t)Epad + 64;
        im2col_kernel<<<(unsigned)ceil_div64(N * L * (int64_t)Epad, 256), 256, 0, st>>>(u, N, L, E, Epad, U);
        BNNP_LAUNCH_CHECK();
        rc = tcg::gemm_tc_splitk(E, E, N * L, 1, tcg::TcColMajor{U, Epad, 0}, tcg::TcColMajor{U, Epad, 0},
                                 batch_factor / (float)N, 1.f, Af, E, 0, wsk, tcg::pick_ksplit(E, E, N * L, 1), st);
      } else {
        rc = gemm_simt(E, E, N * L, 1, UnfoldTA{u, L}, UnfoldNB{u, L},
                       AccumScaled{Af, E, batch_factor / (float)N}, st);
      }
      if (rc) return rc;
    }
  }
  return BNNP_OK;
}
