// Generic tensor-core GEMM with functor (generated) operands -- the tcgen05 counterpart of
// gemm_simt.cuh:   acc(m, n) = sum_k A(m, k) * B(k, n),   fp32-grade via split-bf16 (3 MMAs).
//
// Both operands are *synthesised* by 8 generator warps from functors (strided matrices, squared
// activations, implicit transposes ...): each value is computed in fp32, split into bf16 hi/lo and
// written to shared memory in the UMMA canonical K-major SWIZZLE_64B layout.  One thread issues
// tcgen05.mma (M = 128, N = tile width, K = 16, kind::f16) into one of two TMEM accumulator buffers;
// four dedicated epilogue warps drain the other buffer (tcgen05.ld) concurrently.  Persistent CTAs
// walk (batch, m-tile, n-tile, k-split) tiles round-robin.
//
// Operand functors:
//   struct LA { static constexpr bool K_CONTIG;   // memory order: true = k fastest, false = m fastest
//               using Row = ...;                   // per-row handle, computed once per tile
//               __device__ Row row(int64_t m, int batch) const;
//               __device__ void load8(const Row&, int64_t k0, int64_t K, float* out) const; };
//   (load8 returns A(m, k0..k0+7), zero beyond K; rows m >= M are never requested)
//   struct LB { ... same with n in place of m:  B(k, n) = lb.load8(lb.row(n, b), k0, ...) ... };
// Epilogue functor:
//   struct EP { struct State {...};                       // per-thread scratch carried across one tile row
//               __device__ void begin(State&) const;
//               __device__ void row(State&, int64_t m, int64_t n0, const float* v, int nvalid, int batch, int ks) const;
//               __device__ void end(State&, int64_t m, int batch, int ks) const; };
//   (v[0..nvalid) = acc(m, n0 ..), called by the thread that owns row m, 32 columns at a time)
#pragma once
#include "tc_common.cuh"

namespace bnnp {
namespace tcg {

using namespace bnnp::tcp;

constexpr int BK = 32, TILE_M = 128, MAX_N = 256, STAGES = 4;
constexpr int A_TERM = TILE_M * BK * 2, A_STAGE = 2 * A_TERM;      // hi + lo
constexpr int B_TERM = MAX_N * BK * 2, B_STAGE = 2 * B_TERM;
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;                     // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int NUM_THREADS = 512;      // warp 1: MMA, warp 2: TMEM alloc, warps 4-11: generators, 12-15: epilogue
constexpr int GEN_WARP0 = 4, EPI_WARP0 = 12;
constexpr int kMaxChunksPerAccum = 128;   // 4096 K-elements per TMEM accumulation (truncation bias ~2e-5)

struct Shape {
  long long M, N, K;
  int batch, wn, n_mt, n_nt, ksplit, chunks_total, chunk0;   // this launch covers K-chunks [chunk0, chunk0 + chunks_total)
};

template <class LA, class LB, class EP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(Shape sh, LA la, LB lb, EP ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES] generators -> MMA   (8 warp arrivals)
  uint64_t* empty = bars + STAGES;             // [STAGES] MMA commit -> generators
  uint64_t* acc_full = bars + 2 * STAGES;      // [2] MMA commit -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2; // [2] epilogue (4 warp arrivals) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 8); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_tiles = (long long)sh.batch * sh.n_mt * sh.n_nt * sh.ksplit;
  uint32_t stage = 0, phase = 0;
  uint32_t iter = 0;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
    // m-tile fastest: CTAs that run concurrently share the same B (n-tile) operand through L2
    const int ks = (int)(tile % sh.ksplit);
    const int mt = (int)((tile / sh.ksplit) % sh.n_mt);
    const int nt = (int)((tile / ((long long)sh.ksplit * sh.n_mt)) % sh.n_nt);
    const int b = (int)(tile / ((long long)sh.ksplit * sh.n_nt * sh.n_mt));
    const int c_beg = sh.chunk0 + (int)((long long)sh.chunks_total * ks / sh.ksplit);
    const int c_end = sh.chunk0 + (int)((long long)sh.chunks_total * (ks + 1) / sh.ksplit);
    const int n_chunks = c_end - c_beg;
    const long long m0 = (long long)mt * TILE_M, n0 = (long long)nt * sh.wn;
    const uint32_t buf = iter & 1, buf_phase = (iter >> 1) & 1;

    if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        mbar_wait(&acc_empty[buf], buf_phase ^ 1);        // epilogue has drained this accumulator
        tc_fence_after();
        uint32_t s = stage, ph = phase;
        const uint32_t idesc = umma_idesc_bf16(sh.wn);
        const uint32_t d = tmem_base + buf * MAX_N;
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + A_STAGE;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t dah = umma_desc_sw64(a0 + kk * 32), dal = umma_desc_sw64(a0 + A_TERM + kk * 32);
            const uint64_t dbh = umma_desc_sw64(b0 + kk * 32), dbl = umma_desc_sw64(b0 + B_TERM + kk * 32);
            umma_bf16(d, dah, dbh, idesc, (c | kk) ? 1u : 0u);
            umma_bf16(d, dah, dbl, idesc, 1u);
            umma_bf16(d, dal, dbh, idesc, 1u);
          }
          umma_commit(&empty[s]);
          if (c == n_chunks - 1) umma_commit(&acc_full[buf]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp >= GEN_WARP0 && warp < EPI_WARP0) {
      // ===================== operand generators =====================
      // 8 warps; per K-chunk each warp fills 2 warp-items of the 128 x 32 A tile (items gw, gw+8) and up
      // to 4 of the wn x 32 B tile.  A warp-item is 8 rows x 4 K-groups (k-contiguous operand: 128 B per
      // row) or 32 rows x 1 K-group (row-contiguous operand); every lane owns one 16-byte K-group.
      const int gw = warp - GEN_WARP0;
      const int n_items_b = LB::K_CONTIG ? sh.wn / 8 : ((sh.wn + 31) / 32) * 4;
      typename LA::Row ha[2];
      typename LB::Row hb[4];
      uint32_t offa[2], offb[4];
      int ga[2], gb[4];
      bool oka[2], okb[4], useb[4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int wi = gw + j * 8;
        int row;
        if (LA::K_CONTIG) { row = wi * 8 + (lane >> 2); ga[j] = lane & 3; }
        else { row = (wi >> 2) * 32 + lane; ga[j] = wi & 3; }
        oka[j] = m0 + row < sh.M;
        ha[j] = la.row(oka[j] ? m0 + row : 0, b);
        offa[j] = sw64_offset(row, ga[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int wi = gw + j * 8;
        useb[j] = wi < n_items_b;
        int row;
        if (LB::K_CONTIG) { row = wi * 8 + (lane >> 2); gb[j] = lane & 3; }
        else { row = (wi >> 2) * 32 + lane; gb[j] = wi & 3; }
        okb[j] = useb[j] && n0 + row < sh.N;
        hb[j] = lb.row(okb[j] ? n0 + row : 0, b);
        offb[j] = (uint32_t)A_STAGE + sw64_offset(row, gb[j]);
      }
      uint32_t s = stage, ph = phase;
      for (int c = 0; c < n_chunks; ++c) {
        const long long k0 = (long long)(c_beg + c) * BK;
        float va[2][8], vb[4][8];
        // 1) issue all loads of this chunk (independent of the smem stage)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (oka[j]) la.load8(ha[j], k0 + ga[j] * 8, sh.K, va[j]);
          else {
#pragma unroll
            for (int e = 0; e < 8; ++e) va[j][e] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (okb[j]) lb.load8(hb[j], k0 + gb[j] * 8, sh.K, vb[j]);
          else {
#pragma unroll
            for (int e = 0; e < 8; ++e) vb[j][e] = 0.f;
          }
        }
        // 2) wait for the stage, split to bf16 hi/lo and publish
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* base = smem + s * STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 hi, lo;
          split8(va[j], hi, lo);
          *reinterpret_cast<uint4*>(base + offa[j]) = hi;
          *reinterpret_cast<uint4*>(base + offa[j] + A_TERM) = lo;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (useb[j]) {
            uint4 hi, lo;
            split8(vb[j], hi, lo);
            *reinterpret_cast<uint4*>(base + offb[j]) = hi;
            *reinterpret_cast<uint4*>(base + offb[j] + B_TERM) = lo;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    } else if (warp >= EPI_WARP0) {
      // ===================== epilogue =====================
      const int q = warp - EPI_WARP0;
      mbar_wait(&acc_full[buf], buf_phase);
      tc_fence_after();
      const long long m = m0 + q * 32 + lane;
      const long long ncols = sh.N - n0 < sh.wn ? sh.N - n0 : sh.wn;
      typename EP::State est;
      ep.begin(est);
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * MAX_N + c0), v);
        if (m < sh.M) {
          const int nvalid = ncols - c0 < 32 ? (int)(ncols - c0) : 32;
          ep.row(est, m, n0 + c0, reinterpret_cast<const float*>(v), nvalid, b, ks);
        }
      }
      if (m < sh.M) ep.end(est, m, b, ks);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    // every role advanced the smem ring by n_chunks stages
    const uint32_t tot = stage + (uint32_t)n_chunks;
    phase ^= (tot / STAGES) & 1;
    stage = tot % STAGES;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// pick the n-tile width: fewest tiles, then least padding (multiple of 16, <= 256)
inline int pick_wn(long long N) {
  long long nt = (N + MAX_N - 1) / MAX_N;
  long long w = ((N + nt - 1) / nt + 15) / 16 * 16;
  return (int)(w < 16 ? 16 : w);
}

// chunk_lo / chunk_hi restrict the launch to K-chunks [chunk_lo, chunk_hi) (chunk_hi < 0: to the end) -- used to
// run a long K as sequential segments that accumulate through the output (see gemm_tc_ksegments)
template <class LA, class LB, class EP>
inline int gemm_tc(long long M, long long N, long long K, int batch, LA la, LB lb, EP ep, int ksplit,
                   cudaStream_t st, int chunk_lo = 0, int chunk_hi = -1) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return BNNP_OK;
  Shape sh;
  sh.M = M; sh.N = N; sh.K = K; sh.batch = batch;
  sh.wn = pick_wn(N);
  sh.n_mt = (int)((M + TILE_M - 1) / TILE_M);
  sh.n_nt = (int)((N + sh.wn - 1) / sh.wn);
  {
    const int all = (int)((K + BK - 1) / BK);
    if (chunk_hi < 0 || chunk_hi > all) chunk_hi = all;
    if (chunk_lo >= chunk_hi) return BNNP_OK;
    sh.chunk0 = chunk_lo;
    sh.chunks_total = chunk_hi - chunk_lo;
  }
  sh.ksplit = ksplit < 1 ? 1 : (ksplit > sh.chunks_total ? sh.chunks_total : ksplit);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  auto kern = gemm_tc_kernel<LA, LB, EP>;
  BNNP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const long long n_tiles = (long long)batch * sh.n_mt * sh.n_nt * sh.ksplit;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  kern<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(sh, la, lb, ep);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

// A long K without a split-K workspace: sequential launches over <= kMaxChunksPerAccum-chunk segments; make_ep(seg)
// returns the epilogue of segment `seg` (segment 0 stores, later segments accumulate into the output).
template <class LA, class LB, class MakeEP>
inline int gemm_tc_ksegments(long long M, long long N, long long K, int batch, LA la, LB lb, MakeEP make_ep,
                             cudaStream_t st) {
  const int all = (int)((K + BK - 1) / BK);
  int seg = 0;
  for (int c0 = 0; c0 < all; c0 += kMaxChunksPerAccum, ++seg) {
    int rc = gemm_tc(M, N, K, batch, la, lb, make_ep(seg), 1, st, c0, c0 + kMaxChunksPerAccum);
    if (rc) return rc;
  }
  return BNNP_OK;
}

// split-K factor that fills the machine for an M x N output with `chunks` K-chunks
inline int pick_ksplit(long long M, long long N, long long K, int batch) {
  long long tiles = (long long)batch * ((M + TILE_M - 1) / TILE_M) * ((N + pick_wn(N) - 1) / pick_wn(N));
  long long chunks = (K + BK - 1) / BK;
  long long want = (148 * 2 + tiles - 1) / tiles;
  long long maxs = chunks / 16 > 0 ? chunks / 16 : 1;       // at least 16 chunks per split
  long long s = want < maxs ? want : maxs;
  long long need = (chunks + kMaxChunksPerAccum - 1) / kMaxChunksPerAccum;     // accuracy floor
  if (s < need) s = need;
  return (int)(s < 1 ? 1 : (s > 512 ? 512 : s));
}

// ---- stock operand functors -------------------------------------------------------------------
struct TcRowMajor {        // X(r, k) = p[b*bs + r*ld + k]            (k contiguous)
  static constexpr bool K_CONTIG = true;
  using Row = const float*;
  const float* p; long long ld, bs;
  __device__ __forceinline__ Row row(int64_t r, int b) const { return p + (long long)b * bs + r * ld; }
  __device__ __forceinline__ void load8(const Row& rw, int64_t k0, int64_t K, float* out) const {
    const float* q = rw + k0;
    if (k0 + 8 <= K) {
      if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(q)), y = __ldg(reinterpret_cast<const float4*>(q) + 1);
        out[0] = x.x; out[1] = x.y; out[2] = x.z; out[3] = x.w; out[4] = y.x; out[5] = y.y; out[6] = y.z; out[7] = y.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] = __ldg(q + e);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = (k0 + e < K) ? __ldg(q + e) : 0.f;
    }
  }
};
struct TcColMajor {        // X(r, k) = p[b*bs + k*ld + r]            (r contiguous: transposed storage)
  static constexpr bool K_CONTIG = false;
  using Row = const float*;
  const float* p; long long ld, bs;
  __device__ __forceinline__ Row row(int64_t r, int b) const { return p + (long long)b * bs + r; }
  __device__ __forceinline__ void load8(const Row& rw, int64_t k0, int64_t K, float* out) const {
    const float* q = rw + k0 * ld;
    if (k0 + 8 <= K) {
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = __ldg(q + (long long)e * ld);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = (k0 + e < K) ? __ldg(q + (long long)e * ld) : 0.f;
    }
  }
};

// ---- stock epilogues ----------------------------------------------------------------------------
struct TcNoState {
  struct State {};
  __device__ __forceinline__ void begin(State&) const {}
  __device__ __forceinline__ void end(State&, int64_t, int, int) const {}
};
struct TcStoreAxpby : TcNoState {      // C = alpha*acc + beta*C
  float* p; long long ld, bs; float alpha, beta;
  __host__ __device__ TcStoreAxpby(float* p_, long long ld_, long long bs_, float a_, float b_) : p(p_), ld(ld_), bs(bs_), alpha(a_), beta(b_) {}
  __device__ __forceinline__ void row(State&, int64_t m, int64_t n0, const float* v, int nvalid, int b, int) const {
    float* q = p + (long long)b * bs + m * ld + n0;
    // fully unrolled with predicates so that v[] stays in registers
    if (beta == 0.f) {
#pragma unroll
      for (int e = 0; e < 32; ++e) if (e < nvalid) q[e] = alpha * v[e];
    } else {
      float old[32];                     // loads before stores: interleaved on the same array they serialise
#pragma unroll
      for (int e = 0; e < 32; ++e) old[e] = e < nvalid ? q[e] : 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e) if (e < nvalid) q[e] = alpha * v[e] + beta * old[e];
    }
  }
};
struct TcStorePartial : TcNoState {    // ws[ks][b][m][n] = acc     (deterministic split-K: reduced afterwards)
  float* ws; long long M, N; int batch;
  __host__ __device__ TcStorePartial(float* w_, long long M_, long long N_, int b_) : ws(w_), M(M_), N(N_), batch(b_) {}
  __device__ __forceinline__ void row(State&, int64_t m, int64_t n0, const float* v, int nvalid, int b, int ks) const {
    float* q = ws + (((long long)ks * batch + b) * M + m) * N + n0;
#pragma unroll
    for (int e = 0; e < 32; ++e) if (e < nvalid) q[e] = v[e];
  }
};

// C[b][m][n] = alpha * sum_ks ws[ks][b][m][n] + beta * C[b][m][n]   (fixed order)
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int ksplit, long long M, long long N,
                                     int batch, float alpha, float beta, float* __restrict__ C, long long ldc,
                                     long long bsc);

template <class LA, class LB>
inline int gemm_tc_splitk(long long M, long long N, long long K, int batch, LA la, LB lb, float alpha,
                          float beta, float* C, long long ldc, long long bsc, float* ws, int ksplit,
                          cudaStream_t st) {
  long long chunks = (K + BK - 1) / BK;
  // the TMEM accumulator truncates on every update: with scratch available never let one
  // accumulation span more than kMaxChunksPerAccum K-chunks (partials are summed in fp32 RN)
  if (ws && (chunks + kMaxChunksPerAccum - 1) / kMaxChunksPerAccum > ksplit)
    ksplit = (int)((chunks + kMaxChunksPerAccum - 1) / kMaxChunksPerAccum);
  if (ksplit <= 1 || !ws) return gemm_tc(M, N, K, batch, la, lb, TcStoreAxpby{C, ldc, bsc, alpha, beta}, 1, st);
  if (ksplit > chunks) ksplit = (int)chunks;
  int rc = gemm_tc(M, N, K, batch, la, lb, TcStorePartial{ws, M, N, batch}, ksplit, st);
  if (rc) return rc;
  long long tot = (long long)batch * M * N;
  splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ws, ksplit, M, N, batch, alpha, beta, C, ldc, bsc);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

}  // namespace tcg
}  // namespace bnnp
