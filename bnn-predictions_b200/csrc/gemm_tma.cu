// TMA-fed tensor-core GEMM for operands that are PLAIN matrices (the layer products of the factor paths: forward,
// K-column backward, KFAC Grams A = a^T a / G = g^T g of preds/kron.py:58-69 and preds/laplace.py:69-78):
//
//     C[M,N] = alpha * sum_k A[m,k] * B[n,k] + beta * C          ("NT": both operands K-major, k contiguous)
//
// gemm_tc.cuh synthesises both operands with generator warps (global load -> fp32 -> bf16 hi/lo split -> swizzled
// shared-memory stores) on every tile: measured 17 % tensor-pipe activity on the KFAC Grams, generator bound.  Here the
// operands are split ONCE into bf16 hi/lo copies with 16-byte aligned rows (split_rows / split_transposed below: one
// HBM pass over a matrix that the GEMM then re-reads many times), and `cp.async.bulk.tensor` (SWIZZLE_64B boxes)
// streams them straight into the UMMA layout: no thread touches an operand.  fp32-grade arithmetic as everywhere:
// D += Ah*Bh + Ah*Bl + Al*Bh, three `tcgen05.mma.kind::f16` per 16-wide K-step, fp32 accumulators in TMEM
// (double-buffered: the epilogue of a tile runs under the mainloop of the next one).  One accumulation never spans more
// than 4096 K-elements (TMEM adds truncate): longer K is split into deterministic partials that a second
// kernel sums in fp32 round-to-nearest.
#include <stdlib.h>
#include "tc_common.cuh"
#include "gemm_tma.cuh"

namespace bnnp {
namespace tg {

using namespace bnnp::tcp;

constexpr int BK = 32, TILE_M = 128, MAX_N = 256, STAGES = 4;
constexpr int A_TERM = TILE_M * BK * 2, A_STAGE = 2 * A_TERM;      // hi + lo: 16 KB
constexpr int B_TERM = MAX_N * BK * 2, B_STAGE = 2 * B_TERM;       // 32 KB
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;                     // 48 KB
constexpr int EPI_TILE_BYTES = 32 * 33 * 4 + 32;                   // per epilogue warp: [32][33] fp32 transpose tile
constexpr int EPI_WARPS = 8;          // two per TMEM lane quarter, each draining half of the tile's columns: with K of a few
                                      // hundred (the K-column backward products) a tile's epilogue outlasts its MMAs
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_WARPS * EPI_TILE_BYTES;
constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;   // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-11: epilogue
constexpr int EPI_WARP0 = 4;
constexpr int kMaxChunksPerAccum = 128;

struct Args {
  long long M, N;
  int wn, n_mt, n_nt, ksplit, chunks_total;
  float alpha, beta;
  float* C; long long ldc;
  float* part;                 // [ksplit, M, N] partial sums when ksplit > 1
  const float* bias;           // [N] added to every row (ksplit == 1 only) or nullptr
  int ct_L;                    // > 0: rows are (batch b, position p) = (m / ct_L, m % ct_L) and C is stored transposed per batch:
                               //      C[(b * N + n) * ct_L + p]  (the [R, E, L] buffer of the conv backward-data product)
  const float* had;            // != nullptr: the product is multiplied elementwise by had[m, n] (row pitch ldh) before it enters C
  long long ldh;               //   (function-space kernels: K_c += (G G^T) o (A A^T), preds/gps.py:56-88 per layer)
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                        // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;              // [STAGES] MMA commit -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;       // [2] MMA commit -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2] epilogue (4 warp arrivals) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_tiles = (long long)args.n_mt * args.n_nt * args.ksplit;
  uint32_t stage = 0, phase = 0, iter = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
    // m-tile fastest: CTAs that run concurrently share the same B tile (and K range) through L2
    const int mt = (int)(tile % args.n_mt);
    const int nt = (int)((tile / args.n_mt) % args.n_nt);
    const int ks = (int)(tile / ((long long)args.n_mt * args.n_nt));
    const int c_beg = (int)((long long)args.chunks_total * ks / args.ksplit);
    const int c_end = (int)((long long)args.chunks_total * (ks + 1) / args.ksplit);
    const int n_chunks = c_end - c_beg;
    const long long m0 = (long long)mt * TILE_M, n0 = (long long)nt * args.wn;
    const uint32_t buf = iter & 1, buf_phase = (iter >> 1) & 1;

    if (warp == 0) {
      if (lane == 0) {
        uint32_t s = stage, ph = phase;
        const uint32_t bytes = (uint32_t)(A_STAGE + 2 * args.wn * BK * 2);
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* base = smem + s * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], bytes);
          const int k0 = (c_beg + c) * BK;
          tma_load_2d(&map_ah, &full[s], base, k0, (int)m0);
          tma_load_2d(&map_al, &full[s], base + A_TERM, k0, (int)m0);
          tma_load_2d(&map_bh, &full[s], base + A_STAGE, k0, (int)n0);
          tma_load_2d(&map_bl, &full[s], base + A_STAGE + B_TERM, k0, (int)n0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        mbar_wait(&acc_empty[buf], buf_phase ^ 1);
        tc_fence_after();
        uint32_t s = stage, ph = phase;
        const uint32_t idesc = umma_idesc_bf16(args.wn);
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
    } else if (warp >= EPI_WARP0) {
      // epilogue: tcgen05.ld hands every thread one accumulator row (32 columns per load); a per-warp shared-memory
      // transpose turns that into lane = column, so that every store instruction of the warp writes 128 contiguous
      // bytes of one output row (thread-per-row stores touch 32 cache lines per instruction: measured 41 us per
      // 128 x 256 tile against 6.6 us of MMA time on the K = 512 backward product)
      const int q = (warp - EPI_WARP0) & 3;                   // TMEM lane quarter (= warp index modulo 4)
      const int half = (warp - EPI_WARP0) >> 2;               // which half of the tile's 32-column blocks
      float* sc = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256 + (warp - EPI_WARP0) * EPI_TILE_BYTES);
      mbar_wait(&acc_full[buf], buf_phase);
      tc_fence_after();
      if (args.ct_L > 0) {
        // transposed-per-batch store: the thread's row m = (b, p) is the contiguous index of the output, so the
        // tcgen05.ld layout (thread = row) is already the coalesced one -- no transpose
        const long long m = m0 + q * 32 + lane;
        const int ncols_t = (int)(args.N - n0 < args.wn ? args.N - n0 : args.wn);
        float* pt = args.C + ((m / args.ct_L) * args.N + n0) * args.ct_L + (m % args.ct_L);
        for (int c0 = half * 32; c0 < ncols_t; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * MAX_N + c0), v);
          if (m < args.M) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c0 + e < ncols_t) pt[(long long)(c0 + e) * args.ct_L] = args.alpha * __uint_as_float(v[e]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        const uint32_t tot_t = stage + (uint32_t)n_chunks;
        phase ^= (tot_t / STAGES) & 1;
        stage = tot_t % STAGES;
        continue;
      }
      const long long rbase = m0 + q * 32;
      int nrows = (int)(args.M - rbase < 32 ? (args.M - rbase < 0 ? 0 : args.M - rbase) : 32);
      const int ncols = (int)(args.N - n0 < args.wn ? args.N - n0 : args.wn);
      for (int c0 = half * 32; c0 < ncols; c0 += 64) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * MAX_N + c0), v);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; ++e) sc[lane * 33 + e] = __uint_as_float(v[e]);
        __syncwarp();
        const bool colok = c0 + lane < ncols;
        if (args.ksplit > 1) {
          float* p = args.part + ((long long)ks * args.M + rbase) * args.N + n0 + c0 + lane;
#pragma unroll 8
          for (int u = 0; u < 32; ++u)
            if (colok && u < nrows) p[(long long)u * args.N] = sc[u * 33 + lane];
        } else {
          float* p = args.C + rbase * args.ldc + n0 + c0 + lane;
          const float b = (args.bias && colok) ? __ldg(args.bias + n0 + c0 + lane) : 0.f;
          if (args.had) {
            const float* h = args.had + rbase * args.ldh + n0 + c0 + lane;
            if (args.beta == 1.f) {          // fire-and-forget adds in L2: one add per element per launch (deterministic)
#pragma unroll 8
              for (int u = 0; u < 32; ++u)
                if (colok && u < nrows) atomicAdd(&p[(long long)u * args.ldc], args.alpha * sc[u * 33 + lane] * __ldg(h + (long long)u * args.ldh));
            } else {
              float old[32], hv[32];                       // loads before stores (see the plain beta branch below)
#pragma unroll
              for (int u = 0; u < 32; ++u) {
                const bool on = colok && u < nrows;
                hv[u] = on ? __ldg(h + (long long)u * args.ldh) : 0.f;
                old[u] = (on && args.beta != 0.f) ? p[(long long)u * args.ldc] : 0.f;
              }
#pragma unroll
              for (int u = 0; u < 32; ++u)
                if (colok && u < nrows) {
                  const float v = args.alpha * sc[u * 33 + lane] * hv[u];
                  p[(long long)u * args.ldc] = args.beta == 0.f ? v : fmaf(args.beta, old[u], v);
                }
            }
          } else if (args.beta == 0.f) {
#pragma unroll 8
            for (int u = 0; u < 32; ++u)
              if (colok && u < nrows) p[(long long)u * args.ldc] = fmaf(args.alpha, sc[u * 33 + lane], b);
          } else {
            // all 32 loads of C first: interleaved with the stores to the same array they serialise (one DRAM round trip per
            // row, 71 us per tile -- the Linear-layer A-factor Grams of a 512-example batch took 3 x 71 us for 4 us of math)
            float old[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) old[u] = (colok && u < nrows) ? p[(long long)u * args.ldc] : 0.f;
#pragma unroll
            for (int u = 0; u < 32; ++u)
              if (colok && u < nrows) p[(long long)u * args.ldc] = fmaf(args.alpha, sc[u * 33 + lane], args.beta * old[u]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    const uint32_t tot = stage + (uint32_t)n_chunks;
    phase ^= (tot / STAGES) & 1;
    stage = tot % STAGES;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// C = alpha * sum_ks part[ks] + beta * C   (fixed order)
__global__ void reduce_partials_kernel(const float* __restrict__ part, int ksplit, long long M, long long N, float alpha,
                                       float beta, float* __restrict__ C, long long ldc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float acc = 0.f;
  for (int k = 0; k < ksplit; ++k) acc += part[(long long)k * M * N + i];
  float* q = C + (i / N) * ldc + i % N;
  *q = (beta == 0.f) ? alpha * acc : fmaf(alpha, acc, beta * (*q));
}

// ---- operand preparation ------------------------------------------------------------------------------------------
// hi/lo[r, c] = bf16 split of f(src[r*lds + c]) for c < cols, 0 for cols <= c < pitch   (k = c contiguous).
// One thread per 8 consecutive columns (pitch is a multiple of 8): two 16-byte stores; 16-byte loads when the source
// row allows it.
template <bool SQUARE>
__global__ void split_rows_kernel(const float* __restrict__ src, long long lds, long long rows, long long cols,
                                  long long pitch, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long p8 = pitch >> 3;
  if (i >= rows * p8) return;
  const long long r = i / p8, c = (i % p8) << 3;
  const float* q = src + r * lds + c;
  float v[8];
  if (c + 8 <= cols && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(q)), y = __ldg(reinterpret_cast<const float4*>(q) + 1);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = c + e < cols ? __ldg(q + e) : 0.f;
  }
  if (SQUARE) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= v[e];
  }
  uint4 h, l;
  split8(v, h, l);
  *reinterpret_cast<uint4*>(hi + r * pitch + c) = h;
  *reinterpret_cast<uint4*>(lo + r * pitch + c) = l;
}
// hi/lo[c, r] = bf16 split of f(src[r*lds + c])  (k = r contiguous, pitch >= rows, zero padded); 32 x 32 smem transpose
template <bool SQUARE>
__global__ void split_transposed_kernel(const float* __restrict__ src, long long lds, long long rows, long long cols,
                                        long long pitch, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const long long c0 = (long long)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < pitch) {
      float v = tile[threadIdx.x][i];
      if (SQUARE) v *= v;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[c * pitch + r] = h;
      lo[c * pitch + r] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// Conv gradients gout[r, co, pos] (row pitch gld) -> hi/lo[(r * L + pos), co] (row pitch Cp >= Cout, zero padded): the
// K-major A operand of the backward-data product dU[(r,pos), e] = sum_co gout[r, co, pos] W[co, e].  32 x 32 smem transposes.
__global__ void split_conv_grad_kernel(const float* __restrict__ g, long long gld, int Cout, int L, int Cp,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[32][33];
  const long long r = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < Cout && p < L) ? g[r * gld + (long long)c * L + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < L && c < Cp) {
      const float v = tile[threadIdx.x][i];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const long long o = (r * L + p) * Cp + c;
      hi[o] = h;
      lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// [rows, K] bf16, row pitch `ld` elements; elements beyond K (and rows beyond `rows`) arrive as zeros
static int make_map(CUtensorMap* map, const void* base, long long rows, long long K, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return BNNP_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[bnnp_b200] cuTensorMapEncodeTiled failed: %d (rows %lld K %lld ld %lld box %d)\n", (int)r, rows, K, ld,
            box_rows);
    return BNNP_ERR_CUDA;
  }
  return BNNP_OK;
}

static int pick_wn(long long N) {
  const long long nt = (N + MAX_N - 1) / MAX_N;
  const long long w = ((N + nt - 1) / nt + 15) / 16 * 16;
  return (int)(w < 16 ? 16 : w);
}

}  // namespace tg

int gemm_tma_ksplit(int64_t M, int64_t N, int64_t K) {
  using namespace tg;
  const long long chunks = (K + BK - 1) / BK;
  const long long tiles = ((M + TILE_M - 1) / TILE_M) * ((N + pick_wn(N) - 1) / pick_wn(N));
  long long want = (148 * 2 + tiles - 1) / tiles;                   // fill the machine twice over
  const long long maxs = chunks / 16 > 0 ? chunks / 16 : 1;         // at least 16 chunks per split
  long long s = want < maxs ? want : maxs;
  const long long need = (chunks + kMaxChunksPerAccum - 1) / kMaxChunksPerAccum;     // accuracy floor
  if (s < need) s = need;
  return (int)(s < 1 ? 1 : (s > 512 ? 512 : s));
}

int64_t gemm_tma_scratch_floats(int64_t M, int64_t N, int64_t K) {
  const int ks = gemm_tma_ksplit(M, N, K);
  return ks > 1 ? (int64_t)ks * M * N + 64 : 64;
}

int gemm_tma_nt(int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* Ah, const __nv_bfloat16* Al, int64_t lda,
                const __nv_bfloat16* Bh, const __nv_bfloat16* Bl, int64_t ldb, float beta, float* C, int64_t ldc,
                float* scratch, cudaStream_t st, const float* bias, int ct_L, const float* had, int64_t ldh) {
  using namespace tg;
  if (M <= 0 || N <= 0 || K <= 0) return BNNP_OK;
  if (!Ah || !Al || !Bh || !Bl || !C || (lda & 7) || (ldb & 7)) return BNNP_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(Ah) | reinterpret_cast<uintptr_t>(Al) | reinterpret_cast<uintptr_t>(Bh) |
       reinterpret_cast<uintptr_t>(Bl)) & 15)
    return BNNP_ERR_INVALID;
  Args a;
  a.M = M; a.N = N;
  a.wn = pick_wn(N);
  a.n_mt = (int)((M + TILE_M - 1) / TILE_M);
  a.n_nt = (int)((N + a.wn - 1) / a.wn);
  a.chunks_total = (int)((K + BK - 1) / BK);
  a.ksplit = scratch ? gemm_tma_ksplit(M, N, K) : 1;
  if (a.ksplit > a.chunks_total) a.ksplit = a.chunks_total;
  if (!scratch && a.chunks_total > kMaxChunksPerAccum) return BNNP_ERR_INVALID;      // a long K needs the partials
  a.alpha = alpha; a.beta = beta; a.C = C; a.ldc = ldc; a.part = scratch;
  a.bias = bias;
  a.ct_L = ct_L;
  a.had = had; a.ldh = ldh;
  if (had && (a.ksplit > 1 || bias || ct_L > 0)) return BNNP_ERR_INVALID;
  if (bias && (a.ksplit > 1 || beta != 0.f)) return BNNP_ERR_INVALID;
  if (ct_L > 0 && (a.ksplit > 1 || beta != 0.f || bias)) return BNNP_ERR_INVALID;
  CUtensorMap mah, mal, mbh, mbl;
  int rc;
  if ((rc = make_map(&mah, Ah, M, K, lda, TILE_M))) return rc;
  if ((rc = make_map(&mal, Al, M, K, lda, TILE_M))) return rc;
  if ((rc = make_map(&mbh, Bh, N, K, ldb, a.wn))) return rc;
  if ((rc = make_map(&mbl, Bl, N, K, ldb, a.wn))) return rc;
  static bool attr = false;
  if (!attr) {
    BNNP_CUDA_CHECK(cudaFuncSetAttribute(gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const long long n_tiles = (long long)a.n_mt * a.n_nt * a.ksplit;
  gemm_tma_kernel<<<(int)(n_tiles < sms ? n_tiles : sms), NUM_THREADS, SMEM_BYTES, st>>>(mah, mal, mbh, mbl, a);
  BNNP_LAUNCH_CHECK();
  if (a.ksplit > 1) {
    reduce_partials_kernel<<<(unsigned)ceil_div64(M * N, 256), 256, 0, st>>>(scratch, a.ksplit, M, N, alpha, beta, C, ldc);
    BNNP_LAUNCH_CHECK();
  }
  return BNNP_OK;
}

int split_rows(const float* src, int64_t lds, int64_t rows, int64_t cols, int64_t pitch, __nv_bfloat16* hi,
               __nv_bfloat16* lo, bool square, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return BNNP_OK;
  if (pitch < cols || (pitch & 7)) return BNNP_ERR_INVALID;
  const unsigned grid = (unsigned)ceil_div64(rows * (pitch >> 3), 256);
  if (square) tg::split_rows_kernel<true><<<grid, 256, 0, st>>>(src, lds, rows, cols, pitch, hi, lo);
  else tg::split_rows_kernel<false><<<grid, 256, 0, st>>>(src, lds, rows, cols, pitch, hi, lo);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

int split_transposed(const float* src, int64_t lds, int64_t rows, int64_t cols, int64_t pitch, __nv_bfloat16* hi,
                     __nv_bfloat16* lo, bool square, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return BNNP_OK;
  if (pitch < rows || (pitch & 7)) return BNNP_ERR_INVALID;
  const dim3 grid((unsigned)ceil_div64(pitch, 32), (unsigned)ceil_div64(cols, 32));
  if (grid.y > 65535u) return BNNP_ERR_UNSUPPORTED;
  if (square) tg::split_transposed_kernel<true><<<grid, dim3(32, 8), 0, st>>>(src, lds, rows, cols, pitch, hi, lo);
  else tg::split_transposed_kernel<false><<<grid, dim3(32, 8), 0, st>>>(src, lds, rows, cols, pitch, hi, lo);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

static inline int64_t up8(int64_t v) { return (v + 7) / 8 * 8; }
static inline int64_t up64(int64_t v) { return (v + 63) / 64 * 64; }

int64_t gram_tma_scratch_floats(int64_t R, int64_t m) {
  return up64(m * up8(R)) + 64 + gemm_tma_scratch_floats(m, m, R);
}
int gram_tma(const float* S, int64_t lds, int64_t R, int64_t m, float alpha, float beta, float* C, int64_t ldc,
             float* scratch, cudaStream_t st) {
  const int64_t Rp = up8(R);
  // scratch may start anywhere: align the split copies to 16 bytes
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 15) & ~uintptr_t(15));
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* lo = hi + m * Rp;
  float* part = base + up64(m * Rp);
  int rc = split_transposed(S, lds, R, m, Rp, hi, lo, false, st);
  if (rc) return rc;
  return gemm_tma_nt(m, m, R, alpha, hi, lo, Rp, hi, lo, Rp, beta, C, ldc, part, st);
}

// conv backward-data product on TMA-fed tensor cores: dUt[r, e, pos] = sum_co gout[r, co, pos] W[co, e]
// (W [Cout, E] row-major; the col2im gather that follows reads dUt).  scratch: conv_bwd_tma_scratch_floats().
int64_t conv_bwd_tma_scratch_floats(int64_t R, int64_t Cout, int64_t L, int64_t E) {
  return up64(R * L * up8(Cout)) + up64(E * up8(Cout)) + 128;
}
int conv_bwd_tma(const float* gout, int64_t gld, int64_t R, int Cout, int L, const float* W, int E, float* dUt,
                 float* scratch, cudaStream_t st) {
  if (Cout > 4096 || R > 65535) return BNNP_ERR_UNSUPPORTED;
  const int Cp = (int)up8(Cout);
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 15) & ~uintptr_t(15));
  __nv_bfloat16* ah = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* al = ah + R * L * Cp;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(base + up64(R * L * Cp));
  __nv_bfloat16* wl = wh + (int64_t)E * Cp;
  tg::split_conv_grad_kernel<<<dim3((unsigned)ceil_div64(L, 32), (unsigned)ceil_div64(Cp, 32), (unsigned)R), dim3(32, 8), 0, st>>>(
      gout, gld, Cout, L, Cp, ah, al);
  BNNP_LAUNCH_CHECK();
  int rc = split_transposed(W, E, Cout, E, Cp, wh, wl, false, st);          // [E, Cp]: k = co contiguous
  if (rc) return rc;
  return gemm_tma_nt(R * L, E, Cout, 1.f, ah, al, Cp, wh, wl, Cp, 0.f, dUt, 0, nullptr, st, nullptr, L);
}

int64_t linear_tma_scratch_floats(int64_t R, int64_t k, int64_t n_out) {
  return up64(R * up8(k)) + up64(n_out * up8(k)) + 128;
}
int linear_tma(const float* X, int64_t ldx, int64_t R, int64_t k, const float* W, int64_t ldw, int64_t n_out,
               int w_transposed, const float* bias, float* Y, int64_t ldy, float* scratch, cudaStream_t st,
               int square_x) {
  if (k > 4096) return BNNP_ERR_UNSUPPORTED;             // one TMEM accumulation per output (no partials here)
  const int64_t kp = up8(k);
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 15) & ~uintptr_t(15));
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* xl = xh + R * kp;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(base + up64(R * kp));
  __nv_bfloat16* wl = wh + n_out * kp;
  int rc = split_rows(X, ldx, R, k, kp, xh, xl, square_x != 0, st);
  if (rc) return rc;
  rc = w_transposed ? split_transposed(W, ldw, k, n_out, kp, wh, wl, false, st)
                    : split_rows(W, ldw, n_out, k, kp, wh, wl, false, st);
  if (rc) return rc;
  return gemm_tma_nt(R, n_out, k, 1.f, xh, xl, kp, wh, wl, kp, 0.f, Y, ldy, nullptr, st, bias);
}

}  // namespace bnnp

using namespace bnnp;

extern "C" int64_t bnnp_gemm_nt_scratch_floats(int64_t M, int64_t N, int64_t K) { return gemm_tma_scratch_floats(M, N, K); }

extern "C" int bnnp_gemm_nt_split(int64_t M, int64_t N, int64_t K, float alpha, const void* Ah, const void* Al, int64_t lda,
                                  const void* Bh, const void* Bl, int64_t ldb, float beta, float* C, int64_t ldc,
                                  float* scratch, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return BNNP_ERR_INVALID;
  return gemm_tma_nt(M, N, K, alpha, reinterpret_cast<const __nv_bfloat16*>(Ah), reinterpret_cast<const __nv_bfloat16*>(Al),
                     lda, reinterpret_cast<const __nv_bfloat16*>(Bh), reinterpret_cast<const __nv_bfloat16*>(Bl), ldb, beta, C,
                     ldc, scratch, as_stream(stream), nullptr, 0);
}

extern "C" int bnnp_split_bf16_t(const float* src, int64_t ld_src, int64_t rows, int64_t cols, int64_t pitch, void* hi,
                                 void* lo, void* stream) {
  if (!src || !hi || !lo || rows <= 0 || cols <= 0) return BNNP_ERR_INVALID;
  return split_transposed(src, ld_src, rows, cols, pitch, reinterpret_cast<__nv_bfloat16*>(hi),
                          reinterpret_cast<__nv_bfloat16*>(lo), false, as_stream(stream));
}

extern "C" int bnnp_split_bf16_rows(const float* src, int64_t ld_src, int64_t rows, int64_t cols, int64_t pitch, void* hi,
                                    void* lo, void* stream) {
  if (!src || !hi || !lo || rows <= 0 || cols <= 0) return BNNP_ERR_INVALID;
  return split_rows(src, ld_src, rows, cols, pitch, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo),
                    false, as_stream(stream));
}

extern "C" int bnnp_gemm_nt_split_had(int64_t M, int64_t N, int64_t K, float alpha, const void* Ah, const void* Al, int64_t lda,
                                      const void* Bh, const void* Bl, int64_t ldb, float beta, float* C, int64_t ldc,
                                      const float* had, int64_t ldh, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || !had) return BNNP_ERR_INVALID;
  return gemm_tma_nt(M, N, K, alpha, reinterpret_cast<const __nv_bfloat16*>(Ah), reinterpret_cast<const __nv_bfloat16*>(Al),
                     lda, reinterpret_cast<const __nv_bfloat16*>(Bh), reinterpret_cast<const __nv_bfloat16*>(Bl), ldb, beta, C,
                     ldc, nullptr, as_stream(stream), nullptr, 0, had, ldh);
}
