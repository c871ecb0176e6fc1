// Tensor-core GEMM of the one-off factorisation (preds/laplace.py:43-45: chol(H), cholesky_inverse, chol(Sigma)):
//
//     C[M,N] = alpha * sum_k A[m,k] * B[n,k] + beta * C          ("NT": both operands K-major)
//
// with errors at the level of an fp32 library GEMM, which the two-term split of gemm_tma.cu cannot give (2^-17 per
// product, amplified by the conditioning of H: measured residual |H Sigma - I| 1.8e-3 against potri's 2.2e-4).  Two things
// are different here:
//
//  * THREE bf16 terms per operand (hi + mid + lo = 24 significand bits, i.e. the fp32 value itself) and six MMAs per
//    16-wide K-step: hi*hi into a MAIN accumulator; hi*mid, mid*hi, mid*mid, hi*lo, lo*hi into a separate CORRECTION
//    accumulator.  The tensor core adds into its fp32 accumulator with truncation (scripts/tmem_trunc_probe.py: -3e-8
//    relative per dependent add, a bias for sums of one sign).  An add of a 2^-8-sized correction into the main accumulator
//    would cost a full-sized truncation each; in their own accumulator the corrections truncate relative to their own
//    magnitude (2^-8 of the result), which is negligible.
//  * the MAIN accumulation is cut into chunks of 128 k (8 dependent adds: measured exact-to-rounding, bias < 1e-8); the four
//    epilogue warps drain every chunk from TMEM (double-buffered, under the MMAs of the next chunk) and keep the running
//    sum of the tile in registers with round-to-nearest fp32 adds -- 128 x 128 tile, 128 registers per thread.
//
// Triangular structure is exploited per tile: `lower` computes only tiles on/below the diagonal (SYRK-shaped updates),
// `ktri` starts the K loop of a tile at its row / column origin (operands that are triangular matrices).
#include <math.h>
#include "tc_common.cuh"
#include "gemm3.cuh"

namespace bnnp {
namespace t3 {

using namespace bnnp::tcp;

constexpr int BK = 32, TILE = 128, STAGES = 4, GROUP = 4;           // one main accumulation = GROUP stages = 128 k
constexpr int TERM = TILE * BK * 2;                                 // one bf16 plane of one operand tile: 8 KB
constexpr int A_STAGE = 3 * TERM, STAGE_BYTES = 6 * TERM;           // 48 KB
constexpr int EPI_TILE_BYTES = 32 * 33 * 4 + 32;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + 4 * EPI_TILE_BYTES;
constexpr int NUM_THREADS = 256;      // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-7: epilogue
constexpr int EPI_WARP0 = 4;
constexpr int kMaxStagesPerTile = 4096 / BK;                        // the correction accumulator spans at most 4096 k

struct Args {
  long long M, N, K;
  int n_mt, n_nt, lower, ktri;
  long long koff;
  long long n_tiles;
  float alpha, beta;
  float* C; long long ldc;
  int own_mod, own_rank, own_dim;   // own_mod > 1: this process computes only the tiles whose global column (own_dim 0) /
  long long own_off;                //   row (own_dim 1) tile index (local index + own_off) is congruent to own_rank
};

struct Maps { CUtensorMap a[3], b[3]; };

// Tile order: SUPER x SUPER super-tiles, row fastest inside one, so that the ~148 tiles in flight share 2 x SUPER operand
// row blocks (24 MB of planes at K = 1024) instead of sweeping a whole operand per tile row -- the planes of a 32768-row
// panel are 201 MB, more than L2 (measured before: L2 hit rate 65 %, 15.9 GB of DRAM reads for a 2.1 GB update).
constexpr int SUPER = 16;
__device__ __forceinline__ long long tri_row(long long t) {            // largest r with r (r + 1) / 2 <= t
  long long r = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while (r * (r + 1) / 2 > t) --r;
  while ((r + 1) * (r + 2) / 2 <= t) ++r;
  return r;
}
__device__ __forceinline__ void tile_of(const Args& g, long long t, int& mt, int& nt) {
  if (g.lower) {
    const long long I = tri_row(t) / SUPER;                            // super-row: tile rows SUPER*I ..
    const long long r0 = I * SUPER;
    const long long u = t - r0 * (r0 + 1) / 2;
    const long long h = g.n_mt - r0 < SUPER ? g.n_mt - r0 : SUPER;
    if (u < I * h * SUPER) {                                           // a full super-tile left of the diagonal
      const long long J = u / (h * SUPER), v = u - J * h * SUPER;
      mt = (int)(r0 + v % h);
      nt = (int)(J * SUPER + v / h);
    } else {                                                           // the triangular super-tile on the diagonal
      const long long w = u - I * h * SUPER;
      // rows inside come in blocks: local row lr holds lr + 1 tiles, but the enumeration above put rows r0.. in order of
      // the plain triangle minus the I*SUPER columns each of them has to the left
      const long long lr = tri_row(w);
      mt = (int)(r0 + lr);
      nt = (int)(r0 + (w - lr * (lr + 1) / 2));
    }
  } else {
    const long long per_row = (long long)SUPER * g.n_nt;
    const long long I = t / per_row, u = t - I * per_row;
    const long long r0 = I * SUPER;
    const long long h = g.n_mt - r0 < SUPER ? g.n_mt - r0 : SUPER;
    const long long J = u / (h * SUPER), v = u - J * h * SUPER;
    mt = (int)(r0 + v % h);
    nt = (int)(J * SUPER + v / h);
  }
}
// tiles of other ranks (the factorisation shared between the GPUs of a data-parallel job: every tile has one owner, fixed in
// global tile coordinates, so a rank's copy of the matrix is always complete on the tiles it owns)
__device__ __forceinline__ bool tile_is_mine(const Args& g, int mt, int nt) {
  if (g.own_mod <= 1) return true;
  const long long i = (g.own_dim ? (long long)mt : (long long)nt) + g.own_off;
  return (int)(i % g.own_mod) == g.own_rank;
}
// first K-stage of a tile (operands that are triangular: everything before the tile's origin is zero)
__device__ __forceinline__ int first_stage(const Args& g, int mt, int nt) {
  long long k = 0;
  if (g.ktri == 1) k = (long long)nt * TILE + g.koff;
  else if (g.ktri == 2) k = (long long)mt * TILE + g.koff;
  if (k < 0) k = 0;
  const long long tot = (g.K + BK - 1) / BK;
  const long long s = k / BK;
  return (int)(s < tot ? s : tot);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm3_kernel(const __grid_constant__ Maps maps, Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                          // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;                // [STAGES] MMA commit -> TMA
  uint64_t* main_full = bars + 2 * STAGES;        // [2] MMA commit -> epilogue (one chunk of the main accumulation)
  uint64_t* main_empty = main_full + 2;           // [2] epilogue (4 warp arrivals) -> MMA
  uint64_t* corr_full = main_full + 4;            // [2] MMA commit -> epilogue (corrections of a whole tile)
  uint64_t* corr_empty = main_full + 6;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(main_full + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&main_full[b], 1); mbar_init(&main_empty[b], 4);
      mbar_init(&corr_full[b], 1); mbar_init(&corr_empty[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tot_stages = (int)((args.K + BK - 1) / BK);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (long long tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
        int mt, nt;
        tile_of(args, tile, mt, nt);
        if (!tile_is_mine(args, mt, nt)) continue;
        const int s_beg = first_stage(args, mt, nt);
        for (int c = s_beg; c < tot_stages; ++c) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* base = smem + s * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], (uint32_t)STAGE_BYTES);
          const int k0 = c * BK;
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            tma_load_2d(&maps.a[p], &full[s], base + p * TERM, k0, mt * TILE);
            tma_load_2d(&maps.b[p], &full[s], base + A_STAGE + p * TERM, k0, nt * TILE);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0, gc = 0, cc = 0;
      const uint32_t idesc = umma_idesc_bf16(TILE);
      for (long long tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
        int mt, nt;
        tile_of(args, tile, mt, nt);
        if (!tile_is_mine(args, mt, nt)) continue;
        const int n_st = tot_stages - first_stage(args, mt, nt);
        if (n_st <= 0) continue;
        const uint32_t cb = cc & 1;
        mbar_wait(&corr_empty[cb], ((cc >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_corr = tmem_base + 256 + cb * TILE;
        for (int c = 0; c < n_st;) {
          const uint32_t mb = gc & 1;
          mbar_wait(&main_empty[mb], ((gc >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_main = tmem_base + mb * TILE;
          const int c_end = c + GROUP < n_st ? c + GROUP : n_st;
          for (int j = 0; c < c_end; ++c, ++j) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + A_STAGE;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              const uint64_t ah = umma_desc_sw64(a0 + kk * 32), am = umma_desc_sw64(a0 + TERM + kk * 32),
                             al = umma_desc_sw64(a0 + 2 * TERM + kk * 32);
              const uint64_t bh = umma_desc_sw64(b0 + kk * 32), bm = umma_desc_sw64(b0 + TERM + kk * 32),
                             bl = umma_desc_sw64(b0 + 2 * TERM + kk * 32);
              umma_bf16(d_corr, al, bh, idesc, (c | kk) ? 1u : 0u);       // smallest terms first
              umma_bf16(d_corr, ah, bl, idesc, 1u);
              umma_bf16(d_corr, am, bm, idesc, 1u);
              umma_bf16(d_corr, am, bh, idesc, 1u);
              umma_bf16(d_corr, ah, bm, idesc, 1u);
              umma_bf16(d_main, ah, bh, idesc, (j | kk) ? 1u : 0u);
            }
            umma_commit(&empty[s]);
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
          umma_commit(&main_full[mb]);
          ++gc;
        }
        umma_commit(&corr_full[cb]);
        ++cc;
      }
    }
  } else if (warp >= EPI_WARP0) {
    const int q = warp - EPI_WARP0;
    float* sc = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256 + q * EPI_TILE_BYTES);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t gc = 0, cc = 0;
    for (long long tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
      int mt, nt;
      tile_of(args, tile, mt, nt);
      if (!tile_is_mine(args, mt, nt)) continue;
      const int n_st = tot_stages - first_stage(args, mt, nt);
      float acc[TILE];
#pragma unroll
      for (int e = 0; e < TILE; ++e) acc[e] = 0.f;
      if (n_st > 0) {
        const int n_groups = (n_st + GROUP - 1) / GROUP;
        for (int g = 0; g < n_groups; ++g, ++gc) {
          const uint32_t mb = gc & 1;
          mbar_wait(&main_full[mb], (gc >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < TILE; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(lane_base + mb * TILE + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[c0 + e] += __uint_as_float(v[e]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&main_empty[mb]);
        }
        const uint32_t cb = cc & 1;
        mbar_wait(&corr_full[cb], (cc >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < TILE; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(lane_base + 256 + cb * TILE + c0, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[c0 + e] += __uint_as_float(v[e]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&corr_empty[cb]);
        ++cc;
      } else if (args.beta == 1.f) {
        continue;                                   // nothing to add
      }
      // store: per-warp shared-memory transpose (lane = column -> 128 contiguous bytes per store instruction)
      const long long rbase = (long long)mt * TILE + q * 32, n0 = (long long)nt * TILE;
      const long long left = args.M - rbase;
      const int nrows = (int)(left < 0 ? 0 : (left < 32 ? left : 32));
      const int ncols = (int)(args.N - n0 < TILE ? args.N - n0 : TILE);
#pragma unroll
      for (int c0 = 0; c0 < TILE; c0 += 32) {
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; ++e) sc[lane * 33 + e] = acc[c0 + e];
        __syncwarp();
        if (c0 + lane < ncols) {
          float* p = args.C + rbase * args.ldc + n0 + c0 + lane;
          if (args.beta == 0.f) {
#pragma unroll 8
            for (int u = 0; u < 32; ++u)
              if (u < nrows) p[(long long)u * args.ldc] = args.alpha * sc[u * 33 + lane];
          } else if (args.beta == 1.f) {
            // fire-and-forget adds in L2 (RED.ADD.F32; every element is updated once per launch, so the result is
            // deterministic): a load + store loop here is latency-bound and stalled the MMAs (measured: 16.6 ms for the
            // lower half of a product whose full version took 12.6 ms with plain stores)
#pragma unroll 8
            for (int u = 0; u < 32; ++u)
              if (u < nrows) atomicAdd(&p[(long long)u * args.ldc], args.alpha * sc[u * 33 + lane]);
          } else {
#pragma unroll 8
            for (int u = 0; u < 32; ++u)
              if (u < nrows) p[(long long)u * args.ldc] = fmaf(args.alpha, sc[u * 33 + lane], args.beta * p[(long long)u * args.ldc]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// planes[p][r, c] (p = 0 hi, 1 mid, 2 lo; row pitch `pitch`, plane stride `pstride` elements) <- three-term bf16 split of
// src[r*lds + c]; columns cols..up8(cols)-1 are zeroed, nothing beyond is touched (the destination may be a block of a larger
// plane set).  One thread per 8 columns: three 16-byte stores.
__global__ void split3_rows_kernel(const float* __restrict__ src, long long lds, long long rows, long long cols,
                                   long long pitch, __nv_bfloat16* __restrict__ planes, long long pstride, float scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long p8 = (cols + 7) >> 3;
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
  __nv_bfloat16 h[8], m[8], l[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float x = v[e] * scale;
    h[e] = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h[e]);            // exact
    m[e] = __float2bfloat16_rn(r1);
    l[e] = __float2bfloat16_rn(r1 - __bfloat162float(m[e]));
  }
  __nv_bfloat16* o = planes + r * pitch + c;
  *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(h);
  *reinterpret_cast<uint4*>(o + pstride) = *reinterpret_cast<uint4*>(m);
  *reinterpret_cast<uint4*>(o + 2 * pstride) = *reinterpret_cast<uint4*>(l);
}

// planes[p][c, r] <- split of src[r*lds + c]  (transposed copy; 32 x 32 shared-memory transposes)
__global__ void split3_transposed_kernel(const float* __restrict__ src, long long lds, long long rows, long long cols,
                                         long long pitch, __nv_bfloat16* __restrict__ planes, long long pstride, float scale) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32, c0 = (long long)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < ((rows + 7) & ~7LL)) {
      const float x = tile[threadIdx.x][i] * scale;
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      const float r1 = x - __bfloat162float(h);
      const __nv_bfloat16 m = __float2bfloat16_rn(r1);
      __nv_bfloat16* o = planes + c * pitch + r;
      o[0] = h;
      o[pstride] = m;
      o[2 * pstride] = __float2bfloat16_rn(r1 - __bfloat162float(m));
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
static int make_map(CUtensorMap* map, const void* base, long long rows, long long K, long long ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return BNNP_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)TILE};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[bnnp_b200] cuTensorMapEncodeTiled (gemm3) failed: %d (rows %lld K %lld ld %lld)\n", (int)r, rows, K, ld);
    return BNNP_ERR_CUDA;
  }
  return BNNP_OK;
}

}  // namespace t3

int gemm3_nt(int64_t M, int64_t N, int64_t K, float alpha, const __nv_bfloat16* A, int64_t lda, int64_t a_pstride,
             const __nv_bfloat16* B, int64_t ldb, int64_t b_pstride, float beta, float* C, int64_t ldc, int lower, int ktri,
             int64_t koff, cudaStream_t st, int own_mod, int own_rank, int own_dim, int64_t own_off) {
  using namespace t3;
  if (M <= 0 || N <= 0 || K <= 0) return BNNP_OK;
  if (!A || !B || !C || (lda & 7) || (ldb & 7) || (a_pstride & 7) || (b_pstride & 7)) return BNNP_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) return BNNP_ERR_INVALID;
  if (lower && M != N) return BNNP_ERR_INVALID;
  if (ktri < 0 || ktri > 2) return BNNP_ERR_INVALID;
  if ((K + BK - 1) / BK > kMaxStagesPerTile) return BNNP_ERR_INVALID;    // walk longer K in segments (beta = 1)
  Args a;
  a.M = M; a.N = N; a.K = K;
  a.n_mt = (int)((M + TILE - 1) / TILE);
  a.n_nt = (int)((N + TILE - 1) / TILE);
  a.lower = lower; a.ktri = ktri; a.koff = koff;
  a.n_tiles = lower ? (long long)a.n_mt * (a.n_mt + 1) / 2 : (long long)a.n_mt * a.n_nt;
  a.alpha = alpha; a.beta = beta; a.C = C; a.ldc = ldc;
  if (own_mod > 1 && (own_rank < 0 || own_rank >= own_mod || own_off < 0)) return BNNP_ERR_INVALID;
  a.own_mod = own_mod; a.own_rank = own_rank; a.own_dim = own_dim; a.own_off = own_off;
  Maps maps;
  int rc;
  for (int p = 0; p < 3; ++p) {
    if ((rc = make_map(&maps.a[p], A + p * a_pstride, M, K, lda))) return rc;
    if ((rc = make_map(&maps.b[p], B + p * b_pstride, N, K, ldb))) return rc;
  }
  static bool attr = false;
  if (!attr) {
    BNNP_CUDA_CHECK(cudaFuncSetAttribute(gemm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  gemm3_kernel<<<(int)(a.n_tiles < sms ? a.n_tiles : sms), NUM_THREADS, SMEM_BYTES, st>>>(maps, a);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

int split3(const float* src, int64_t lds, int64_t rows, int64_t cols, __nv_bfloat16* planes, int64_t pitch, int64_t pstride,
           int transposed, float scale, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return BNNP_OK;
  if ((pitch & 7) || (pstride & 7) || (reinterpret_cast<uintptr_t>(planes) & 15)) return BNNP_ERR_INVALID;
  if (!transposed) {
    if (pitch < ((cols + 7) & ~int64_t(7))) return BNNP_ERR_INVALID;
    t3::split3_rows_kernel<<<(unsigned)ceil_div64(rows * ((cols + 7) >> 3), 256), 256, 0, st>>>(src, lds, rows, cols, pitch, planes,
                                                                                          pstride, scale);
  } else {
    if (pitch < ((rows + 7) & ~int64_t(7))) return BNNP_ERR_INVALID;
    const dim3 grid((unsigned)ceil_div64(rows, 32), (unsigned)ceil_div64(cols, 32));
    if (grid.y > 65535u) return BNNP_ERR_UNSUPPORTED;
    t3::split3_transposed_kernel<<<grid, dim3(32, 8), 0, st>>>(src, lds, rows, cols, pitch, planes, pstride, scale);
  }
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

}  // namespace bnnp

using namespace bnnp;

extern "C" int bnnp_split3_bf16(const float* src, int64_t ld_src, int64_t rows, int64_t cols, void* planes, int64_t pitch,
                                int64_t plane_stride, int transposed, float scale, void* stream) {
  if (!src || !planes || rows <= 0 || cols <= 0) return BNNP_ERR_INVALID;
  return split3(src, ld_src, rows, cols, reinterpret_cast<__nv_bfloat16*>(planes), pitch, plane_stride, transposed, scale,
                as_stream(stream));
}

extern "C" int bnnp_gemm3_nt(int64_t M, int64_t N, int64_t K, float alpha, const void* A, int64_t lda, int64_t a_plane_stride,
                             const void* B, int64_t ldb, int64_t b_plane_stride, float beta, float* C, int64_t ldc, int lower,
                             int ktri, int64_t koff, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return BNNP_ERR_INVALID;
  return gemm3_nt(M, N, K, alpha, reinterpret_cast<const __nv_bfloat16*>(A), lda, a_plane_stride,
                  reinterpret_cast<const __nv_bfloat16*>(B), ldb, b_plane_stride, beta, C, ldc, lower, ktri, koff,
                  as_stream(stream), 1, 0, 0, 0);
}

extern "C" int bnnp_gemm3_nt_owned(int64_t M, int64_t N, int64_t K, float alpha, const void* A, int64_t lda,
                                   int64_t a_plane_stride, const void* B, int64_t ldb, int64_t b_plane_stride, float beta,
                                   float* C, int64_t ldc, int lower, int ktri, int64_t koff, int own_mod, int own_rank,
                                   int own_dim, int64_t own_off, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return BNNP_ERR_INVALID;
  return gemm3_nt(M, N, K, alpha, reinterpret_cast<const __nv_bfloat16*>(A), lda, a_plane_stride,
                  reinterpret_cast<const __nv_bfloat16*>(B), ldb, b_plane_stride, beta, C, ldc, lower, ktri, koff,
                  as_stream(stream), own_mod, own_rank, own_dim, own_off);
}
