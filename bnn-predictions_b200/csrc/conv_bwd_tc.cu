// Backward-data of a stride-1 convolution as an IMPLICIT GEMM on TMA-fed tensor cores (K-column backward of the conv nets,
// BackPACK's sqrt-GGN / Jacobian back-propagation behind preds/laplace.py:47-78):
//
//     dX[r, ci, y, x] = sum_{a, b} sum_co g[r, co, y + pt - a, x + pl - b] * W[co, ci, a, b]
//
// Until now this ran as a plain product dU[(r,pos), (ci,a,b)] = g W followed by a col2im gather: the [R, E, L] buffer between
// the two is 1.7 GB for CIFAR10Net's conv2 at 512 examples (written by the product, re-read by the gather at 2.9 TB/s) and the
// pair was 36 % of the KFAC pass.  Here the gradient is split ONCE into K-major bf16 hi / lo planes [(r, y', x'), co] -- viewed by
// a 4-D tensor map {co, x', y', r} -- and for every filter tap (a, b) the TMA engine fetches the box of gradient pixels that
// feeds a tile of 128 OUTPUT pixels, shifted by the tap: coordinates that fall outside the gradient map are zero-filled by the
// hardware, which is exactly the convolution's padding.  One accumulator tile D[128 pixels, Cin] sums all kh * kw taps x Cout/32
// chunks in TMEM; the epilogue stores dX directly in its [r, ci, y, x] layout.  Neither the unfolded gradient nor dU ever exists.
#include "tc_common.cuh"
#include "conv_bwd_tc.cuh"

namespace bnnp {
namespace cb {

using namespace bnnp::tcp;

constexpr int BK = 32, TILE_M = 128, MAX_N = 256, STAGES = 6;
constexpr int A_TERM = TILE_M * BK * 2;                 // 8 KB: one bf16 plane of the pixel tile
constexpr int B_TERM = MAX_N * BK * 2;                  // 16 KB (upper bound: Cin <= 256)
constexpr int STAGE_BYTES = 2 * A_TERM + 2 * B_TERM;    // 48 KB
constexpr int SMEM_BYTES = 4 * STAGE_BYTES + 1024 + 256;
constexpr int NUM_THREADS = 256;                        // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-7: epilogue
constexpr int NSTG = 4;

struct Args {
  int R, H, W, Cin, Nt;            // Nt = up16(Cin): MMA N
  int wide;                        // 1 (Nt <= 128): hi and lo weight planes side by side as ONE N = 2 Nt operand
  int kh, kw, pl, pt;
  int BX, BY, BR;                  // box of output pixels: BX * BY * BR = 128
  int tiles_y;                     // ceil(H / BY)
  long long n_tiles;               // tiles_y * ceil(R / BR)
  int n_cc;                        // Cout chunks of 32
  float* dX; long long ld;         // dX[r * ld + ci * H * W + y * W + x]
  const float* bias;               // [Cin] added in the epilogue (forward convolution), or nullptr
};

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_bwd_kernel(const __grid_constant__ CUtensorMap map_gh, const __grid_constant__ CUtensorMap map_gl,
                const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl, Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTG * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + NSTG;
  uint64_t* acc_full = bars + 2 * NSTG;        // [2]
  uint64_t* acc_empty = bars + 2 * NSTG + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTG + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTG; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_taps = args.kh * args.kw;
  const int n_st = n_taps * args.n_cc;
  const uint32_t stage_tx = (uint32_t)(2 * A_TERM + 2 * args.Nt * BK * 2);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (long long tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
        const int ty = (int)(tile % args.tiles_y);
        const int r0 = (int)(tile / args.tiles_y) * args.BR;
        const int y0 = ty * args.BY;
        for (int t = 0; t < n_taps; ++t) {
          const int a = t / args.kw, b = t - a * args.kw;
          const int cx = args.pl - b, cy = y0 + args.pt - a;            // gradient pixel of output pixel (y0, 0) for this tap
          for (int c = 0; c < args.n_cc; ++c) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* base = smem + s * STAGE_BYTES;
            mbar_arrive_expect_tx(&full[s], stage_tx);
            tma_load_4d(&map_gh, &full[s], base, c * BK, cx, cy, r0);
            tma_load_4d(&map_gl, &full[s], base + A_TERM, c * BK, cx, cy, r0);
            tma_load_2d(&map_wh, &full[s], base + 2 * A_TERM, c * BK, t * args.Cin);
            tma_load_2d(&map_wl, &full[s], base + 2 * A_TERM + args.Nt * 64, c * BK, t * args.Cin);
            if (++s == NSTG) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0, it = 0;
      // wide: g_hi x [W_hi; W_lo] is one MMA of N = 2 Nt (the lo plane follows the hi plane in shared memory) into columns
      // [0, Nt) and [Nt, 2 Nt), g_lo x W_hi adds into [0, Nt); the epilogue sums the two ranges.  Two reads of the 128-row
      // gradient tile per k step instead of three: these narrow-N MMAs are bound by their shared-memory operand reads.
      const uint32_t idesc = umma_idesc_bf16(args.Nt), idesc2 = umma_idesc_bf16(2 * args.Nt);
      const uint32_t b_lo = (uint32_t)args.Nt * 64;
      for (long long tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + buf * MAX_N;
        for (int c = 0; c < n_st; ++c) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + 2 * A_TERM;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t dah = umma_desc_sw64(a0 + kk * 32), dal = umma_desc_sw64(a0 + A_TERM + kk * 32);
            const uint64_t dbh = umma_desc_sw64(b0 + kk * 32);
            if (args.wide) {
              umma_bf16(d, dah, dbh, idesc2, (c | kk) ? 1u : 0u);
            } else {
              umma_bf16(d, dah, dbh, idesc, (c | kk) ? 1u : 0u);
              umma_bf16(d, dah, umma_desc_sw64(b0 + b_lo + kk * 32), idesc, 1u);
            }
            umma_bf16(d, dal, dbh, idesc, 1u);
          }
          umma_commit(&empty[s]);
          if (++s == NSTG) { s = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int p = q * 32 + lane;                       // accumulator row = output pixel of the box, x fastest
    const int px = p % args.BX, py = (p / args.BX) % args.BY, pr = p / (args.BX * args.BY);
    const int HW = args.H * args.W;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      const int ty = (int)(tile % args.tiles_y);
      const int r = (int)(tile / args.tiles_y) * args.BR + pr;
      const int y = ty * args.BY + py;
      const bool ok = px < args.W && y < args.H && r < args.R;
      float* out = args.dX + (long long)r * args.ld + (long long)y * args.W + px;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < args.Cin; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * MAX_N + c0), v);
        if (args.wide) {
          uint32_t w[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * MAX_N + args.Nt + c0), w);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(w[e]));
        }
        if (ok) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c0 + e < args.Cin)
              out[(long long)(c0 + e) * HW] = __uint_as_float(v[e]) + (args.bias ? __ldg(args.bias + c0 + e) : 0.f);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// planes[(r * L + pos), co] <- bf16 hi / lo of g[r * gld + co * L + pos]  (co >= Cout zero): 32 x 32 shared-memory transposes
__global__ void split_grad_nhwc_kernel(const float* __restrict__ g, long long gld, int Cout, int L, int Cp,
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
// Same planes, one block per (row r, chunk of PC positions, ALL channels): the [Cout, L] slab of a row is contiguous in g and
// the [L, Cp] slab of the planes is contiguous in the output, so both sides move in long runs (the 32 x 32 tiles above write
// 64-byte segments from 77 000 four-load blocks: 2.4 TB/s); channel pairs leave as bf16x2.  Dynamic smem: Cout * (PC + 1) floats.
__global__ void __launch_bounds__(256)
split_grad_nhwc_slab_kernel(const float* __restrict__ g, long long gld, int Cout, int L, int Cp, int PC,
                            __nv_bfloat162* __restrict__ hi, __nv_bfloat162* __restrict__ lo) {
  extern __shared__ float slab[];
  const long long r = blockIdx.y;
  const int p0 = blockIdx.x * PC, np = min(PC, L - p0), pitch = PC + 1;
  const float* src = g + r * gld + p0;
  for (int i = threadIdx.x; i < Cout * np; i += 256) {
    const int c = i / np, p = i - c * np;
    slab[c * pitch + p] = src[(long long)c * L + p];
  }
  __syncthreads();
  const int hp = Cp >> 1;                                    // channel pairs per position (Cp is a multiple of 8)
  const long long o0 = (r * L + p0) * hp;
  for (int i = threadIdx.x; i < np * hp; i += 256) {
    const int p = i / hp, c = (i - p * hp) * 2;
    const float v0 = c < Cout ? slab[c * pitch + p] : 0.f, v1 = c + 1 < Cout ? slab[(c + 1) * pitch + p] : 0.f;
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    hi[o0 + i] = h;
    lo[o0 + i] = __floats2bfloat162_rn(v0 - __low2float(h), v1 - __high2float(h));
  }
}
// planes[(t * Cin + ci), co] <- bf16 hi / lo of W[co, ci, a, b], t = a * kw + b   (co >= Cout zero)
// fwd: the forward convolution through the same kernel -- the roles of the channel axes swap and the taps are mirrored:
// planes[(t * Cin + ci), co] <- W[ci, co, taps - 1 - t]  (W stored [Cin][Cout][taps] in the kernel's naming)
__global__ void split_weight_taps_kernel(const float* __restrict__ W, int Cout, int Cin, int taps, int Cp, int fwd,
                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= taps * Cin * Cp) return;
  const int co = i % Cp, row = i / Cp, ci = row % Cin, t = row / Cin;
  const float v = co >= Cout ? 0.f
                             : (fwd ? W[((long long)ci * Cout + co) * taps + (taps - 1 - t)] : W[((long long)co * Cin + ci) * taps + t]);
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
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
static int encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* gdim, const cuuint64_t* gstride,
                  const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return BNNP_ERR_CUDA;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[bnnp_b200] cuTensorMapEncodeTiled (conv bwd, rank %d) failed: %d\n", rank, (int)r);
    return BNNP_ERR_CUDA;
  }
  return BNNP_OK;
}

static inline int64_t up8(int64_t v) { return (v + 7) / 8 * 8; }
static inline int64_t up64(int64_t v) { return (v + 63) / 64 * 64; }
static inline int pow2_ge(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace cb

bool conv_bwd_implicit_ok(int Cin, int Cout, int H, int W, int OH, int OW, int kh, int kw, int sh, int sw, int64_t R) {
  return sh == 1 && sw == 1 && kh >= 1 && kw >= 1 && kh * kw <= 32 && Cin >= 1 && Cin <= cb::MAX_N && Cout >= 1 &&
         cb::pow2_ge(W) <= cb::TILE_M && H >= 1 && OH >= 1 && OW >= 1 && OW <= 256 && OH <= 256 && R >= 1 && R <= 65535;
}

int64_t conv_bwd_implicit_scratch_floats(int64_t R, int Cout, int L, int Cin, int taps) {
  return cb::up64(R * L * cb::up8(Cout)) + cb::up64((int64_t)taps * Cin * cb::up8(Cout)) + 256;
}

int conv_bwd_implicit(const float* gout, int64_t gld, int64_t R, int Cin, int Cout, int H, int W, int OH, int OW, int kh, int kw,
                      int pl, int pt, const float* Wt, float* dX, int64_t dld, float* scratch, cudaStream_t st, int fwd,
                      const float* bias) {
  using namespace cb;
  if (!conv_bwd_implicit_ok(Cin, Cout, H, W, OH, OW, kh, kw, 1, 1, R)) return BNNP_ERR_UNSUPPORTED;
  const int L = OH * OW, taps = kh * kw;
  const int Cp = (int)up8(Cout);
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 127) & ~uintptr_t(127));
  __nv_bfloat16* gh = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* gl = gh + R * L * Cp;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(base + up64(R * L * Cp));
  __nv_bfloat16* wl = wh + (int64_t)taps * Cin * Cp;
  if (Cout <= 1024) {
    const int pc_max = 8192 / Cout - 1;                      // slab of at most 32 KB
    const int chunks = (L + pc_max - 1) / pc_max, PC = (L + chunks - 1) / chunks;
    split_grad_nhwc_slab_kernel<<<dim3((unsigned)((L + PC - 1) / PC), (unsigned)R), 256, (size_t)Cout * (PC + 1) * sizeof(float), st>>>(
        gout, gld, Cout, L, Cp, PC, reinterpret_cast<__nv_bfloat162*>(gh), reinterpret_cast<__nv_bfloat162*>(gl));
  } else {
    split_grad_nhwc_kernel<<<dim3((unsigned)ceil_div64(L, 32), (unsigned)ceil_div64(Cp, 32), (unsigned)R), dim3(32, 8), 0, st>>>(
        gout, gld, Cout, L, Cp, gh, gl);
  }
  BNNP_LAUNCH_CHECK();
  split_weight_taps_kernel<<<(unsigned)ceil_div64((int64_t)taps * Cin * Cp, 256), 256, 0, st>>>(Wt, Cout, Cin, taps, Cp, fwd, wh, wl);
  BNNP_LAUNCH_CHECK();

  Args a;
  a.R = (int)R; a.H = H; a.W = W; a.Cin = Cin; a.Nt = (Cin + 15) / 16 * 16;
  a.wide = a.Nt <= 128 ? 1 : 0;
  a.kh = kh; a.kw = kw; a.pl = pl; a.pt = pt;
  a.BX = pow2_ge(W);
  a.BY = pow2_ge(H) < TILE_M / a.BX ? pow2_ge(H) : TILE_M / a.BX;
  a.BR = TILE_M / (a.BX * a.BY);
  a.tiles_y = (H + a.BY - 1) / a.BY;
  a.n_tiles = (long long)a.tiles_y * ((R + a.BR - 1) / a.BR);
  a.n_cc = (Cout + BK - 1) / BK;
  a.dX = dX; a.ld = dld; a.bias = bias;

  CUtensorMap mgh, mgl, mwh, mwl;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cout, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)R};
    cuuint64_t gstr[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)OW * Cp * 2, (cuuint64_t)L * Cp * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)a.BX, (cuuint32_t)a.BY, (cuuint32_t)a.BR};
    int rc;
    if ((rc = encode(&mgh, gh, 4, gdim, gstr, box))) return rc;
    if ((rc = encode(&mgl, gl, 4, gdim, gstr, box))) return rc;
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)Cout, (cuuint64_t)taps * Cin};
    cuuint64_t gstr[1] = {(cuuint64_t)Cp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)a.Nt};
    int rc;
    if ((rc = encode(&mwh, wh, 2, gdim, gstr, box))) return rc;
    if ((rc = encode(&mwl, wl, 2, gdim, gstr, box))) return rc;
  }
  static bool attr = false;
  if (!attr) {
    BNNP_CUDA_CHECK(cudaFuncSetAttribute(conv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  conv_bwd_kernel<<<(int)(a.n_tiles < sms ? a.n_tiles : sms), NUM_THREADS, SMEM_BYTES, st>>>(mgh, mgl, mwh, mwl, a);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

}  // namespace bnnp
