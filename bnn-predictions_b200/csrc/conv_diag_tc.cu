// Diagonal GGN of a conv layer's weight on TMA-fed tensor cores (BackPACK DiagGGNExact, preds/laplace.py:47-59):
//
//     d[co, e] += factor * sum_{r = (n, v)} ( sum_pos g[r, co, pos] * U[n, pos, e] )^2 ,     U = unfold(layer input)
//
// i.e. one small product [Cout x L] . [L x E] per back-propagated row r whose result is squared and summed -- 157 M MACs per
// example on CIFAR10Net, the largest item of the diag pass.  Round 1 ran it on the generated-operand GEMM (operands
// re-synthesised by generator warps for every tile, a read-modify-write of a per-CTA partial in every tile's epilogue:
// 4.8 ms per 512 examples, 35 % of the pass).  Here both operands are plain K-major matrices made ONCE per batch,
//
//     A = g as [(r, co), pos]        bf16 hi / lo, row pitch Lp          (split_conv_grad_rows_kernel)
//     B = U as [(n, e),  pos]        bf16 hi / lo, row pitch Lp          (im2col_rows_split_kernel)
//
// streamed by TMA; a 128-row M tile holds 128 / Cout consecutive rows r of the SAME example (they share the B tile); a CTA
// owns one 128-wide e block and a contiguous slice of the M tiles and keeps the running sum of squares of its slice in
// registers (thread = accumulator row, 128 columns): the epilogue of a tile is tcgen05.ld + 128 FMAs, no memory traffic.
// Each CTA writes its sums once; a second kernel adds the partials in a fixed order (deterministic).
#include "tc_common.cuh"
#include "conv_diag_tc.cuh"

namespace bnnp {
namespace cd {

using namespace bnnp::tcp;

constexpr int BK = 32, TILE = 128, STAGES = 6;
constexpr int TERM = TILE * BK * 2;                    // 8 KB: one bf16 plane of one operand tile
constexpr int STAGE_BYTES = 4 * TERM;                  // A hi, A lo, B hi, B lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int NUM_THREADS = 256;                       // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-7: epilogue

struct Args {
  long long R;            // back-propagated rows (N * V)
  int V, Cout, E, L;
  int ipt;                // rows r per M tile (128 / Cout when that divides V, else 1)
  int n_et;               // 128-wide e blocks
  int cpe;                // CTAs per e block
  long long n_mt;         // M tiles = ceil(R / ipt)
  float* part;            // [n_et * cpe][128][128]
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_diag_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                 const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;        // [4]
  uint64_t* acc_empty = bars + 2 * STAGES + 4;   // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 4; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int et = blockIdx.x % args.n_et;                // this CTA's e block
  const int slot = blockIdx.x / args.n_et;              // and its slice of the M tiles
  const long long t_beg = args.n_mt * slot / args.cpe, t_end = args.n_mt * (slot + 1) / args.cpe;
  const int n_st = (args.L + BK - 1) / BK;
  const int rows_item = args.Cout;
  const long long rows_tile = (long long)args.ipt * rows_item;          // A rows between consecutive M tiles

  if (warp == 0) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      // the gradient planes stream through once, the (small) unfolded input of an example is re-used by its V rows: without
      // the hints the stream evicts it first (ncu on conv1: L2 hit rate 28 %, 1.9 GB of DRAM reads for 1.2 GB of operands)
      const uint64_t pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
      for (long long t = t_beg; t < t_end; ++t) {
        const long long r0 = t * args.ipt;
        const long long n = r0 / args.V;
        for (int c = 0; c < n_st; ++c) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* base = smem + s * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], (uint32_t)STAGE_BYTES);
          if (args.n_et == 1) {           // a single e block: every gradient tile is read exactly once
            tma_load_2d_hint(&map_ah, &full[s], base, c * BK, (int)(t * rows_tile), pol_a);
            tma_load_2d_hint(&map_al, &full[s], base + TERM, c * BK, (int)(t * rows_tile), pol_a);
          } else {                        // the CTAs of the other e blocks read the same tile: leave it to the default policy
            tma_load_2d(&map_ah, &full[s], base, c * BK, (int)(t * rows_tile));
            tma_load_2d(&map_al, &full[s], base + TERM, c * BK, (int)(t * rows_tile));
          }
          tma_load_2d_hint(&map_bh, &full[s], base + 2 * TERM, c * BK, (int)(n * args.E + (long long)et * TILE), pol_b);
          tma_load_2d_hint(&map_bl, &full[s], base + 3 * TERM, c * BK, (int)(n * args.E + (long long)et * TILE), pol_b);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0, it = 0;
      const uint32_t idesc = umma_idesc_bf16(TILE);
      for (long long t = t_beg; t < t_end; ++t, ++it) {
        const uint32_t buf = it & 3;
        mbar_wait(&acc_empty[buf], ((it >> 2) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + buf * TILE;
        for (int c = 0; c < n_st; ++c) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + 2 * TERM;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t dah = umma_desc_sw64(a0 + kk * 32), dal = umma_desc_sw64(a0 + TERM + kk * 32);
            const uint64_t dbh = umma_desc_sw64(b0 + kk * 32), dbl = umma_desc_sw64(b0 + TERM + kk * 32);
            umma_bf16(d, dah, dbh, idesc, (c | kk) ? 1u : 0u);
            umma_bf16(d, dah, dbl, idesc, 1u);
            umma_bf16(d, dal, dbh, idesc, 1u);
          }
          umma_commit(&empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int row = q * 32 + lane;                        // accumulator row = (item within the tile, co)
    const int item = row / rows_item;
    const bool row_ok = row < args.ipt * rows_item;
    float sq[TILE];
#pragma unroll
    for (int e = 0; e < TILE; ++e) sq[e] = 0.f;
    uint32_t it = 0;
    for (long long t = t_beg; t < t_end; ++t, ++it) {
      const uint32_t buf = it & 3;
      mbar_wait(&acc_full[buf], (it >> 2) & 1);
      tc_fence_after();
      const bool ok = row_ok && (t * args.ipt + item) < args.R;          // (the last tile of a ragged R)
#pragma unroll
      for (int c0 = 0; c0 < TILE; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TILE + c0), v);
        if (ok) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float x = __uint_as_float(v[e]);
            sq[c0 + e] = fmaf(x, x, sq[c0 + e]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    // this CTA's sums: part[cta][row][col]   (thread = row: 512-byte runs per thread; written once per launch)
    float* p = args.part + ((long long)blockIdx.x * TILE + row) * TILE;
#pragma unroll
    for (int e = 0; e < TILE; e += 4)
      *reinterpret_cast<float4*>(p + e) = make_float4(sq[e], sq[e + 1], sq[e + 2], sq[e + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// d[co, e] += factor * sum_{slots, items} part[(et, slot)][item * Cout + co][e - et * 128]      (fixed order)
__global__ void conv_diag_reduce_kernel(const float* __restrict__ part, int n_et, int cpe, int ipt, int Cout, int E,
                                        float factor, float* __restrict__ d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * E) return;
  const int co = (int)(i / E), e = (int)(i % E);
  const int et = e / TILE, c = e % TILE;
  float acc = 0.f;
  for (int s = 0; s < cpe; ++s) {
    const float* p = part + ((long long)(s * n_et + et) * TILE) * TILE;
    for (int j = 0; j < ipt; ++j) acc += p[(long long)(j * Cout + co) * TILE + c];
  }
  d[i] = fmaf(factor, acc, d[i]);
}

// planes[(r * Cout + co), pos] <- bf16 hi / lo of g[r * gld + co * L + pos]     (one thread per 8 positions)
__global__ void split_conv_grad_rows_kernel(const float* __restrict__ g, long long gld, long long R, int Cout, int L,
                                            int Lp, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int p8 = (L + 7) >> 3;
  if (i >= R * Cout * p8) return;
  const int pos = (int)(i % p8) << 3;
  const long long rc = i / p8;
  const float* q = g + (rc / Cout) * gld + (rc % Cout) * (long long)L + pos;
  float v[8];
  if (pos + 8 <= L && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(q)), y = __ldg(reinterpret_cast<const float4*>(q) + 1);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = pos + k < L ? __ldg(q + k) : 0.f;
  }
  uint4 h, l;
  split8(v, h, l);
  *reinterpret_cast<uint4*>(hi + rc * Lp + pos) = h;
  *reinterpret_cast<uint4*>(lo + rc * Lp + pos) = l;
}

// planes[(n * E + e), pos] <- bf16 hi / lo of unfold(in[n])[e, pos]        grid (ceil(E * p8 / 256), N)
__global__ void im2col_rows_split_kernel(ConvUnfold u, int L, int E, int Lp, __nv_bfloat16* __restrict__ hi,
                                         __nv_bfloat16* __restrict__ lo) {
  const int p8 = (L + 7) >> 3;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= E * p8) return;
  const long long n = blockIdx.y;
  const int e = j / p8, pos = (j - e * p8) << 3;
  const int b = e % u.kw, t = e / u.kw, a = t % u.kh, ci = t / u.kh;
  int oh = pos / u.OW, ow = pos - oh * u.OW;
  const float* xn = u.in + n * u.ild;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = pos + k < L ? u.at_dec(xn, ci, a, b, oh, ow) : 0.f;
    if (++ow == u.OW) { ow = 0; ++oh; }
  }
  uint4 h, l;
  split8(v, h, l);
  const long long o = (n * E + e) * Lp + pos;
  *reinterpret_cast<uint4*>(hi + o) = h;
  *reinterpret_cast<uint4*>(lo + o) = l;
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
    fprintf(stderr, "[bnnp_b200] cuTensorMapEncodeTiled (conv diag) failed: %d (rows %lld K %lld ld %lld)\n", (int)r, rows, K, ld);
    return BNNP_ERR_CUDA;
  }
  return BNNP_OK;
}

static inline int64_t up64(int64_t v) { return (v + 63) / 64 * 64; }
static inline int pick_ipt(int Cout, int V) {
  if (Cout > TILE) return 0;
  int ipt = TILE / Cout;
  while (ipt > 1 && (V % ipt != 0 || ipt * Cout > TILE)) --ipt;
  return ipt;
}

}  // namespace cd

bool conv_diag_tma_ok(int Cout, int L, long long N, int V) {
  return Cout >= 1 && Cout <= cd::TILE && L >= 1 && N * V * Cout < (1LL << 31) && N <= 65535;   // (N: grid.y of the unfold pass)
}

int64_t conv_diag_tma_scratch_floats(int64_t N, int V, int Cout, int L, int E) {
  const int64_t Lp = cd::up64(L);
  // bf16 planes (two per operand: 1 float holds two elements) + per-CTA partials for up to 160 CTAs
  return cd::up64(N * V * Cout * Lp) + cd::up64(N * E * Lp) + 160LL * cd::TILE * cd::TILE + 256;
}

int conv_diag_tma(const ConvUnfold& u, const float* G, int64_t gld, int64_t N, int V, int Cout, int L, int E, float factor,
                  float* d_w, float* scratch, cudaStream_t st, const __nv_bfloat16* pre_hi, const __nv_bfloat16* pre_lo) {
  using namespace cd;
  const int64_t R = N * V;
  if (!conv_diag_tma_ok(Cout, L, N, V)) return BNNP_ERR_UNSUPPORTED;
  const int Lp = (int)up64(L);
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 127) & ~uintptr_t(127));
  __nv_bfloat16* ah = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* al = ah + R * Cout * Lp;
  __nv_bfloat16* bh = reinterpret_cast<__nv_bfloat16*>(base + up64(R * Cout * Lp));
  __nv_bfloat16* bl = bh + N * E * (int64_t)Lp;
  float* part = base + up64(R * Cout * Lp) + up64(N * E * (int64_t)Lp);
  const int p8 = (L + 7) >> 3;
  if (pre_hi) {            // the planes were made by the producer of the gradient (row pitch L, L % 8 == 0)
    if (!pre_lo || (L & 7)) return BNNP_ERR_INVALID;
  } else {
    split_conv_grad_rows_kernel<<<(unsigned)ceil_div64(R * Cout * p8, 256), 256, 0, st>>>(G, gld, R, Cout, L, Lp, ah, al);
    BNNP_LAUNCH_CHECK();
  }
  im2col_rows_split_kernel<<<dim3((unsigned)ceil_div64((int64_t)E * p8, 256), (unsigned)N), 256, 0, st>>>(u, L, E, Lp, bh, bl);
  BNNP_LAUNCH_CHECK();
  Args a;
  a.R = R; a.V = V; a.Cout = Cout; a.E = E; a.L = L;
  a.ipt = pick_ipt(Cout, V);
  a.n_et = (E + TILE - 1) / TILE;
  a.n_mt = (R + a.ipt - 1) / a.ipt;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if (sms > 160) sms = 160;
  }
  a.cpe = sms / a.n_et;
  if (a.cpe < 1) a.cpe = 1;
  if (a.cpe > a.n_mt) a.cpe = (int)a.n_mt;
  if ((long long)a.cpe * a.n_et > 160) return BNNP_ERR_UNSUPPORTED;
  a.part = part;
  CUtensorMap mah, mal, mbh, mbl;
  int rc;
  if ((rc = make_map(&mah, pre_hi ? pre_hi : ah, R * Cout, L, pre_hi ? L : Lp))) return rc;
  if ((rc = make_map(&mal, pre_hi ? pre_lo : al, R * Cout, L, pre_hi ? L : Lp))) return rc;
  if ((rc = make_map(&mbh, bh, N * E, L, Lp))) return rc;
  if ((rc = make_map(&mbl, bl, N * E, L, Lp))) return rc;
  static bool attr = false;
  if (!attr) {
    BNNP_CUDA_CHECK(cudaFuncSetAttribute(conv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  conv_diag_kernel<<<a.cpe * a.n_et, NUM_THREADS, SMEM_BYTES, st>>>(mah, mal, mbh, mbl, a);
  BNNP_LAUNCH_CHECK();
  conv_diag_reduce_kernel<<<(unsigned)ceil_div64((int64_t)Cout * E, 256), 256, 0, st>>>(part, a.n_et, a.cpe, a.ipt, Cout, E,
                                                                                       factor, d_w);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

}  // namespace bnnp
