// Full-GGN accumulation on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, B operand by
// TMA, A operand synthesised on chip).  Replaces the dominant einsum of preds/laplace.py:42
//     precision += einsum('mpk,mkl,mql->pq', Js, Hess, Js)
// for Linear-only networks without ever forming Js.
//
// Algebra.  For a Linear layer  J_n[k, (i,j)] = g_n[k,i] * a_n[j]  (a_n[j] := 1 for the bias), so
//     H[(i,j),(i',j')] = sum_n  C_n[i,i'] * a_n[j] * a'_n[j'],     C_n = G_n^T Lambda_n G_n  (M x M)
// i.e. for every pair of units (i,i') a *weighted Gram* of the layer inputs.  Contracting the K
// outputs into C_n first removes the factor K from the 2*N*K*P^2 flops of the einsum as written.
//
// Mapping.  One CTA tile = 128 consecutive parameter rows of layer l (weights then biases) x
// two column tiles (units i'_a, i'_b of layer l') x the same <=256-wide range of j'.  Per 32-example
// K-chunk:
//   B  = X^T tile [wc x 32]  bf16 hi/lo, loaded by TMA (SWIZZLE_64B, K-major), shared by both units;
//   A  = c_n[i(p), i'] * x_n[j(p)]  computed in fp32 by 8 generator warps, split into bf16 hi/lo and
//        written to shared memory in the UMMA canonical K-major SWIZZLE_64B layout;
//   D += Ah*Bh + Ah*Bl + Al*Bh   (3 tcgen05.mma kind::f16 per 16-wide K-step; products exact, fp32
//        accumulation in TMEM => ~2^-17 relative per term, "fp32-grade" after summation).
// Epilogue: tcgen05.ld -> registers -> H[p,q] += acc (tiles on/below the diagonal only).
#include <math.h>
#include <stdlib.h>
#include "tc_common.cuh"
#include "gram_tc.cuh"

namespace bnnp {
namespace tc {

constexpr int BK = 32;                 // examples per K-chunk (32 bf16 = 64 B = one SWIZZLE_64B row)
constexpr int TILE_M = 128;
constexpr int MAX_WC = 256;
constexpr int STAGES = 3;
constexpr int A_TERM_BYTES = TILE_M * BK * 2;            // 8 KB
constexpr int A_STAGE_BYTES = 2 * 2 * A_TERM_BYTES;      // 2 accumulators x (hi, lo)
constexpr int B_TERM_BYTES = MAX_WC * BK * 2;            // 16 KB
constexpr int B_STAGE_BYTES = 2 * B_TERM_BYTES;
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;   // 64 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 384;       // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: generators + epilogue
constexpr int GEN_WARP0 = 4;

using namespace bnnp::tcp;

struct BlockArgs {
  // row side (layer l)
  int m, n, has_bias, unit0, col0;     // rows r in [0, m*n (+m)): r < m*n -> (i=r/n, j=r%n), else bias unit r-m*n
  long long row_theta0;                // theta offset of row 0 (weights, biases follow contiguously)
  // column side (layer l')
  int mp, np, unit0p;                  // column tiles: unit i' in [0,mp), j' in [0,np)
  long long col_theta0;
  int wc, n_jt, n_pairs;               // column tile width, #j' tiles, #unit pairs
  int rt0, n_rt;                       // row tiles [rt0, rt0 + n_rt) of this launch (single-CTA kernel)
  int rtp0, n_rtp;                     // row-tile pairs [rtp0, rtp0 + n_rtp) of this launch (CTA-pair kernels)
  int diag_block;                      // 1 when l == l' (skip tiles strictly above the diagonal)
  int n_chunks, chunk0, Mtot;          // K-chunks of this accumulation segment: [chunk0, chunk0 + n_chunks)
  long long Npad, P;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
               const float* __restrict__ XT32, const float* __restrict__ Ct, float* __restrict__ H,
               int* __restrict__ tile_counter, BlockArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_a = bars;                    // [STAGES]  generators -> MMA
  uint64_t* full_b = bars + STAGES;           // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + 2 * STAGES;        // [STAGES]  MMA (tcgen05.commit) -> TMA + generators
  uint64_t* acc_full = bars + 3 * STAGES;     // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);
  int* tile_slot = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_a[s], 8); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles = args.n_pairs * args.n_jt * args.n_rt;
  const int rows_total = args.m * args.n + (args.has_bias ? args.m : 0);
  uint32_t stage = 0, phase = 0;       // smem ring position (every role walks the same sequence)
  uint32_t acc_phase = 0;

  for (;;) {
    if (threadIdx.x == 0) *tile_slot = atomicAdd(tile_counter, 1);
    __syncthreads();
    const int tile = *tile_slot;
    __syncthreads();
    if (tile >= n_tiles) break;
    const int rt = args.rt0 + tile % args.n_rt;
    const int jt = (tile / args.n_rt) % args.n_jt;
    const int pair = tile / (args.n_rt * args.n_jt);
    const int ipa = 2 * pair, ipb = 2 * pair + 1;
    const int nacc = ipb < args.mp ? 2 : 1;
    const int jc0 = jt * args.wc;
    const int row0 = rt * TILE_M;
    if (args.diag_block) {
      // needed iff some (row, col) of the tile is on/below the diagonal of H
      long long max_row = args.row_theta0 + (row0 + TILE_M - 1 < rows_total - 1 ? row0 + TILE_M - 1 : rows_total - 1);
      long long min_col = args.col_theta0 + (long long)ipa * args.np + jc0;
      if (max_row < min_col) continue;          // uniform across the CTA
    }

    if (warp == 0) {
      // ===================== TMA producer: B = X^T rows [jc0, jc0+wc) of layer l', hi and lo =====================
      if (lane == 0) {
        uint32_t s = stage, ph = phase;
        const uint32_t bytes = 2u * (uint32_t)args.wc * BK * 2u;
        for (int c = 0; c < args.n_chunks; ++c) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* bdst = smem + s * STAGE_BYTES + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_b[s], bytes);
          tma_load_2d(&map_bh, &full_b[s], bdst, (args.chunk0 + c) * BK, jc0);
          tma_load_2d(&map_bl, &full_b[s], bdst + B_TERM_BYTES, (args.chunk0 + c) * BK, jc0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        uint32_t s = stage, ph = phase;
        // the ragged last column tile issues a narrower MMA (the TMA box keeps its full height; rows past the
        // layer's fan-in arrive as zeros and are simply not multiplied)
        const int wvalid = args.np - jc0 < args.wc ? args.np - jc0 : args.wc;
        const uint32_t idesc = umma_idesc_bf16((wvalid + 15) / 16 * 16);
        for (int c = 0; c < args.n_chunks; ++c) {
          mbar_wait(&full_a[s], ph);
          mbar_wait(&full_b[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b0 = a0 + A_STAGE_BYTES;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            if (a < nacc) {
              const uint32_t ah = a0 + a * 2 * A_TERM_BYTES, al = ah + A_TERM_BYTES;
              const uint32_t d = tmem_base + a * MAX_WC;
#pragma unroll
              for (int ks = 0; ks < BK / 16; ++ks) {
                const uint64_t dah = umma_desc_sw64(ah + ks * 32), dal = umma_desc_sw64(al + ks * 32);
                const uint64_t dbh = umma_desc_sw64(b0 + ks * 32), dbl = umma_desc_sw64(b0 + B_TERM_BYTES + ks * 32);
                umma_bf16(d, dah, dbh, idesc, (c | ks) ? 1u : 0u);
                umma_bf16(d, dah, dbl, idesc, 1u);
                umma_bf16(d, dal, dbh, idesc, 1u);
              }
            }
          }
          umma_commit(&empty[s]);                 // frees the smem stage when these MMAs retire
          if (c == args.n_chunks - 1) umma_commit(acc_full);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp >= GEN_WARP0) {
      // ===================== A generators (then epilogue) =====================
      const int gw = warp - GEN_WARP0;
      const int g = lane & 3;                  // 16-byte K-group (8 examples) inside the 32-example chunk
      // this thread's two rows: gw*16 + it*8 + lane/4
      const float* xsrc[2];
      const float* csrc[2][2];
      bool rvalid[2];
      uint32_t soff[2];
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int rl = gw * 16 + it * 8 + (lane >> 2);
        const int r = row0 + rl;
        rvalid[it] = r < rows_total;
        int u = args.unit0, src = args.col0;
        if (rvalid[it]) {
          if (r < args.m * args.n) { u += r / args.n; src += r % args.n; }
          else { u += r - args.m * args.n; src += args.n; }        // bias row: the constant-1 column
        }
        xsrc[it] = XT32 + (long long)src * args.Npad + (long long)args.chunk0 * BK + g * 8;
        csrc[0][it] = Ct + ((long long)(args.unit0p + ipa) * args.Mtot + u) * args.Npad + (long long)args.chunk0 * BK + g * 8;
        csrc[1][it] = Ct + ((long long)(args.unit0p + (nacc > 1 ? ipb : ipa)) * args.Mtot + u) * args.Npad + (long long)args.chunk0 * BK + g * 8;
        soff[it] = (uint32_t)((rl >> 3) * 512 + (rl & 7) * 64 + ((g ^ ((rl >> 1) & 3)) << 4));
      }
      uint32_t s = stage, ph = phase;
      const int pf_chunk = args.n_chunks > 48 ? args.n_chunks - 40 : 0;
      float4 xv[2][2], cv[2][2][2];
      auto load_chunk = [&](int c) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const float4* xp = reinterpret_cast<const float4*>(xsrc[it] + c * BK);
          xv[it][0] = __ldg(xp); xv[it][1] = __ldg(xp + 1);
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const float4* cp = reinterpret_cast<const float4*>(csrc[a][it] + c * BK);
            cv[a][it][0] = __ldg(cp); cv[a][it][1] = __ldg(cp + 1);
          }
        }
      };
      load_chunk(0);
      for (int c = 0; c < args.n_chunks; ++c) {
        float4 x[2][2], cc[2][2][2];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          x[it][0] = xv[it][0]; x[it][1] = xv[it][1];
#pragma unroll
          for (int a = 0; a < 2; ++a) { cc[a][it][0] = cv[a][it][0]; cc[a][it][1] = cv[a][it][1]; }
        }
        if (c + 1 < args.n_chunks) load_chunk(c + 1);      // software prefetch of the next chunk
        if (c == pf_chunk) {
          // pull this tile of H towards L2 so that the read-modify-write epilogue does not pay DRAM latency
          const int ncols_pf = args.np - jc0 < args.wc ? args.np - jc0 : args.wc;
          const int lines = (ncols_pf * 4 + 127) / 128 + 1;
          for (int a = 0; a < nacc; ++a) {
            const int ip = a == 0 ? ipa : ipb;
            for (int w = (warp - GEN_WARP0) * 32 + lane; w < TILE_M * lines; w += 256) {
              const int rr = w / lines, ln = w % lines;
              if (row0 + rr < rows_total) {
                const float* ptr = H + (args.row_theta0 + row0 + rr) * args.P + args.col_theta0 +
                                   (long long)ip * args.np + jc0 + ln * 32;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
              }
            }
          }
        }
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* abase = smem + s * STAGE_BYTES;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          if (a < nacc) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
              float v[8];
              const float* xf = reinterpret_cast<const float*>(&x[it][0]);
              const float* cf = reinterpret_cast<const float*>(&cc[a][it][0]);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = rvalid[it] ? xf[e] * cf[e] : 0.f;
              uint4 hi, lo;
              split8(v, hi, lo);
              uint8_t* dst = abase + a * 2 * A_TERM_BYTES + soff[it];
              *reinterpret_cast<uint4*>(dst) = hi;
              *reinterpret_cast<uint4*>(dst + A_TERM_BYTES) = lo;
            }
          }
        }
        fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      // ---------------- epilogue: H[p, q] += D ----------------
      // TMEM -> registers (thread = row) -> per-warp smem transpose -> coalesced read-modify-write of
      // H (lane = column, 128 B per request, 8 rows in flight).  All MMAs of the tile have retired, so
      // the stage-0 operand buffers are free to serve as the transpose scratch.
      mbar_wait(acc_full, acc_phase);
      tc_fence_after();
      const int a = gw >> 2, quarter = gw & 3;
      if (a < nacc) {
        float* sc = reinterpret_cast<float*>(smem + gw * (32 * 33 * 4 + 128));     // [32][33]
        const int ip = a == 0 ? ipa : ipb;
        const int ncols = args.np - jc0 < args.wc ? args.np - jc0 : args.wc;
        const int rbase = row0 + quarter * 32;
        float* hbase = H + (args.row_theta0 + rbase) * args.P + args.col_theta0 + (long long)ip * args.np + jc0;
        int nrows = rows_total - rbase;
        nrows = nrows < 0 ? 0 : (nrows > 32 ? 32 : nrows);
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * MAX_WC + c0), v);
          __syncwarp();
#pragma unroll
          for (int e = 0; e < 32; ++e) sc[lane * 33 + e] = __uint_as_float(v[e]);
          __syncwarp();
          const bool colok = c0 + lane < ncols;
          {
            // fire-and-forget reductions (RED.ADD.F32, 128 B per warp request): the add happens in L2, the epilogue
            // never waits for H.  Every element of H receives exactly one add per launch, so the result does not
            // depend on the order in which the reductions land.
#pragma unroll
            for (int u = 0; u < 32; ++u)
              if (colok && u < nrows) atomicAdd(hbase + (long long)u * args.P + c0 + lane, sc[u * 33 + lane]);
          }
        }
      }
      tc_fence_before();
    }
    // every role advanced the ring by n_chunks stages
    {
      uint32_t adv = (uint32_t)args.n_chunks;
      uint32_t tot = stage + adv;
      phase ^= (tot / STAGES) & 1;
      stage = tot % STAGES;
      acc_phase ^= 1;
    }
    tc_fence_before();
    __syncthreads();            // TMEM drained by the epilogue before the next tile's first MMA overwrites it
    tc_fence_after();
  }

  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---- CTA-pair version -------------------------------------------------------------------------------------------
// gram_tc_kernel is bound by shared-memory traffic: per 32-example chunk the generators store 32 KB of A, TMA stores
// 32 KB of B and the tensor core reads 144 KB of operands (measured: 43 ms per launch, 36 without the A stores, 32
// with neither stores nor H epilogue).  Here two CTAs of a cluster (one TPC) run ONE M = 256 `tcgen05.mma.cta_group::2`
// over two consecutive 128-row tiles: each CTA generates the A rows of its own tile and stages only HALF of the B tile
// (the instruction reads the other half from the peer), which also leaves room for a fourth pipeline stage.
// Rank 0 fetches the tile index for both CTAs, issues the MMAs and commits to the barriers of both; the generators of
// rank 1 arrive on rank 0's `full_a` barrier, every TMA load credits rank 0's `full_b`.
#ifdef BNNP_EXPERIMENTS      // kept for A/B timing against gram_wide_kernel (build with BNNP_EXPERIMENTS=1)
constexpr int P_STAGES = 4;
constexpr int P_B_TERM_BYTES = (MAX_WC / 2) * BK * 2;                      // 8 KB
constexpr int P_STAGE_BYTES = A_STAGE_BYTES + 2 * P_B_TERM_BYTES;          // 48 KB
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gram_pair_kernel(const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                 const float* __restrict__ XT32, const float* __restrict__ Ct, float* __restrict__ H,
                 int* __restrict__ tile_counter, BlockArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* full_a = bars;                      // [P_STAGES] rank 0 only: generator warps of both CTAs
  uint64_t* full_b = bars + P_STAGES;           // [P_STAGES] rank 0 only: TMA bytes of both CTAs
  uint64_t* empty = bars + 2 * P_STAGES;        // [P_STAGES] both CTAs (multicast commit)
  uint64_t* acc_full = bars + 3 * P_STAGES;     // both CTAs (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * P_STAGES + 1);
  int* tile_slot = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full_a[s], 16); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_rtp = args.n_rtp;
  const int n_tiles = args.n_pairs * args.n_jt * n_rtp;
  const int rows_total = args.m * args.n + (args.has_bias ? args.m : 0);
  uint32_t stage = 0, phase = 0;
  uint32_t acc_phase = 0;

  for (;;) {
    if (threadIdx.x == 0 && rank == 0) {
      const int t = atomicAdd(tile_counter, 1);
      *tile_slot = t;
      st_shared_cluster_u32(mapa_shared(smem_u32(tile_slot), 1), (uint32_t)t);
    }
    cluster_sync_all();                         // release/acquire at cluster scope: the slot is visible in both CTAs
    const int tile = *tile_slot;
    if (tile >= n_tiles) break;
    const int rtp = args.rtp0 + tile % n_rtp;
    const int jt = (tile / n_rtp) % args.n_jt;
    const int pair = tile / (n_rtp * args.n_jt);
    const int ipa = 2 * pair, ipb = 2 * pair + 1;
    const int nacc = ipb < args.mp ? 2 : 1;
    const int jc0 = jt * args.wc;
    const int row0 = (2 * rtp + (int)rank) * TILE_M;
    bool own_needed = row0 < rows_total;
    bool pair_needed = true;
    if (args.diag_block) {
      const long long min_col = args.col_theta0 + (long long)ipa * args.np + jc0;
      const int prow_last = 2 * rtp * TILE_M + 2 * TILE_M - 1;
      const long long pair_max_row = args.row_theta0 + (prow_last < rows_total - 1 ? prow_last : rows_total - 1);
      pair_needed = pair_max_row >= min_col;
      const long long own_max_row = args.row_theta0 + (row0 + TILE_M - 1 < rows_total - 1 ? row0 + TILE_M - 1 : rows_total - 1);
      own_needed = own_needed && own_max_row >= min_col;
    }
    if (pair_needed) {
      const int wvalid = args.np - jc0 < args.wc ? args.np - jc0 : args.wc;
      const int nmma = (wvalid + 15) / 16 * 16;                 // instruction N; each CTA supplies nmma/2 rows of B
      if (warp == 0) {
        // ===================== TMA producer: this CTA's half of B =====================
        if (lane == 0) {
          uint32_t s = stage, ph = phase;
          const uint32_t bytes = 2u * 2u * (uint32_t)(args.wc / 2) * BK * 2u;        // both CTAs, hi + lo
          const int brow = jc0 + (int)rank * (nmma / 2);
          for (int c = 0; c < args.n_chunks; ++c) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* bdst = smem + s * P_STAGE_BYTES + A_STAGE_BYTES;
            const uint32_t lead_full = mapa_shared(smem_u32(&full_b[s]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_b[s], bytes);
            tma_load_2d_pair(&map_bh, lead_full, bdst, (args.chunk0 + c) * BK, brow);
            tma_load_2d_pair(&map_bl, lead_full, bdst + P_B_TERM_BYTES, (args.chunk0 + c) * BK, brow);
            if (++s == P_STAGES) { s = 0; ph ^= 1; }
          }
        }
      } else if (warp == 1) {
        // ===================== MMA issuer (rank 0) =====================
        if (lane == 0 && rank == 0) {
          uint32_t s = stage, ph = phase;
          const uint32_t idesc = umma_idesc_bf16_pair(nmma);
          for (int c = 0; c < args.n_chunks; ++c) {
            mbar_wait_cluster(&full_a[s], ph);
            mbar_wait(&full_b[s], ph);
            tc_fence_after();
            const uint32_t a0 = smem_u32(smem + s * P_STAGE_BYTES);
            const uint32_t b0 = a0 + A_STAGE_BYTES;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              if (a < nacc) {
                const uint32_t ah = a0 + a * 2 * A_TERM_BYTES, al = ah + A_TERM_BYTES;
                const uint32_t d = tmem_base + a * MAX_WC;
#pragma unroll
                for (int ks = 0; ks < BK / 16; ++ks) {
                  const uint64_t dah = umma_desc_sw64(ah + ks * 32), dal = umma_desc_sw64(al + ks * 32);
                  const uint64_t dbh = umma_desc_sw64(b0 + ks * 32), dbl = umma_desc_sw64(b0 + P_B_TERM_BYTES + ks * 32);
                  umma_bf16_pair(d, dah, dbh, idesc, (c | ks) ? 1u : 0u);
                  umma_bf16_pair(d, dah, dbl, idesc, 1u);
                  umma_bf16_pair(d, dal, dbh, idesc, 1u);
                }
              }
            }
            umma_commit_pair(&empty[s]);
            if (c == args.n_chunks - 1) umma_commit_pair(acc_full);
            if (++s == P_STAGES) { s = 0; ph ^= 1; }
          }
        }
      } else if (warp >= GEN_WARP0) {
        // ===================== A generators (then epilogue), rows of this CTA's own tile =====================
        const int gw = warp - GEN_WARP0;
        const int g = lane & 3;
        const float* xsrc[2];
        const float* csrc[2][2];
        bool rvalid[2];
        uint32_t soff[2];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int rl = gw * 16 + it * 8 + (lane >> 2);
          const int r = row0 + rl;
          rvalid[it] = r < rows_total;
          int u = args.unit0, src = args.col0;
          if (rvalid[it]) {
            if (r < args.m * args.n) { u += r / args.n; src += r % args.n; }
            else { u += r - args.m * args.n; src += args.n; }
          }
          xsrc[it] = XT32 + (long long)src * args.Npad + (long long)args.chunk0 * BK + g * 8;
          csrc[0][it] = Ct + ((long long)(args.unit0p + ipa) * args.Mtot + u) * args.Npad + (long long)args.chunk0 * BK + g * 8;
          csrc[1][it] = Ct + ((long long)(args.unit0p + (nacc > 1 ? ipb : ipa)) * args.Mtot + u) * args.Npad + (long long)args.chunk0 * BK + g * 8;
          soff[it] = (uint32_t)((rl >> 3) * 512 + (rl & 7) * 64 + ((g ^ ((rl >> 1) & 3)) << 4));
        }
        uint32_t s = stage, ph = phase;
        const int pf_chunk = args.n_chunks > 48 ? args.n_chunks - 40 : 0;
        float4 xv[2][2], cv[2][2][2];
        auto load_chunk = [&](int c) {
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const float4* xp = reinterpret_cast<const float4*>(xsrc[it] + c * BK);
            xv[it][0] = __ldg(xp); xv[it][1] = __ldg(xp + 1);
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              const float4* cp = reinterpret_cast<const float4*>(csrc[a][it] + c * BK);
              cv[a][it][0] = __ldg(cp); cv[a][it][1] = __ldg(cp + 1);
            }
          }
        };
        load_chunk(0);
        for (int c = 0; c < args.n_chunks; ++c) {
          float4 x[2][2], cc[2][2][2];
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            x[it][0] = xv[it][0]; x[it][1] = xv[it][1];
#pragma unroll
            for (int a = 0; a < 2; ++a) { cc[a][it][0] = cv[a][it][0]; cc[a][it][1] = cv[a][it][1]; }
          }
          if (c + 1 < args.n_chunks) load_chunk(c + 1);
          if (c == pf_chunk && own_needed) {
            const int lines = (wvalid * 4 + 127) / 128 + 1;
            for (int a = 0; a < nacc; ++a) {
              const int ip = a == 0 ? ipa : ipb;
              for (int w = (warp - GEN_WARP0) * 32 + lane; w < TILE_M * lines; w += 256) {
                const int rr = w / lines, ln = w % lines;
                if (row0 + rr < rows_total) {
                  const float* ptr = H + (args.row_theta0 + row0 + rr) * args.P + args.col_theta0 +
                                     (long long)ip * args.np + jc0 + ln * 32;
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
                }
              }
            }
          }
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* abase = smem + s * P_STAGE_BYTES;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            if (a < nacc) {
#pragma unroll
              for (int it = 0; it < 2; ++it) {
                float v[8];
                const float* xf = reinterpret_cast<const float*>(&x[it][0]);
                const float* cf = reinterpret_cast<const float*>(&cc[a][it][0]);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = rvalid[it] ? xf[e] * cf[e] : 0.f;
                uint4 hi, lo;
                split8(v, hi, lo);
                uint8_t* dst = abase + a * 2 * A_TERM_BYTES + soff[it];
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + A_TERM_BYTES) = lo;
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&full_a[s]);
            else mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&full_a[s]), 0));
          }
          if (++s == P_STAGES) { s = 0; ph ^= 1; }
        }
        // ---------------- epilogue: H[p, q] += D for this CTA's rows ----------------
        mbar_wait(acc_full, acc_phase);
        tc_fence_after();
        const int a = gw >> 2, quarter = gw & 3;
        if (a < nacc && own_needed) {
          float* sc = reinterpret_cast<float*>(smem + gw * (32 * 33 * 4 + 128));
          const int ip = a == 0 ? ipa : ipb;
          const int rbase = row0 + quarter * 32;
          float* hbase = H + (args.row_theta0 + rbase) * args.P + args.col_theta0 + (long long)ip * args.np + jc0;
          int nrows = rows_total - rbase;
          nrows = nrows < 0 ? 0 : (nrows > 32 ? 32 : nrows);
          for (int c0 = 0; c0 < wvalid; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * MAX_WC + c0), v);
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 32; ++e) sc[lane * 33 + e] = __uint_as_float(v[e]);
            __syncwarp();
            const bool colok = c0 + lane < wvalid;
#pragma unroll
            for (int u = 0; u < 32; ++u)
              if (colok && u < nrows) atomicAdd(hbase + (long long)u * args.P + c0 + lane, sc[u * 33 + lane]);
          }
        }
        tc_fence_before();
      }
      {
        const uint32_t tot = stage + (uint32_t)args.n_chunks;
        phase ^= (tot / P_STAGES) & 1;
        stage = tot % P_STAGES;
        acc_phase ^= 1;
      }
    }
    tc_fence_before();
    cluster_sync_all();         // both CTAs drained their TMEM rows (and read the tile slot) before the next tile
    tc_fence_after();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}
#endif  // BNNP_EXPERIMENTS

// ---- CTA-pair version, ONE column unit x TWO column tiles ("wide") -------------------------------------------------
// In gram_pair_kernel the two accumulators belong to two column units i'_a, i'_b: they share the TMA-fed B tile but each
// needs its own generated A = c_n[i(p), i'] * x_n[j(p)] (32 KB of generator stores per 32-example chunk).  Here both
// accumulators belong to the SAME unit i' and cover two adjacent j' ranges [jc0, jc0+wc), [jc0+wc, jc0+2wc): they share
// the generated A (16 KB per chunk: half the generator loads, multiplies, bf16 splits and shared-memory stores, each
// generated element now feeds up to 448 output columns) and take two B tiles, which the TMA engine delivers without any
// thread work.  The MMA sequence per accumulator is unchanged, so H stays bit-identical to the other kernels.  A second
// column tile that lies entirely above the diagonal of H is skipped on its own (finer than whole-tile skipping).
// `args.rtp0 / args.n_rtp` restrict a launch to a range of row-tile pairs (row-chunked accumulation, so that finished
// row chunks of H can be exchanged between GPUs while later chunks still accumulate).
constexpr int W_STAGES = 4;
constexpr int W_A_STAGE_BYTES = 2 * A_TERM_BYTES;                     // hi + lo: 16 KB
constexpr int W_B_TERM_BYTES = (MAX_WC / 2) * BK * 2;                 // this CTA's half of one column tile, one term: 8 KB
constexpr int W_B_STAGE_BYTES = 2 * 2 * W_B_TERM_BYTES;               // 2 column tiles x (hi, lo): 32 KB
constexpr int W_STAGE_BYTES = W_A_STAGE_BYTES + W_B_STAGE_BYTES;      // 48 KB
constexpr int W_SMEM_BYTES = W_STAGES * W_STAGE_BYTES + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gram_wide_kernel(const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                 const float* __restrict__ XT32, const float* __restrict__ Ct, float* __restrict__ H,
                 int* __restrict__ tile_counter, BlockArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W_STAGES * W_STAGE_BYTES);
  uint64_t* full_a = bars;                      // [W_STAGES] rank 0 only: generator warps of both CTAs
  uint64_t* full_b = bars + W_STAGES;           // [W_STAGES] rank 0 only: TMA bytes of both CTAs
  uint64_t* empty = bars + 2 * W_STAGES;        // [W_STAGES] both CTAs (multicast commit)
  uint64_t* acc_full = bars + 3 * W_STAGES;     // both CTAs (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * W_STAGES + 1);
  int* tile_slot = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int s = 0; s < W_STAGES; ++s) { mbar_init(&full_a[s], 16); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_rtp = args.n_rtp;
  const int n_jtp = (args.n_jt + 1) / 2;
  const int n_tiles = args.mp * n_jtp * n_rtp;
  const int rows_total = args.m * args.n + (args.has_bias ? args.m : 0);
  uint32_t stage = 0, phase = 0;
  uint32_t acc_phase = 0;

  for (;;) {
    if (threadIdx.x == 0 && rank == 0) {
      const int t = atomicAdd(tile_counter, 1);
      *tile_slot = t;
      st_shared_cluster_u32(mapa_shared(smem_u32(tile_slot), 1), (uint32_t)t);
    }
    cluster_sync_all();                         // release/acquire at cluster scope: the slot is visible in both CTAs
    const int tile = *tile_slot;
    if (tile >= n_tiles) break;
    const int rtp = args.rtp0 + tile % n_rtp;   // row-tile pair fastest: concurrent CTAs share c_n rows and B tiles in L2
    const int jtp = (tile / n_rtp) % n_jtp;
    const int ip = tile / (n_rtp * n_jtp);
    const int jc0 = jtp * 2 * args.wc;
    int nacc = jc0 + args.wc < args.np ? 2 : 1;
    const int row0 = (2 * rtp + (int)rank) * TILE_M;
    bool own_needed = row0 < rows_total;
    bool pair_needed = true;
    bool own_acc1 = true;                        // this CTA's rows reach the second column tile
    if (args.diag_block) {
      const long long min_col = args.col_theta0 + (long long)ip * args.np + jc0;
      const int prow_last = 2 * rtp * TILE_M + 2 * TILE_M - 1;
      const long long pair_max_row = args.row_theta0 + (prow_last < rows_total - 1 ? prow_last : rows_total - 1);
      pair_needed = pair_max_row >= min_col;
      if (nacc == 2 && pair_max_row < min_col + args.wc) nacc = 1;
      const long long own_max_row = args.row_theta0 + (row0 + TILE_M - 1 < rows_total - 1 ? row0 + TILE_M - 1 : rows_total - 1);
      own_needed = own_needed && own_max_row >= min_col;
      own_acc1 = own_max_row >= min_col + args.wc;
    }
    if (pair_needed) {
      int wvalid[2], nmma[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int left = args.np - (jc0 + a * args.wc);
        wvalid[a] = left < args.wc ? (left > 0 ? left : 0) : args.wc;
        nmma[a] = (wvalid[a] + 15) / 16 * 16;                  // instruction N; each CTA supplies nmma/2 rows of B
      }
      if (warp == 0) {
        // ===================== TMA producer: this CTA's half of each column tile, hi and lo =====================
        if (lane == 0) {
          uint32_t s = stage, ph = phase;
          const uint32_t bytes = 2u * (uint32_t)nacc * 2u * (uint32_t)(args.wc / 2) * BK * 2u;    // both CTAs
          for (int c = 0; c < args.n_chunks; ++c) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* bdst = smem + s * W_STAGE_BYTES + W_A_STAGE_BYTES;
            const uint32_t lead_full = mapa_shared(smem_u32(&full_b[s]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_b[s], bytes);
            for (int a = 0; a < nacc; ++a) {
              const int brow = jc0 + a * args.wc + (int)rank * (nmma[a] / 2);
              tma_load_2d_pair(&map_bh, lead_full, bdst + a * 2 * W_B_TERM_BYTES, (args.chunk0 + c) * BK, brow);
              tma_load_2d_pair(&map_bl, lead_full, bdst + a * 2 * W_B_TERM_BYTES + W_B_TERM_BYTES, (args.chunk0 + c) * BK, brow);
            }
            if (++s == W_STAGES) { s = 0; ph ^= 1; }
          }
        }
      } else if (warp == 1) {
        // ===================== MMA issuer (rank 0) =====================
        if (lane == 0 && rank == 0) {
          uint32_t s = stage, ph = phase;
          const uint32_t idesc0 = umma_idesc_bf16_pair(nmma[0]), idesc1 = umma_idesc_bf16_pair(nmma[1] ? nmma[1] : 16);
          for (int c = 0; c < args.n_chunks; ++c) {
            mbar_wait_cluster(&full_a[s], ph);
            mbar_wait(&full_b[s], ph);
            tc_fence_after();
            const uint32_t ah = smem_u32(smem + s * W_STAGE_BYTES), al = ah + A_TERM_BYTES;
            const uint32_t b0 = ah + W_A_STAGE_BYTES;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              if (a < nacc) {
                const uint32_t bh = b0 + a * 2 * W_B_TERM_BYTES, bl = bh + W_B_TERM_BYTES;
                const uint32_t d = tmem_base + a * MAX_WC;
                const uint32_t idesc = a ? idesc1 : idesc0;
#pragma unroll
                for (int ks = 0; ks < BK / 16; ++ks) {
                  const uint64_t dah = umma_desc_sw64(ah + ks * 32), dal = umma_desc_sw64(al + ks * 32);
                  const uint64_t dbh = umma_desc_sw64(bh + ks * 32), dbl = umma_desc_sw64(bl + ks * 32);
                  umma_bf16_pair(d, dah, dbh, idesc, (c | ks) ? 1u : 0u);
                  umma_bf16_pair(d, dah, dbl, idesc, 1u);
                  umma_bf16_pair(d, dal, dbh, idesc, 1u);
                }
              }
            }
            umma_commit_pair(&empty[s]);
            if (c == args.n_chunks - 1) umma_commit_pair(acc_full);
            if (++s == W_STAGES) { s = 0; ph ^= 1; }
          }
        }
      } else if (warp >= GEN_WARP0) {
        // ===================== A generators (then epilogue), rows of this CTA's own tile =====================
        const int gw = warp - GEN_WARP0;
        const int g = lane & 3;
        const float* xsrc[2];
        const float* csrc[2];
        bool rvalid[2];
        uint32_t soff[2];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int rl = gw * 16 + it * 8 + (lane >> 2);
          const int r = row0 + rl;
          rvalid[it] = r < rows_total;
          int u = args.unit0, src = args.col0;
          if (rvalid[it]) {
            if (r < args.m * args.n) { u += r / args.n; src += r % args.n; }
            else { u += r - args.m * args.n; src += args.n; }
          }
          xsrc[it] = XT32 + (long long)src * args.Npad + (long long)args.chunk0 * BK + g * 8;
          csrc[it] = Ct + ((long long)(args.unit0p + ip) * args.Mtot + u) * args.Npad + (long long)args.chunk0 * BK + g * 8;
          soff[it] = (uint32_t)((rl >> 3) * 512 + (rl & 7) * 64 + ((g ^ ((rl >> 1) & 3)) << 4));
        }
        uint32_t s = stage, ph = phase;
        const int pf_chunk = args.n_chunks > 48 ? args.n_chunks - 40 : 0;
        float4 xv[2][2], cv[2][2];
        auto load_chunk = [&](int c) {
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const float4* xp = reinterpret_cast<const float4*>(xsrc[it] + c * BK);
            xv[it][0] = __ldg(xp); xv[it][1] = __ldg(xp + 1);
            const float4* cp = reinterpret_cast<const float4*>(csrc[it] + c * BK);
            cv[it][0] = __ldg(cp); cv[it][1] = __ldg(cp + 1);
          }
        };
        load_chunk(0);
        for (int c = 0; c < args.n_chunks; ++c) {
          float4 x[2][2], cc[2][2];
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            x[it][0] = xv[it][0]; x[it][1] = xv[it][1];
            cc[it][0] = cv[it][0]; cc[it][1] = cv[it][1];
          }
          if (c + 1 < args.n_chunks) load_chunk(c + 1);
          if (c == pf_chunk && own_needed) {
            // pull this tile of H towards L2 so that the reductions of the epilogue do not pay DRAM latency
            for (int a = 0; a < nacc; ++a) {
              if (a == 1 && !own_acc1) break;
              const int lines = (wvalid[a] * 4 + 127) / 128 + 1;
              for (int w = (warp - GEN_WARP0) * 32 + lane; w < TILE_M * lines; w += 256) {
                const int rr = w / lines, ln = w % lines;
                if (row0 + rr < rows_total) {
                  const float* ptr = H + (args.row_theta0 + row0 + rr) * args.P + args.col_theta0 +
                                     (long long)ip * args.np + jc0 + a * args.wc + ln * 32;
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
                }
              }
            }
          }
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* abase = smem + s * W_STAGE_BYTES;
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            float v[8];
            const float* xf = reinterpret_cast<const float*>(&x[it][0]);
            const float* cf = reinterpret_cast<const float*>(&cc[it][0]);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = rvalid[it] ? xf[e] * cf[e] : 0.f;
            uint4 hi, lo;
            split8(v, hi, lo);
            uint8_t* dst = abase + soff[it];
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + A_TERM_BYTES) = lo;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&full_a[s]);
            else mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&full_a[s]), 0));
          }
          if (++s == W_STAGES) { s = 0; ph ^= 1; }
        }
        // ---------------- epilogue: H[p, q] += D for this CTA's rows ----------------
        // two groups of four warps (one warp per 32-lane TMEM quarter): with two column tiles each group drains one
        // accumulator, with one they split its 32-column steps
        mbar_wait(acc_full, acc_phase);
        tc_fence_after();
        const int grp = gw >> 2, quarter = gw & 3;
        if (own_needed) {
          float* sc = reinterpret_cast<float*>(smem + gw * (32 * 33 * 4 + 128));
          const int a = nacc == 2 ? grp : 0;
          const int nsteps = (wvalid[a] + 31) / 32;
          const int s_beg = nacc == 2 ? 0 : (grp == 0 ? 0 : (nsteps + 1) / 2);
          const int s_end = nacc == 2 ? nsteps : (grp == 0 ? (nsteps + 1) / 2 : nsteps);
          const int rbase = row0 + quarter * 32;
          float* hbase = H + (args.row_theta0 + rbase) * args.P + args.col_theta0 + (long long)ip * args.np + jc0 + a * args.wc;
          int nrows = rows_total - rbase;
          nrows = nrows < 0 ? 0 : (nrows > 32 ? 32 : nrows);
          if (!(a == 1 && !own_acc1)) {
            for (int st = s_beg; st < s_end; ++st) {
              const int c0 = st * 32;
              uint32_t v[32];
              tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * MAX_WC + c0), v);
              __syncwarp();
#pragma unroll
              for (int e = 0; e < 32; ++e) sc[lane * 33 + e] = __uint_as_float(v[e]);
              __syncwarp();
              const bool colok = c0 + lane < wvalid[a];
              // fire-and-forget reductions (RED.ADD.F32, 128 B per warp request): the add happens in L2; every element
              // of H receives exactly one add per launch, so the result does not depend on the order they land in
#pragma unroll
              for (int u = 0; u < 32; ++u)
                if (colok && u < nrows) atomicAdd(hbase + (long long)u * args.P + c0 + lane, sc[u * 33 + lane]);
            }
          }
        }
        tc_fence_before();
      }
      {
        const uint32_t tot = stage + (uint32_t)args.n_chunks;
        phase ^= (tot / W_STAGES) & 1;
        stage = tot % W_STAGES;
        acc_phase ^= 1;
      }
    }
    tc_fence_before();
    cluster_sync_all();         // both CTAs drained their TMEM rows (and read the tile slot) before the next tile
    tc_fence_after();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

// ---------------------------------------------------------------- pre-kernels
// XT32[c, n] = A[n, c];  XTh/XTl = bf16 hi/lo split;  zero for n >= N   (32x32 smem transpose)
__global__ void transpose_split_kernel(const float* __restrict__ A, long long ald, long long N, int D,
                                       long long Npad, float* __restrict__ XT32,
                                       __nv_bfloat16* __restrict__ XTh, __nv_bfloat16* __restrict__ XTl) {
  __shared__ float tile[32][33];
  const long long n0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    long long n = n0 + i; int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (n < N && c < D) ? A[n * ald + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i; long long n = n0 + threadIdx.x;
    if (c < D && n < Npad) {
      float v = tile[threadIdx.x][i];
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      XT32[c * Npad + n] = v;
      XTh[c * Npad + n] = h;
      XTl[c * Npad + n] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// GL[(n,k), u] = sum_l Lambda[n,k,l] G[(n,l), u]
__global__ void tc_lambda_g_kernel(const float* __restrict__ Lam, const float* __restrict__ G, long long N,
                                   int K, int M, float* __restrict__ GL) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K * M) return;
  int u = (int)(i % M); long long r = i / M; long long n = r / K; int k = (int)(r % K);
  float acc = 0.f;
  for (int l = 0; l < K; ++l) acc = fmaf(Lam[(n * K + k) * K + l], G[(n * K + l) * M + u], acc);
  GL[i] = acc;
}

// Ct[(u'*M + u), n] = sum_k GL[(n,k), u] * G[(n,k), u']   (0 for n >= N).  One block per NB examples
// (NB = 32 / 16 / 8, the largest whose two [NB, K*M] staging tiles fit in shared memory).
template <int NB>
__global__ void ct_kernel(const float* __restrict__ GL, const float* __restrict__ G, long long N, int K,
                          int M, long long Npad, float* __restrict__ Ct) {
  extern __shared__ float sh[];
  const int KM = K * M;
  float* sgl = sh;
  float* sg = sh + NB * KM;
  const long long n0 = (long long)blockIdx.x * NB;
  for (int i = threadIdx.x; i < NB * KM; i += blockDim.x) {
    long long n = n0 + i / KM;
    bool ok = n < N;
    sgl[i] = ok ? GL[n * KM + i % KM] : 0.f;
    sg[i] = ok ? G[n * KM + i % KM] : 0.f;
  }
  __syncthreads();
  const int nl = threadIdx.x % NB;
  for (int e = threadIdx.x / NB; e < M * M; e += blockDim.x / NB) {
    int up = e / M, u = e % M;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(sgl[nl * KM + k * M + u], sg[nl * KM + k * M + up], acc);
    Ct[(long long)e * Npad + n0 + nl] = acc;
  }
}
template <int NB>
static int launch_ct(const float* GL, const float* G, long long N, int K, int M, long long Npad, float* Ct,
                     cudaStream_t st) {
  size_t sh = (size_t)2 * NB * K * M * sizeof(float);
  BNNP_CUDA_CHECK(cudaFuncSetAttribute(ct_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  ct_kernel<NB><<<(unsigned)(Npad / NB), 256, sh, st>>>(GL, G, N, K, M, Npad, Ct);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

// Bias columns (SIMT): H[row_theta0 + r, bias_theta0p + i'] += sum_n Ct[(u'_i' * M + u(r)), n] * XT32[src(r), n]
// one warp per (r, i').
__global__ void bias_cols_kernel(const float* __restrict__ XT32, const float* __restrict__ Ct,
                                 float* __restrict__ H, BlockArgs a, long long bias_theta0p, int r_begin, int r_end) {
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwork = (long long)(r_end - r_begin) * a.mp;
  if (w >= nwork) return;
  const int r = r_begin + (int)(w / a.mp), ip = (int)(w % a.mp);
  int u = a.unit0, src = a.col0;
  if (r < a.m * a.n) { u += r / a.n; src += r % a.n; } else { u += r - a.m * a.n; src += a.n; }
  const float* x = XT32 + (long long)src * a.Npad;
  const float* c = Ct + ((long long)(a.unit0p + ip) * a.Mtot + u) * a.Npad;
  float acc = 0.f;
  for (long long n = lane; n < a.Npad; n += 32) acc = fmaf(c[n], x[n], acc);
  acc = warp_sum(acc);
  if (lane == 0) H[(a.row_theta0 + r) * a.P + bias_theta0p + ip] += acc;
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

static int make_b_map(CUtensorMap* map, const __nv_bfloat16* base, long long rows, long long Npad, int wc) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return BNNP_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)Npad, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)Npad * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)wc};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), gdim, gstride, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[bnnp_b200] cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return BNNP_ERR_CUDA;
  }
  return BNNP_OK;
}

}  // namespace tc

// ---------------------------------------------------------------- host side
static inline long long pad_to(long long v, long long a) { return (v + a - 1) / a * a; }
// examples per tensor-core accumulation (measured truncation bias ~ -5.6e-9 per example: 2.3e-5 at 4096)
constexpr int kMaxAccumExamples = 4096;

// Optional CUDA-event timing of every gram_tc_kernel launch (bench.py's live roofline number).
struct GramProf {
  bool on = false;
  int n = 0;
  static constexpr int CAP = 4096;
  cudaEvent_t beg[CAP], end[CAP];
  double executed[CAP], useful[CAP];
  bool created = false;
};
static GramProf g_prof;

// MMA flops of one launch: executed (padded tiles as issued, 3 split passes) and useful (elements of H on or below
// its diagonal only, valid rows / columns only -- the strict count the roofline fraction is quoted on).
// mode: 0 single-CTA kernel, 1 CTA-pair kernel (two units), 2 wide kernel (one unit, two column tiles)
static void gram_flops(const tc::BlockArgs& a, int mode, double* executed, double* useful) {
  const long long rows_total = (long long)a.m * a.n + (a.has_bias ? a.m : 0);
  const double kflops = 3.0 * 2.0 * (double)a.n_chunks * tc::BK;
  double ex = 0, us = 0;
  // lower-triangle element count of rows [r0, r1) x cols [c0, c1) (global indices)
  auto lower = [](long long r0, long long r1, long long c0, long long c1) {
    double n = 0;
    for (long long r = r0; r < r1; ++r) { long long k = r - c0 + 1; n += (double)(k < 0 ? 0 : (k > c1 - c0 ? c1 - c0 : k)); }
    return n;
  };
  const int tile_rows = mode == 0 ? tc::TILE_M : 2 * tc::TILE_M;
  const int t0 = mode == 0 ? a.rt0 : a.rtp0, nt = mode == 0 ? a.n_rt : a.n_rtp;
  for (int t = t0; t < t0 + nt; ++t) {
    const long long row0 = (long long)t * tile_rows;
    const long long row1 = row0 + tile_rows < rows_total ? row0 + tile_rows : rows_total;
    const long long gr0 = a.row_theta0 + row0, gr1 = a.row_theta0 + row1;
    for (int ip = 0; ip < a.mp; ++ip)
      for (int jt = 0; jt < a.n_jt; ++jt) {
        const long long jc0 = (long long)jt * a.wc;
        const long long vc = a.np - jc0 < a.wc ? a.np - jc0 : a.wc;
        const long long gc0 = a.col_theta0 + (long long)ip * a.np + jc0;
        if (a.diag_block) {
          // tile-skipping granularity of each kernel: unit pair x column tile (0, 1) / unit x column tile, decided
          // on the first column of the group the accumulator belongs to (second tile of a wide pair: its own)
          long long group_c0 = gc0;
          if (mode != 2) group_c0 = a.col_theta0 + (long long)(ip & ~1) * a.np + jc0;
          else if (jt & 1) {
            const long long first = a.col_theta0 + (long long)ip * a.np + (long long)(jt - 1) * a.wc;
            group_c0 = (gr1 - 1 >= first + a.wc) ? first : gr1;      // skipped on its own when entirely above the diagonal
          }
          if (gr1 - 1 < group_c0) continue;
          us += kflops * lower(gr0, gr1, gc0, gc0 + vc);
        } else {
          us += kflops * (double)(row1 - row0) * (double)vc;
        }
        ex += kflops * tile_rows * (double)((vc + 15) / 16 * 16);
      }
  }
  *executed = ex; *useful = us;
}

struct TcLayout {
  long long Npad, off_counter, off_xt32, off_xth, off_xtl, off_gl, off_ct, total;
};
static TcLayout tc_layout(const bnnp_plan_t* plan, long long N) {
  TcLayout L;
  L.Npad = pad_to(N, 64);
  long long D = plan->Dtot, M = plan->Mtot, K = plan->K;
  long long cur = 0;
  auto take = [&](long long floats) { long long o = cur; cur += pad_to(floats, 256); return o; };
  L.off_counter = take(4096);
  L.off_xt32 = take(D * L.Npad);
  L.off_xth = take((D * L.Npad + 1) / 2);
  L.off_xtl = take((D * L.Npad + 1) / 2);
  L.off_gl = take(N * K * M);
  L.off_ct = take(M * M * L.Npad);
  L.total = cur;
  return L;
}

int64_t gram_tc_scratch_floats(const bnnp_plan_t* plan, int64_t N) { return tc_layout(plan, N).total + 256; }

bool gram_tc_preferred(const bnnp_op_t*, int, const bnnp_plan_t* plan, int64_t N) {
  // tiny problems are launch/padding bound on 128 x 256 tiles: keep them on the SIMT path
  return plan->P >= 2048 && N >= 256 && 8LL * plan->K * plan->Mtot * 2 * 4 <= 200 * 1024;
}

// rows [row_begin, row_end) of H only (row_end < 0: all rows).  A boundary that falls inside a layer's row block must
// sit on a multiple of 256 rows of that block (one CTA-pair row tile) -- gram_row_chunks() produces such boundaries.
// prepared != 0: the per-batch operands in `scratch` (transposed activations, C_n) were already built by an earlier call
// for the same batch and are reused.
int gram_tc_full_accum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, const float* ws,
                       const float* Lambda, float* H, float* scratch, int64_t row_begin, int64_t row_end, int prepared,
                       cudaStream_t st) {
  using namespace tc;
  const int K = plan->K, M = plan->Mtot, D = plan->Dtot;
  if ((long long)8 * K * M * 2 * 4 > 200 * 1024) return BNNP_ERR_UNSUPPORTED;
  if (row_end < 0) { row_begin = 0; row_end = plan->P; }
  if (row_begin < 0 || row_end > plan->P || row_begin > row_end) return BNNP_ERR_INVALID;
  // scratch must be 1 KB aligned for the tensor maps / vector loads
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~uintptr_t(1023));
  const TcLayout L = tc_layout(plan, N);
  int* counters = reinterpret_cast<int*>(base + L.off_counter);
  float* XT32 = base + L.off_xt32;
  __nv_bfloat16* XTh = reinterpret_cast<__nv_bfloat16*>(base + L.off_xth);
  __nv_bfloat16* XTl = reinterpret_cast<__nv_bfloat16*>(base + L.off_xtl);
  float* GL = base + L.off_gl;
  float* Ct = base + L.off_ct;
  const float* A = ws + plan->acat_off;
  const float* G = ws + plan->gcat_off;

  BNNP_CUDA_CHECK(cudaMemsetAsync(counters, 0, 4096 * sizeof(int), st));
  if (!prepared) {
    dim3 grid((unsigned)(L.Npad / 32), (unsigned)((D + 31) / 32));
    transpose_split_kernel<<<grid, dim3(32, 8), 0, st>>>(A, D, N, D, L.Npad, XT32, XTh, XTl);
    BNNP_LAUNCH_CHECK();
    tc_lambda_g_kernel<<<(unsigned)ceil_div64(N * K * M, 256), 256, 0, st>>>(Lambda, G, N, K, M, GL);
    BNNP_LAUNCH_CHECK();
    const long long per = 2LL * K * M * (long long)sizeof(float);
    int rc;
    if (32 * per <= 200 * 1024) rc = launch_ct<32>(GL, G, N, K, M, L.Npad, Ct, st);
    else if (16 * per <= 200 * 1024) rc = launch_ct<16>(GL, G, N, K, M, L.Npad, Ct, st);
    else rc = launch_ct<8>(GL, G, N, K, M, L.Npad, Ct, st);
    if (rc) return rc;
  }
  static bool attrs_set = false;
  if (!attrs_set) {
    BNNP_CUDA_CHECK(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    BNNP_CUDA_CHECK(cudaFuncSetAttribute(gram_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM_BYTES));
#ifdef BNNP_EXPERIMENTS
    BNNP_CUDA_CHECK(cudaFuncSetAttribute(gram_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
#endif
    attrs_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  int pair_idx = 0;
  for (int i = 0; i < n_ops; ++i) {
    if (ops[i].kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_t& R = ops[i];
    if (R.has_bias && R.theta_off_b != R.theta_off_w + (long long)R.out_c * R.in_c) return BNNP_ERR_UNSUPPORTED;
    const long long rows_total = (long long)R.out_c * R.in_c + (R.has_bias ? R.out_c : 0);
    // rows of this layer's block inside [row_begin, row_end)
    long long lr0 = row_begin - R.theta_off_w, lr1 = row_end - R.theta_off_w;
    lr0 = lr0 < 0 ? 0 : (lr0 > rows_total ? rows_total : lr0);
    lr1 = lr1 < 0 ? 0 : (lr1 > rows_total ? rows_total : lr1);
    if (lr0 >= lr1) continue;
    if (lr0 % (2 * TILE_M) != 0 || (lr1 % (2 * TILE_M) != 0 && lr1 != rows_total)) return BNNP_ERR_INVALID;
    for (int j = 0; j <= i; ++j) {
      if (ops[j].kind != BNNP_OP_LINEAR) continue;
      const bnnp_op_t& Cc = ops[j];
      BlockArgs a;
      a.m = R.out_c; a.n = R.in_c; a.has_bias = R.has_bias;
      a.unit0 = plan->op[i].lin_unit; a.col0 = plan->op[i].lin_col; a.row_theta0 = R.theta_off_w;
      a.mp = Cc.out_c; a.np = Cc.in_c; a.unit0p = plan->op[j].lin_unit; a.col_theta0 = Cc.theta_off_w;
      a.rt0 = (int)(lr0 / TILE_M);
      a.n_rt = (int)((lr1 + TILE_M - 1) / TILE_M) - a.rt0;
      a.rtp0 = a.rt0 / 2;
      a.n_rtp = (int)((lr1 + 2 * TILE_M - 1) / (2 * TILE_M)) - a.rtp0;
      // kernel: wide CTA-pair kernel (one unit, two column tiles sharing the generated A operand, M = 256 over two row
      // tiles) whenever the block has at least two row tiles; the single-CTA kernel otherwise (or by option)
      int mode = (rows_total + TILE_M - 1) / TILE_M >= 2 ? 2 : 0;
      if (g_bnnp_opt.gram_kernel == 1) mode = 0;
#ifdef BNNP_EXPERIMENTS
      if (g_bnnp_opt.gram_kernel == 2 && mode == 2) mode = 1;
#endif
      // column tile width: widest box (<= 256, multiple of 16) whose tiling of np wastes the least MMA work
      // when the last tile issues a narrower instruction (e.g. 784 = 3 x 224 + 112 instead of 4 x 208); the wide
      // kernel pays the per-tile generator / epilogue overhead once per PAIR of column tiles
      {
        int best_w = 16; long long best_cost = -1;
        for (int w = 256; w >= 64; w -= 16) {
          const int nj = (a.np + w - 1) / w;
          const int last = a.np - (nj - 1) * w;
          const int last16 = (last + 15) / 16 * 16;
          // executed MMA columns (a very narrow last tile is smem-bound: count it as >= 64) + ~64 columns worth of
          // per-tile generator / epilogue work
          const long long groups = mode == 2 ? (nj + 1) / 2 : nj;
          const long long cost = (long long)(nj - 1) * w + (last16 < 64 ? 64 : last16) + 64LL * groups;
          if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_w = w; }
        }
        a.wc = a.np < 64 ? (a.np + 15) / 16 * 16 : best_w;
      }
      a.n_jt = (a.np + a.wc - 1) / a.wc;
      a.n_pairs = (a.mp + 1) / 2;
      a.diag_block = (i == j);
      a.Mtot = M; a.Npad = L.Npad; a.P = plan->P;
      CUtensorMap mh, ml;
      int rc = make_b_map(&mh, XTh + (long long)plan->op[j].lin_col * L.Npad, a.np, L.Npad, mode ? a.wc / 2 : a.wc);
      if (rc) return rc;
      rc = make_b_map(&ml, XTl + (long long)plan->op[j].lin_col * L.Npad, a.np, L.Npad, mode ? a.wc / 2 : a.wc);
      if (rc) return rc;
      long long n_tiles;
      if (mode == 0) n_tiles = (long long)a.n_pairs * a.n_jt * a.n_rt;
      else if (mode == 1) n_tiles = (long long)a.n_pairs * a.n_jt * a.n_rtp;
      else n_tiles = (long long)a.mp * ((a.n_jt + 1) / 2) * a.n_rtp;
      const int grid = mode ? 2 * (int)(n_tiles < sms / 2 ? n_tiles : sms / 2) : (int)(n_tiles < sms ? n_tiles : sms);
      // The TMEM accumulators truncate on every update, so one accumulation never spans more than
      // kMaxAccumExamples examples; the segments are combined in fp32 (round-to-nearest) through H.
      const int total_chunks = (int)(L.Npad / BK), seg_chunks = kMaxAccumExamples / BK;
      for (int c0 = 0; c0 < total_chunks; c0 += seg_chunks) {
        a.chunk0 = c0;
        a.n_chunks = total_chunks - c0 < seg_chunks ? total_chunks - c0 : seg_chunks;
        if (pair_idx >= 4096) return BNNP_ERR_UNSUPPORTED;
        const bool prof = g_prof.on && g_prof.n < GramProf::CAP;
        if (prof) {
          if (!g_prof.created) {
            for (int e = 0; e < GramProf::CAP; ++e) { cudaEventCreate(&g_prof.beg[e]); cudaEventCreate(&g_prof.end[e]); }
            g_prof.created = true;
          }
          gram_flops(a, mode, &g_prof.executed[g_prof.n], &g_prof.useful[g_prof.n]);
          cudaEventRecord(g_prof.beg[g_prof.n], st);
        }
        if (mode == 2) gram_wide_kernel<<<grid, NUM_THREADS, W_SMEM_BYTES, st>>>(mh, ml, XT32, Ct, H, counters + pair_idx, a);
#ifdef BNNP_EXPERIMENTS
        else if (mode == 1) gram_pair_kernel<<<grid, NUM_THREADS, P_SMEM_BYTES, st>>>(mh, ml, XT32, Ct, H, counters + pair_idx, a);
#endif
        else gram_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mh, ml, XT32, Ct, H, counters + pair_idx, a);
        if (prof) { cudaEventRecord(g_prof.end[g_prof.n], st); ++g_prof.n; }
        BNNP_LAUNCH_CHECK();
        ++pair_idx;
      }
      a.chunk0 = 0;
      a.n_chunks = total_chunks;
      // bias columns of layer l' (SIMT): all rows for l' < l, bias rows only on the diagonal block
      if (Cc.has_bias) {
        long long r_begin = (i == j) ? (long long)a.m * a.n : 0;
        if (r_begin < lr0) r_begin = lr0;
        const long long nwork = (lr1 - r_begin) * a.mp;
        if (nwork > 0) {
          bias_cols_kernel<<<(unsigned)ceil_div64(nwork * 32, 256), 256, 0, st>>>(XT32, Ct, H, a, Cc.theta_off_b, (int)r_begin,
                                                                                  (int)lr1);
          BNNP_LAUNCH_CHECK();
        }
      }
    }
  }
  return BNNP_OK;
}

// Row boundaries b[0] = 0 < b[1] < ... < b[n] = P that split the lower triangle of H into n_chunks chunks whose AREAS
// decrease geometrically (ratio 0.6): the exchange of a chunk overlaps the accumulation of the next ones, only the
// last chunk's exchange is exposed, so the last chunk is made small (3 % of the triangle for 6 chunks) while the first
// ones stay large enough to keep the per-launch tail negligible.  Each boundary is legal for gram_tc_full_accum (a
// multiple of 256 rows inside its layer block, or a block edge).  Returns the number of chunks actually produced
// (<= n_chunks; boundaries that collapse are dropped).
int gram_row_chunks(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int n_chunks, int64_t* bounds) {
  const long long P = plan->P;
  int n = 0;
  bounds[n++] = 0;
  double total = 0, w = 1;
  for (int c = 0; c < n_chunks; ++c) { total += w; w *= 0.6; }
  double cum = 0;
  w = 1;
  for (int c = 1; c < n_chunks; ++c) {
    cum += w / total;
    w *= 0.6;
    long long want = (long long)((double)P * sqrt(cum));
    long long snapped = want;
    for (int i = 0; i < n_ops; ++i) {
      if (ops[i].kind != BNNP_OP_LINEAR) continue;
      const long long b0 = ops[i].theta_off_w;
      const long long rows = (long long)ops[i].out_c * ops[i].in_c + (ops[i].has_bias ? ops[i].out_c : 0);
      if (want >= b0 && want < b0 + rows) { snapped = b0 + (want - b0) / (2 * tc::TILE_M) * (2 * tc::TILE_M); break; }
    }
    if (snapped > bounds[n - 1] && snapped < P) bounds[n++] = snapped;
  }
  bounds[n] = P;
  return n;
}

}  // namespace bnnp

extern "C" int bnnp_prof_enable(int on) {
  bnnp::g_prof.on = on != 0;
  bnnp::g_prof.n = 0;
  bnnp::glm_prof_set(on != 0);
  return BNNP_OK;
}
// Copies up to cap records {ms, executed flops, useful flops} of the gram launches since enable;
// returns the number of records (or a negative error).
extern "C" int bnnp_prof_read(float* ms, double* executed, double* useful, int cap) {
  int n = bnnp::g_prof.n < cap ? bnnp::g_prof.n : cap;
  for (int i = 0; i < n; ++i) {
    if (cudaEventSynchronize(bnnp::g_prof.end[i]) != cudaSuccess) return BNNP_ERR_CUDA;
    if (cudaEventElapsedTime(&ms[i], bnnp::g_prof.beg[i], bnnp::g_prof.end[i]) != cudaSuccess) return BNNP_ERR_CUDA;
    executed[i] = bnnp::g_prof.executed[i];
    useful[i] = bnnp::g_prof.useful[i];
  }
  return n;
}
