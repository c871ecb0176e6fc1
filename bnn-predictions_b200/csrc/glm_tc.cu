// GLM predictive, full covariance, on tensor cores with TMA-fed operands.
//
//   f_var[n] = J_n Sigma J_n^T        (preds/predictives.py:63, einsum('nkp,pq,ncq->nkc'))
//
// For Linear layers J_n[k,(i,j)] = g_n[k,i] a_n[j], hence
//   f_var[n,k,l] = sum_{u,u'} G[n,k,u] T[n,u,u'] G[n,l,u'],
//   T[n,u,u']    = sum_{p in unit u} a~_n[j(p)] * Y_u'[n,p],    Y_u'[n,p] = sum_{q in unit u'} Sigma[p,q] a~_n[j(q)].
// Y is a plain GEMM per unit u' (M = inputs, N = all P parameters, K = inputs of the unit's layer) whose
// operands are ordinary matrices: the activations and Sigma itself.  Both are pre-split into bf16 hi/lo
// copies with 16-byte aligned rows and streamed by TMA (SWIZZLE_64B boxes) straight into the UMMA
// layout -- no generator warps.  Y never leaves the chip: the epilogue multiplies each accumulator row
// by a~_n[j(p)] and reduces over the columns of a unit (adding the unit's bias column, a rank-one term,
// from the fp32 Sigma on the way).  P^2 MACs per input instead of the reference's K*P^2.
#include <stdlib.h>
#include <vector>
#include "tc_common.cuh"

namespace bnnp {
namespace glmtc {

using namespace bnnp::tcp;

constexpr int BK = 32, TILE_M = 128, MAX_N = 256, STAGES = 3;
constexpr int A_TERM = TILE_M * BK * 2, A_STAGE = 2 * A_TERM;
constexpr int B_TERM = MAX_N * BK * 2, B_STAGE = 2 * B_TERM;
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;               // 48 KB
constexpr int META_BYTES = 4 * MAX_N * 12;                  // per accumulator buffer x 2 (consecutive tiles of a warp group): unit id, activation column, bias value
constexpr int EPI_STAGE_BYTES_FWD = 8 * 32 * 33 * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + META_BYTES + EPI_STAGE_BYTES_FWD;
constexpr int NUM_THREADS = 384;     // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-11: epilogue (two per TMEM lane quarter)
constexpr int EPI_WARPS = 8;          // warps 4-7 own accumulator buffer 0 (even tiles), warps 8-11 buffer 1 (odd tiles):
                                     // each group has two mainloop periods for its tile, which hides the exposed
                                     // global-load latencies of the epilogue under the TMA traffic
constexpr int EPI_WARP0 = 4;

struct Args {
  long long N, P;                    // inputs, parameters
  int n_units, nin, has_bias;        // units of the column-side layer l', its fan-in
  long long offw, offb;              // theta offsets of layer l' weights / biases
  int unit0p;                        // global index of the layer's first unit
  int unit_first, unit_step, delta;  // this launch covers units unit_first + i*unit_step (same K-start misalignment delta)
  int wn, n_mt, n_nt, n_chunks;
  int n_mtp;                         // CTA-pair kernel: input tiles are taken two at a time
  long long ald; int Mtot;
  const float* Sigma; const float* A; const int32_t* pu; const int32_t* pc; float* T;
  const long long* tile_start;       // [n_units + 1] prefix of tiles per unit of this launch (lower triangle only)
};

// epilogue of one accumulator tile: T[n, u(p), u'] += a~[n, c(p)] * (Y[n,p] + Sigma[p, bias(u')]) for the 128 input
// rows m0.. held by this CTA's TMEM; q = epilogue warp 0..3 (TMEM lane quarter).
// Each thread owns one input row, so reading a~[n, c(p)] directly costs 32 L1 wavefronts per load instruction (32 rows
// = 32 cache lines) -- ncu showed l1tex as the busiest unit with these gathers as ~80 % of its load, and halving the
// operand bytes (CTA-pair kernel) changed nothing.  The 32 x 32 activation block of a column step is therefore staged
// through a per-warp shared-memory tile: lanes run along the columns for the global reads (1-2 wavefronts per row),
// along the rows for the (padded, conflict-free) reads back.
constexpr int EPI_STAGE_LD = 33;
constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * EPI_STAGE_LD * 4;
__device__ __forceinline__ void glm_epilogue_tile(const Args& args, int* meta, float* stage_all, int ew, int lane,
                                                  long long m0, long long n0, int up, uint32_t buf, uint32_t buf_phase,
                                                  uint64_t* acc_full_bar, uint32_t tmem_base) {
  const int q = ew & 3;                              // TMEM lane quarter (rows); ew >> 2 == buf (checked by the caller)
  const long long n = m0 + q * 32 + lane;
  const bool rowok = n < args.N;
  const int ncols = (int)(args.P - n0 < args.wn ? args.P - n0 : args.wn);
  // the column metadata is the same for all 128 rows of the tile: stage it in shared memory once
  // (unit id, activation column, bias-column value of Sigma -- row offb+u' of the symmetric matrix)
  // a group's consecutive tiles alternate between two copies: a fast warp staging tile i + 2 must not overwrite what a
  // slower warp of the group still reads for tile i
  int* m_u = meta + (buf * 2 + buf_phase) * 3 * MAX_N;
  int* m_c = m_u + MAX_N;
  float* m_b = reinterpret_cast<float*>(m_c + MAX_N);
  float* sA = stage_all + ew * 32 * EPI_STAGE_LD;
  {
    const float* sbias = args.Sigma + (args.offb + up) * args.P;
    for (int e = (q * 32 + lane); e < MAX_N; e += 128) {
      const long long p = n0 + (e < ncols ? e : ncols - 1);
      m_u[e] = e < ncols ? __ldg(args.pu + p) : -1;          // -1: padding column, contributes nothing
      m_c[e] = __ldg(args.pc + p);
      m_b[e] = (args.has_bias && e < ncols) ? __ldg(sbias + p) : 0.f;
    }
  }
  asm volatile("bar.sync %0, 128;" ::"r"(1 + (int)buf) : "memory");     // the four warps of this buffer's group
  // rows of this warp: first row and how many are inside the batch
  const long long r0 = m0 + q * 32;
  const int nrows = (int)(args.N - r0 < 32 ? (args.N - r0 < 0 ? 0 : args.N - r0) : 32);
  const float* __restrict__ abase = args.A + r0 * args.ald;
  float* trow = args.T + ((rowok ? n : 0) * args.Mtot) * (long long)args.Mtot + args.unit0p + up;
  float acc = 0.f;
  int ucur = -1;
  bool waited = false;
  // a~[r0 + r, c(c0 + lane)] for the warp's 32 rows, one column step ahead: the loads of step c0 + 32 are in flight
  // while step c0 is reduced (their latency under the TMA traffic is what bounded the epilogue)
  const int cbeg = 0, cend = ncols;
  float t[32];
  if (cbeg < cend) {
    const int cc = m_c[cbeg + lane];
#pragma unroll
    for (int r = 0; r < 32; ++r) t[r] = r < nrows ? __ldg(abase + (long long)r * args.ald + cc) : 0.f;
  }
  for (int c0 = cbeg; c0 < cend; c0 += 32) {
    __syncwarp();                                                // previous step's reads of sA are done
#pragma unroll
    for (int r = 0; r < 32; ++r) sA[r * EPI_STAGE_LD + lane] = t[r];
    __syncwarp();
    if (c0 + 32 < cend) {
      const int cc = m_c[c0 + 32 + lane];
#pragma unroll
      for (int r = 0; r < 32; ++r) t[r] = r < nrows ? __ldg(abase + (long long)r * args.ald + cc) : 0.f;
    }
    if (!waited) {
      mbar_wait(acc_full_bar, buf_phase);
      tc_fence_after();
      waited = true;
    }
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * MAX_N + c0), v);
    const float* my = sA + lane * EPI_STAGE_LD;
    const float4* b4 = reinterpret_cast<const float4*>(m_b + c0);
    const int u_first = m_u[c0], u_last = m_u[c0 + 31];
    if (u_first == u_last) {
      // all 32 columns belong to one unit (runs are nin + 1 long): branch-free dot product
      if (u_first != ucur) {
        if (ucur >= 0 && rowok) atomicAdd(trow + (long long)ucur * args.Mtot, acc);
        ucur = u_first; acc = 0.f;
      }
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 bb = b4[g];
        s0 = fmaf(my[g * 4 + 0], __uint_as_float(v[g * 4 + 0]) + bb.x, s0);
        s1 = fmaf(my[g * 4 + 1], __uint_as_float(v[g * 4 + 1]) + bb.y, s1);
        s0 = fmaf(my[g * 4 + 2], __uint_as_float(v[g * 4 + 2]) + bb.z, s0);
        s1 = fmaf(my[g * 4 + 3], __uint_as_float(v[g * 4 + 3]) + bb.w, s1);
      }
      acc += s0 + s1;
    } else {
      // a unit boundary (or the padding, unit -1) inside the step: run-length reduction, one atomic per run
      for (int e = 0; e < 32; ++e) {
        const int u = m_u[c0 + e];
        if (u != ucur) {
          if (ucur >= 0 && rowok) atomicAdd(trow + (long long)ucur * args.Mtot, acc);
          ucur = u; acc = 0.f;
        }
        float y;
        // v[] must stay in registers: select with a fully unrolled predicated chain instead of dynamic indexing
        y = __uint_as_float(v[0]);
#pragma unroll
        for (int k = 1; k < 32; ++k) y = (e == k) ? __uint_as_float(v[k]) : y;
        acc = fmaf(my[e], y + m_b[c0 + e], acc);
      }
    }
  }
  if (!waited) {                                       // no columns for this warp: still follow the accumulator hand-over
    mbar_wait(acc_full_bar, buf_phase);
    tc_fence_after();
  }
  if (rowok && ucur >= 0) atomicAdd(trow + (long long)ucur * args.Mtot, acc);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
glm_fvar_tma_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                    const __grid_constant__ CUtensorMap map_sh, const __grid_constant__ CUtensorMap map_sl,
                    Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                        // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;              // [STAGES] MMA commit -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;       // [2]
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  int* meta = reinterpret_cast<int*>(smem + STAGES * STAGE_BYTES + 256);     // [2][3][MAX_N]
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256 + META_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EPI_WARPS / 2); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // T[n,u,u'] is symmetric: for unit u' only the parameter rows from the unit's own first weight onward are
  // needed (units u >= u' of the layer, its later biases, all later layers); tile_start[] is the prefix of the
  // per-unit tile counts.  Entries above the diagonal receive partial sums and are never read.
  const long long n_tiles = args.tile_start[args.n_units];
  uint32_t stage = 0, phase = 0, iter = 0;
  int ui = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
    while (args.tile_start[ui + 1] <= tile) ++ui;                 // tiles are visited in increasing order
    const int up = args.unit_first + args.unit_step * ui;
    const int nt0 = (int)((args.offw + (long long)up * args.nin) / args.wn);
    const long long local = tile - args.tile_start[ui];
    // input tile fastest: concurrently running CTAs stream the same Sigma tile (shared through L2)
    const int mt = (int)(local % args.n_mt);
    const int nt = nt0 + (int)(local / args.n_mt);
    const long long m0 = (long long)mt * TILE_M, n0 = (long long)nt * args.wn;
    const uint32_t buf = iter & 1, buf_phase = (iter >> 1) & 1;

    if (warp == 0) {
      if (lane == 0) {
        uint32_t s = stage, ph = phase;
        const uint32_t bytes = (uint32_t)(A_STAGE + 2 * args.wn * BK * 2);
        // first Sigma column of unit u', moved down to a 16-byte boundary (the activation copy is shifted by delta)
        const int kcol0 = (int)(args.offw + (long long)up * args.nin) - args.delta;
        for (int c = 0; c < args.n_chunks; ++c) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* base = smem + s * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], bytes);
          tma_load_2d(&map_ah, &full[s], base, c * BK, (int)m0);
          tma_load_2d(&map_al, &full[s], base + A_TERM, c * BK, (int)m0);
          tma_load_2d(&map_sh, &full[s], base + A_STAGE, kcol0 + c * BK, (int)n0);
          tma_load_2d(&map_sl, &full[s], base + A_STAGE + B_TERM, kcol0 + c * BK, (int)n0);
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
        for (int c = 0; c < args.n_chunks; ++c) {
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
          if (c == args.n_chunks - 1) umma_commit(&acc_full[buf]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp >= EPI_WARP0 && (uint32_t)((warp - EPI_WARP0) >> 2) == buf) {
      glm_epilogue_tile(args, meta, epi_stage, warp - EPI_WARP0, lane, m0, n0, up, buf, buf_phase, &acc_full[buf], tmem_base);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    const uint32_t tot = stage + (uint32_t)args.n_chunks;
    phase ^= (tot / STAGES) & 1;
    stage = tot % STAGES;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---- CTA-pair version ------------------------------------------------------------------------------------
// The single-CTA kernel above is bound by the bytes each SM has to pull from L2: 48 KB per K-chunk against 786 tensor
// cycles (ncu: 6.7 TB/s L2->SM, tensor pipe 30 %).  Here two CTAs of a cluster (one TPC) run ONE M = 256
// `tcgen05.mma.cta_group::2`: each CTA stages its own 128 input rows and only HALF of the Sigma tile (wn/2 rows), the
// tensor core reads both halves -- 32 KB per chunk per SM for the same work.  Rank 0 issues the MMAs and commits to the
// barriers of both CTAs; every TMA load credits rank 0's `full` barrier; both CTAs run the epilogue on their own TMEM
// rows and release the accumulator buffer on rank 0's `acc_empty` barrier.
constexpr int P_STAGES = 5;
constexpr int P_B_TERM = (MAX_N / 2) * BK * 2, P_STAGE_BYTES = A_STAGE + 2 * P_B_TERM;      // 32 KB
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 256 + META_BYTES + EPI_STAGE_BYTES;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
glm_fvar_pair_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                     const __grid_constant__ CUtensorMap map_sh, const __grid_constant__ CUtensorMap map_sl,
                     Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* full = bars;                          // [P_STAGES] used in rank 0 only: all TMA bytes of the pair
  uint64_t* empty = bars + P_STAGES;              // [P_STAGES] in both CTAs: multicast commit of rank 0
  uint64_t* acc_full = bars + 2 * P_STAGES;       // [2] in both CTAs
  uint64_t* acc_empty = bars + 2 * P_STAGES + 2;  // [2] used in rank 0 only: 8 epilogue warps x 2 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P_STAGES + 4);
  int* meta = reinterpret_cast<int*>(smem + P_STAGES * P_STAGE_BYTES + 256);
  float* epi_stage = reinterpret_cast<float*>(smem + P_STAGES * P_STAGE_BYTES + 256 + META_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EPI_WARPS); }
    fence_barrier_init();
  }
  cluster_sync_all();                             // barriers of both CTAs exist before anything remote touches them
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_tiles = args.tile_start[args.n_units];
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int half = args.wn / 2;
  uint32_t stage = 0, phase = 0, iter = 0;
  int ui = 0;
  for (long long tile = pair; tile < n_tiles; tile += n_pairs, ++iter) {
    while (args.tile_start[ui + 1] <= tile) ++ui;
    const int up = args.unit_first + args.unit_step * ui;
    const int nt0 = (int)((args.offw + (long long)up * args.nin) / args.wn);
    const long long local = tile - args.tile_start[ui];
    const int mtp = (int)(local % args.n_mtp);
    const int nt = nt0 + (int)(local / args.n_mtp);
    const long long m0 = ((long long)mtp * 2 + rank) * TILE_M, n0 = (long long)nt * args.wn;
    const uint32_t buf = iter & 1, buf_phase = (iter >> 1) & 1;

    if (warp == 0) {
      if (lane == 0) {
        uint32_t s = stage, ph = phase;
        const uint32_t bytes = 2u * (uint32_t)(A_STAGE + 2 * half * BK * 2);
        const int kcol0 = (int)(args.offw + (long long)up * args.nin) - args.delta;
        for (int c = 0; c < args.n_chunks; ++c) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* base = smem + s * P_STAGE_BYTES;
          const uint32_t lead_full = mapa_shared(smem_u32(&full[s]), 0);
          // Only rank 0 arrives, expecting the bytes of BOTH CTAs.  Rank 1's loads may complete first (the transaction
          // count goes negative, the phase cannot end before rank 0's arrival) but never a phase early: its own `empty`
          // wait orders them after the MMA that consumed the previous use of the stage.  (A per-chunk remote
          // mbarrier.arrive.release.cluster here costs a GPU-scope MEMBAR per chunk and throttled the whole pipeline.)
          if (rank == 0) mbar_arrive_expect_tx(&full[s], bytes);
          tma_load_2d_pair(&map_ah, lead_full, base, c * BK, (int)m0);
          tma_load_2d_pair(&map_al, lead_full, base + A_TERM, c * BK, (int)m0);
          tma_load_2d_pair(&map_sh, lead_full, base + A_STAGE, kcol0 + c * BK, (int)(n0 + rank * half));
          tma_load_2d_pair(&map_sl, lead_full, base + A_STAGE + P_B_TERM, kcol0 + c * BK, (int)(n0 + rank * half));
          if (++s == P_STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && rank == 0) {
        mbar_wait(&acc_empty[buf], buf_phase ^ 1);
        tc_fence_after();
        uint32_t s = stage, ph = phase;
        const uint32_t idesc = umma_idesc_bf16_pair(args.wn);
        const uint32_t d = tmem_base + buf * MAX_N;
        for (int c = 0; c < args.n_chunks; ++c) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * P_STAGE_BYTES), b0 = a0 + A_STAGE;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t dah = umma_desc_sw64(a0 + kk * 32), dal = umma_desc_sw64(a0 + A_TERM + kk * 32);
            const uint64_t dbh = umma_desc_sw64(b0 + kk * 32), dbl = umma_desc_sw64(b0 + P_B_TERM + kk * 32);
            umma_bf16_pair(d, dah, dbh, idesc, (c | kk) ? 1u : 0u);
            umma_bf16_pair(d, dah, dbl, idesc, 1u);
            umma_bf16_pair(d, dal, dbh, idesc, 1u);
          }
          umma_commit_pair(&empty[s]);
          if (c == args.n_chunks - 1) umma_commit_pair(&acc_full[buf]);
          if (++s == P_STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp >= EPI_WARP0 && (uint32_t)((warp - EPI_WARP0) >> 2) == buf) {
      glm_epilogue_tile(args, meta, epi_stage, warp - EPI_WARP0, lane, m0, n0, up, buf, buf_phase, &acc_full[buf], tmem_base);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));
    }
    const uint32_t tot = stage + (uint32_t)args.n_chunks;
    phase ^= (tot / P_STAGES) & 1;
    stage = tot % P_STAGES;
  }
  tc_fence_before();
  cluster_sync_all();                             // nobody leaves while the peer may still signal its barriers / read its smem
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

// dst_hi/lo[r, c] = bf16 split of src[r*lds + c] for c < cols, 0 for cols <= c < cpad
__global__ void split_bf16_kernel(const float* __restrict__ src, long long lds, long long rows, long long cols,
                                  long long cpad, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                  int shift) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cpad) return;
  long long r = i / cpad, c = i % cpad - shift;
  float v = (c >= 0 && c < cols) ? src[r * lds + c] : 0.f;
  __nv_bfloat16 h = __float2bfloat16_rn(v);
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
// bf16 matrix [rows, cols] with row stride ld elements, boxes [box_rows x 32]
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return BNNP_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[bnnp_b200] cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return BNNP_ERR_CUDA;
  }
  return BNNP_OK;
}

}  // namespace glmtc
}  // namespace bnnp

using namespace bnnp;
using namespace bnnp::glmtc;

static inline long long up8(long long v) { return (v + 7) / 8 * 8; }
static inline long long up32(long long v) { return (v + 31) / 32 * 32; }

// Split a fp32 matrix into bf16 hi/lo copies with a 16-byte aligned row pitch (cpad = roundup8(cols)).
extern "C" int bnnp_split_bf16(const float* src, int64_t lds, int64_t rows, int64_t cols, void* hi, void* lo,
                               void* stream) {
  if (!src || !hi || !lo || rows <= 0 || cols <= 0) return BNNP_ERR_INVALID;
  const long long cpad = up8(cols);
  split_bf16_kernel<<<(unsigned)ceil_div64(rows * cpad, 256), 256, 0, as_stream(stream)>>>(
      src, lds, rows, cols, cpad, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), 0);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

// scratch layout (floats): [maps 2*Ppad ints][T N*M*M][tile prefix tables][per-layer activation splits]
extern "C" int64_t bnnp_glm_fvar_full_tma_scratch_floats(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                                         int64_t N) {
  const int64_t Ppad = (plan->P + 31) / 32 * 32;
  int64_t act = 0;
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind == BNNP_OP_LINEAR) act += 8 * (N * up32(ops[i].in_c + 8) + 64);   // <= 8 shifted copies, hi + lo bf16
  return 2 * Ppad + N * (int64_t)plan->Mtot * plan->Mtot + act + 2 * (int64_t)(plan->Mtot + 64) * 8 + 4096;
}

__global__ void glmtc_param_maps_kernel(int32_t* __restrict__ pu, int32_t* __restrict__ pc, int64_t off, int m,
                                        int nin, int unit0, int col0, int is_bias) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t cnt = is_bias ? m : (int64_t)m * nin;
  if (e >= cnt) return;
  if (is_bias) { pu[off + e] = unit0 + (int)e; pc[off + e] = col0 + nin; }
  else { pu[off + e] = unit0 + (int)(e / nin); pc[off + e] = col0 + (int)(e % nin); }
}
__global__ void glmtc_zero_kernel(float* p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}
// f_var[n,k,l] = sum_{u,u'} G[(n,k),u] Tsym[n,u,u'] G[(n,l),u'],  Tsym read from the lower triangle (u >= u')
__global__ void glmtc_fvar_from_units_kernel(const float* __restrict__ G, int64_t gld, const float* __restrict__ T,
                                             int K, int M, float* __restrict__ fvar) {
  extern __shared__ float shm[];
  const int64_t n = blockIdx.x;
  const float* Tn = T + n * M * M;
  for (int i = threadIdx.x; i < K * M; i += blockDim.x) {
    const int k = i / M, up = i % M;
    float acc = 0.f;
    for (int u = 0; u < M; ++u) acc = fmaf(G[(n * K + k) * gld + u], u >= up ? Tn[u * M + up] : Tn[up * M + u], acc);
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

// Optional CUDA-event timing of the f_var kernel launches of every call (bench.py's live roofline of the GLM
// predictive); one record per (layer, alignment group) launch, useful flops = 3 split passes x 2 x inputs x the
// elements of Sigma on/below the block diagonal that the launch contracts.
struct GlmProf {
  bool on = false, created = false;
  int n = 0;
  static constexpr int CAP = 1024;
  cudaEvent_t beg[CAP], end[CAP];
  double useful[CAP];
};
static GlmProf g_glm_prof;
namespace bnnp { void glm_prof_set(bool on) { g_glm_prof.on = on; g_glm_prof.n = 0; } }
extern "C" int bnnp_glm_prof_read(float* ms, double* useful, int cap) {
  int n = g_glm_prof.n < cap ? g_glm_prof.n : cap;
  for (int i = 0; i < n; ++i) {
    if (cudaEventSynchronize(g_glm_prof.end[i]) != cudaSuccess) return BNNP_ERR_CUDA;
    if (cudaEventElapsedTime(&ms[i], g_glm_prof.beg[i], g_glm_prof.end[i]) != cudaSuccess) return BNNP_ERR_CUDA;
    useful[i] = g_glm_prof.useful[i];
  }
  return n;
}

// Sigma: fp32 [P,P]; Sh/Sl: its bf16 hi/lo split with row pitch roundup8(P) (bnnp_split_bf16).
extern "C" int bnnp_glm_fvar_full_tma(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                                      const float* ws, const float* Sigma, const void* Sh, const void* Sl,
                                      float* f_var, float* scratch, void* stream) {
  if (!ops || !plan || !ws || !Sigma || !Sh || !Sl || !f_var || !scratch || N <= 0) return BNNP_ERR_INVALID;
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind == BNNP_OP_CONV2D || ops[i].kind == BNNP_OP_MAXPOOL2D || ops[i].kind == BNNP_OP_AVGPOOL2D)
      return BNNP_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int K = plan->K, M = plan->Mtot;
  const int64_t P = plan->P, Ppad = (P + 31) / 32 * 32, Ppitch = up8(P);
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255));
  int32_t* pu = reinterpret_cast<int32_t*>(base);
  int32_t* pc = pu + Ppad;
  float* T = base + 2 * Ppad;
  long long* prefix_dev = reinterpret_cast<long long*>(T + ((N * (int64_t)M * M + 63) / 64) * 64);
  float* act = reinterpret_cast<float*>(prefix_dev + (M + 64) * 8);
  long long prefix_used = 0;
  const float* A = ws + plan->acat_off;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    glmtc_param_maps_kernel<<<(unsigned)ceil_div64((int64_t)o.out_c * o.in_c, 256), 256, 0, st>>>(
        pu, pc, o.theta_off_w, o.out_c, o.in_c, q.lin_unit, q.lin_col, 0);
    BNNP_LAUNCH_CHECK();
    if (o.has_bias) {
      glmtc_param_maps_kernel<<<(unsigned)ceil_div64(o.out_c, 256), 256, 0, st>>>(pu, pc, o.theta_off_b, o.out_c,
                                                                                  o.in_c, q.lin_unit, q.lin_col, 1);
      BNNP_LAUNCH_CHECK();
    }
  }
  glmtc_zero_kernel<<<(unsigned)ceil_div64(N * (int64_t)M * M, 256), 256, 0, st>>>(T, N * (int64_t)M * M);
  BNNP_LAUNCH_CHECK();
  BNNP_CUDA_CHECK(cudaFuncSetAttribute(glm_fvar_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CUtensorMap msh, msl;
  Args a;
  a.N = N; a.P = P; a.ald = plan->Dtot; a.Mtot = M;
  a.Sigma = Sigma; a.A = A; a.pu = pu; a.pc = pc; a.T = T;
  {
    long long nt = (P + MAX_N - 1) / MAX_N;
    a.wn = (int)(((P + nt - 1) / nt + 15) / 16 * 16);
    a.n_nt = (int)((P + a.wn - 1) / a.wn);
    a.n_mt = (int)((N + TILE_M - 1) / TILE_M);
  }
  // CTA-pair kernel (two input tiles per M = 256 MMA, each CTA stages half of the Sigma tile) whenever there are at
  // least two input tiles; option glm_single forces the single-CTA kernel (A/B measurements)
  const bool pair = a.n_mt >= 2 && !g_bnnp_opt.glm_single;
  a.n_mtp = (a.n_mt + 1) / 2;
  int rc = make_map(&msh, Sh, P, P, Ppitch, pair ? a.wn / 2 : a.wn);
  if (rc) return rc;
  rc = make_map(&msl, Sl, P, P, Ppitch, pair ? a.wn / 2 : a.wn);
  if (rc) return rc;
  if (pair) BNNP_CUDA_CHECK(cudaFuncSetAttribute(glm_fvar_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind != BNNP_OP_LINEAR) continue;
    const bnnp_op_plan_t& q = plan->op[i];
    // TMA boxes must start on a 16-byte boundary, i.e. at a Sigma column that is a multiple of 8.  The K-range
    // of unit u' starts at offw + u'*nin; units are grouped by the misalignment delta of that start (periodic in
    // u'), and each group reads a copy of the layer's activations shifted right by delta (zero filled), so that
    // the Sigma columns pulled in below the unit's range meet zeros.
    int g8 = o.in_c % 8, period = 1;
    if (g8) { int a_ = g8, b_ = 8; while (b_) { int t = a_ % b_; a_ = b_; b_ = t; } period = 8 / a_; }
    for (int grp = 0; grp < period && grp < o.out_c; ++grp) {
      const int delta = (int)((o.theta_off_w + (long long)grp * o.in_c) % 8);
      const long long apitch = up32(o.in_c + delta);
      __nv_bfloat16* ah = reinterpret_cast<__nv_bfloat16*>(act);
      __nv_bfloat16* al = ah + N * apitch;
      act += N * apitch + 64;
      split_bf16_kernel<<<(unsigned)ceil_div64(N * apitch, 256), 256, 0, st>>>(A + q.lin_col, plan->Dtot, N, o.in_c,
                                                                               apitch, ah, al, delta);
      BNNP_LAUNCH_CHECK();
      CUtensorMap mah, mal;
      rc = make_map(&mah, ah, N, apitch, apitch, TILE_M);
      if (rc) return rc;
      rc = make_map(&mal, al, N, apitch, apitch, TILE_M);
      if (rc) return rc;
      a.unit_first = grp; a.unit_step = period; a.delta = delta;
      a.n_units = (o.out_c - grp + period - 1) / period;
      a.nin = o.in_c; a.has_bias = o.has_bias;
      a.offw = o.theta_off_w; a.offb = o.has_bias ? o.theta_off_b : 0; a.unit0p = q.lin_unit;
      a.n_chunks = (o.in_c + delta + BK - 1) / BK;
      // per-unit tile counts (lower triangle): n-tiles from the one holding the unit's first weight row
      std::vector<long long> prefix(a.n_units + 1, 0);
      for (int ui = 0; ui < a.n_units; ++ui) {
        const long long up = grp + (long long)period * ui;
        const long long nt0 = (o.theta_off_w + up * o.in_c) / a.wn;
        prefix[ui + 1] = prefix[ui] + (a.n_nt - nt0) * (pair ? a.n_mtp : a.n_mt);
      }
      if (prefix_used + a.n_units + 1 > (long long)(M + 64) * 8) return BNNP_ERR_UNSUPPORTED;
      BNNP_CUDA_CHECK(cudaMemcpyAsync(prefix_dev + prefix_used, prefix.data(), sizeof(long long) * (a.n_units + 1),
                                      cudaMemcpyHostToDevice, st));
      a.tile_start = prefix_dev + prefix_used;
      prefix_used += a.n_units + 1;
      const long long n_tiles = prefix[a.n_units];
      const bool prof = g_glm_prof.on && g_glm_prof.n < GlmProf::CAP;
      if (prof) {
        if (!g_glm_prof.created) {
          for (int e = 0; e < GlmProf::CAP; ++e) { cudaEventCreate(&g_glm_prof.beg[e]); cudaEventCreate(&g_glm_prof.end[e]); }
          g_glm_prof.created = true;
        }
        // unit u' contracts its fan-in (+ bias handled in the epilogue) with the Sigma rows from its first weight row on
        double us = 0;
        for (int ui = 0; ui < a.n_units; ++ui) {
          const long long up = grp + (long long)period * ui;
          us += (double)o.in_c * (double)(P - (o.theta_off_w + up * o.in_c));
        }
        g_glm_prof.useful[g_glm_prof.n] = 3.0 * 2.0 * (double)N * us;
        cudaEventRecord(g_glm_prof.beg[g_glm_prof.n], st);
      }
      if (pair) {
        const long long pairs = sms / 2;
        const int grid = 2 * (int)(n_tiles < pairs ? n_tiles : pairs);
        glm_fvar_pair_kernel<<<grid, NUM_THREADS, P_SMEM_BYTES, st>>>(mah, mal, msh, msl, a);
      } else {
        const int grid = (int)(n_tiles < sms ? n_tiles : sms);
        glm_fvar_tma_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mah, mal, msh, msl, a);
      }
      if (prof) { cudaEventRecord(g_glm_prof.end[g_glm_prof.n], st); ++g_glm_prof.n; }
      BNNP_LAUNCH_CHECK();
      if (g_bnnp_opt.debug_sync) {
        cudaError_t e = cudaStreamSynchronize(st);
        fprintf(stderr, "[bnnp debug] glm_fvar_tma op %d grp %d delta %d: units %d nin %d wn %d n_nt %d n_mt %d chunks %d -> %s\n",
                i, grp, delta, a.n_units, a.nin, a.wn, a.n_nt, a.n_mt, a.n_chunks, cudaGetErrorString(e));
        if (e != cudaSuccess) return BNNP_ERR_CUDA;
      }
    }
  }
  glmtc_fvar_from_units_kernel<<<(unsigned)N, 256, (size_t)K * M * sizeof(float), st>>>(ws + plan->gcat_off, M, T, K,
                                                                                       M, f_var);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}
