// The ONE exchange step of the data-parallel full-covariance path (preds/laplace.py:39-42 sharded over examples,
// SURVEY.md 8e): every rank holds its local GGN  H_r = sum_{n in shard r} J_n^T Lambda_n J_n  in the lower triangle of a
// row-major [P,P] fp32 matrix that lives at the same offset of a SYMMETRIC allocation on every GPU of the NVSwitch
// domain (torch.distributed._symmetric_memory provides the mappings; that is plumbing).  This kernel is the collective:
//
//   H[p, 0..p] <- sum_r H_r[p, 0..p]  (+ prior on the diagonal),  identical bits on every rank,
//
// over peer memory, touching only the triangle (half the bytes of an all-reduce of the square), with the prior
// precision fused into the owner's pass.  Row p is owned by rank p mod W:
//   * NVLS mode (multicast mapping available): `multimem.ld_reduce.add.f32` pulls the sum of all W copies THROUGH THE
//     SWITCH (the reduction happens in NVSwitch, one response crosses the owner's link), `multimem.st` broadcasts the
//     result to all W copies.  Per GPU and direction: (1 + 1/W) x triangle bytes.
//   * P2P mode (no multicast): the owner loads the row from every peer mapping in rank order (fixed summation order),
//     and stores the sum to every peer.  Per GPU and direction: 2 (W-1)/W x triangle bytes.
// Each element is reduced exactly once, by one rank, and the same value is written everywhere: no rank-dependent
// rounding, so the redundant factorisations that follow see bit-identical inputs.
// Row ranges [row0, row1) let the caller exchange finished row chunks on a side stream while later chunks still
// accumulate (overlap of the collective with the tensor-core kernel).  Cross-rank ordering (all ranks finished
// accumulating the chunk / all owners finished broadcasting it) is the caller's job: a symmetric-memory barrier on the
// same stream before and after the launch.
#include "common.cuh"

namespace bnnp {
namespace xch {

constexpr int MAX_WORLD = 16;
constexpr int THREADS = 512;
// 16-byte requests in flight per thread.  The exchange runs beside the persistent tensor-core kernel, so it must
// saturate NVLink from FEW SMs: measured at W = 2, P = 59 635 with 4 requests per thread the time halves with every
// doubling of the CTA count up to 128 CTAs (memory-level parallelism bound, ~4 MB in flight needed); with 16 (NVLS) /
// 8 (P2P: two register sets, the running sum and the peer's data) the same bytes are in flight from 32 / 64 CTAs.
constexpr int UNROLL_MC = 16;
constexpr int UNROLL_P2P = 8;

struct Args {
  float* peer[MAX_WORLD];      // rank w's H in this process's address space
  float* mc;                   // multicast mapping of H (NVLS) or nullptr
  const float* prior_vec;      // [P] or nullptr
  float prior_scalar;
  int rank, world;
  long long P, row0, row1;
};

__device__ __forceinline__ float4 mc_ldreduce4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float mc_ldreduce1(const float* p) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st1(float* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float4 sys_ld4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float sys_ld1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sys_st4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sys_st1(float* p, float v) {
  asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <bool MC>
__device__ __forceinline__ float reduce1(const Args& a, long long e) {
  if (MC) return mc_ldreduce1(a.mc + e);
  float s = 0.f;
  for (int w = 0; w < a.world; ++w) s += sys_ld1(a.peer[w] + e);
  return s;
}
template <bool MC>
__device__ __forceinline__ void bcast1(const Args& a, long long e, float v) {
  if (MC) { mc_st1(a.mc + e, v); return; }
  for (int w = 0; w < a.world; ++w) sys_st1(a.peer[w] + e, v);
}

// grid: any number of CTAs; row p of [row0,row1) with p % world == rank is handled by CTA ((p / world) % gridDim.x)
template <bool MC>
__global__ void __launch_bounds__(THREADS)
exchange_lower_kernel(const Args a) {
  constexpr int UNROLL = MC ? UNROLL_MC : UNROLL_P2P;
  const int W = a.world;
  long long first = a.row0 + ((a.rank - a.row0 % W) % W + W) % W;      // first owned row >= row0
  for (long long p = first + (long long)blockIdx.x * W; p < a.row1; p += (long long)gridDim.x * W) {
    const long long e0 = p * a.P, e1 = e0 + p + 1;                     // flat element range of the row's lower part
    const float prior = a.prior_vec ? a.prior_vec[p] : a.prior_scalar;
    const long long v0 = (e0 + 3) & ~3LL, v1 = e1 & ~3LL;              // 16-byte aligned body [v0, v1)
    if (v0 >= v1) {                                                    // (very short rows) scalars only
      for (long long e = e0 + threadIdx.x; e < e1; e += THREADS) {
        float s = reduce1<MC>(a, e);
        if (e == e1 - 1) s += prior;
        bcast1<MC>(a, e, s);
      }
      continue;
    }
    // head [e0, v0) and tail [v1, e1): at most 3 + 3 scalars, one thread each
    {
      const int nh = (int)(v0 - e0), nt = (int)(e1 - v1);
      if ((int)threadIdx.x < nh + nt) {
        const long long e = (int)threadIdx.x < nh ? e0 + threadIdx.x : v1 + (threadIdx.x - nh);
        float s = reduce1<MC>(a, e);
        if (e == e1 - 1) s += prior;
        bcast1<MC>(a, e, s);
      }
    }
    // body: UNROLL independent 16-byte requests in flight per thread
    const long long nvec = (v1 - v0) >> 2;
    for (long long i = threadIdx.x; i < nvec; i += (long long)THREADS * UNROLL) {
      float4 acc[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long j = i + (long long)u * THREADS;
        if (j < nvec) {
          const long long e = v0 + (j << 2);
          if (MC) acc[u] = mc_ldreduce4(a.mc + e);
          else acc[u] = sys_ld4(a.peer[0] + e);
        }
      }
      if (!MC) {
        for (int w = 1; w < W; ++w) {
          float4 t[UNROLL];
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            const long long j = i + (long long)u * THREADS;
            if (j < nvec) t[u] = sys_ld4(a.peer[w] + v0 + (j << 2));
          }
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            const long long j = i + (long long)u * THREADS;
            if (j < nvec) { acc[u].x += t[u].x; acc[u].y += t[u].y; acc[u].z += t[u].z; acc[u].w += t[u].w; }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long j = i + (long long)u * THREADS;
        if (j < nvec) {
          const long long e = v0 + (j << 2);
          const long long d = e1 - 1 - e;                                   // position of the diagonal element, if inside
          if (d == 0) acc[u].x += prior; else if (d == 1) acc[u].y += prior;
          else if (d == 2) acc[u].z += prior; else if (d == 3) acc[u].w += prior;
          if (MC) mc_st4(a.mc + e, acc[u]);
          else for (int w = 0; w < W; ++w) sys_st4(a.peer[w] + e, acc[u]);
        }
      }
    }
  }
}

// ---- packed lower triangle (the library-collective path: pack -> NCCL all-reduce -> unpack) ------------------------
// packed[(p(p+1)/2 - row0(row0+1)/2) + q] = H[p, q], q <= p, rows [row0, row1)
__global__ void pack_lower_kernel(const float* __restrict__ H, long long P, long long row0, long long row1,
                                  float* __restrict__ packed) {
  const long long base0 = row0 * (row0 + 1) / 2;
  for (long long p = row0 + blockIdx.x; p < row1; p += gridDim.x) {
    const float* src = H + p * P;
    float* dst = packed + (p * (p + 1) / 2 - base0);
    for (long long q = threadIdx.x; q <= p; q += blockDim.x) dst[q] = src[q];
  }
}
__global__ void unpack_lower_kernel(const float* __restrict__ packed, long long P, long long row0, long long row1,
                                    float* __restrict__ H, const float* __restrict__ prior_vec, float prior_scalar) {
  const long long base0 = row0 * (row0 + 1) / 2;
  for (long long p = row0 + blockIdx.x; p < row1; p += gridDim.x) {
    float* dst = H + p * P;
    const float* src = packed + (p * (p + 1) / 2 - base0);
    const float prior = prior_vec ? prior_vec[p] : prior_scalar;
    for (long long q = threadIdx.x; q <= p; q += blockDim.x) dst[q] = src[q] + (q == p ? prior : 0.f);
  }
}

// H[p, 0..p] = 0: the accumulation only ever adds into the lower triangle (plus the above-diagonal part of tiles that
// cross the diagonal, which the mirror overwrites), so a fresh accumulator needs half the bytes of a full memset.
__global__ void zero_lower_kernel(float* __restrict__ H, long long P) {
  for (long long p = blockIdx.x; p < P; p += gridDim.x) {
    float* row = H + p * P;
    const long long n = p + 1;
    const long long head = ((4 - ((reinterpret_cast<uintptr_t>(row) >> 2) & 3)) & 3);      // floats up to a 16-byte boundary
    const long long h = head < n ? head : n;
    if ((long long)threadIdx.x < h) row[threadIdx.x] = 0.f;
    const long long nv = (n - h) >> 2;
    float4* v = reinterpret_cast<float4*>(row + h);
    for (long long i = threadIdx.x; i < nv; i += blockDim.x) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long t0 = h + (nv << 2);
    if (t0 + (long long)threadIdx.x < n) row[t0 + threadIdx.x] = 0.f;
  }
}

}  // namespace xch
}  // namespace bnnp

using namespace bnnp;

extern "C" int bnnp_exchange_lower_allreduce(float* const* peers, float* multicast, int rank, int world, int64_t P,
                                             int64_t row0, int64_t row1, const float* prior_vec, float prior_scalar,
                                             int n_ctas, void* stream) {
  if (!peers || world < 1 || world > xch::MAX_WORLD || rank < 0 || rank >= world || P <= 0 || row0 < 0 || row1 > P ||
      row0 > row1)
    return BNNP_ERR_INVALID;
  if (row0 == row1) return BNNP_OK;
  xch::Args a;
  for (int w = 0; w < xch::MAX_WORLD; ++w) a.peer[w] = w < world ? peers[w] : nullptr;
  for (int w = 0; w < world; ++w)
    if (!a.peer[w]) return BNNP_ERR_INVALID;
  // vector path: every mapping must be 16-byte aligned (a symmetric allocation is)
  for (int w = 0; w < world; ++w)
    if (reinterpret_cast<uintptr_t>(a.peer[w]) & 15) return BNNP_ERR_INVALID;
  if (multicast && (reinterpret_cast<uintptr_t>(multicast) & 15)) return BNNP_ERR_INVALID;
  a.mc = multicast; a.prior_vec = prior_vec; a.prior_scalar = prior_scalar;
  a.rank = rank; a.world = world; a.P = P; a.row0 = row0; a.row1 = row1;
  if (n_ctas <= 0) n_ctas = multicast ? 16 : 64;      // measured at W = 8: NVLS 16 CTAs 17.0 ms, 32: 16.7, 128: 17.2
  const long long mine = (row1 - row0 + world - 1) / world;
  if (n_ctas > mine) n_ctas = (int)(mine > 0 ? mine : 1);
  if (multicast) xch::exchange_lower_kernel<true><<<n_ctas, xch::THREADS, 0, as_stream(stream)>>>(a);
  else xch::exchange_lower_kernel<false><<<n_ctas, xch::THREADS, 0, as_stream(stream)>>>(a);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" int64_t bnnp_packed_lower_floats(int64_t row0, int64_t row1) {
  return row1 * (row1 + 1) / 2 - row0 * (row0 + 1) / 2;
}

extern "C" int bnnp_pack_lower(const float* H, int64_t P, int64_t row0, int64_t row1, float* packed, void* stream) {
  if (!H || !packed || P <= 0 || row0 < 0 || row1 > P || row0 > row1) return BNNP_ERR_INVALID;
  if (row0 == row1) return BNNP_OK;
  const int64_t rows = row1 - row0;
  xch::pack_lower_kernel<<<(unsigned)(rows < 148 * 8 ? rows : 148 * 8), 512, 0, as_stream(stream)>>>(H, P, row0, row1, packed);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" int bnnp_unpack_lower(const float* packed, int64_t P, int64_t row0, int64_t row1, float* H,
                                 const float* prior_vec, float prior_scalar, void* stream) {
  if (!H || !packed || P <= 0 || row0 < 0 || row1 > P || row0 > row1) return BNNP_ERR_INVALID;
  if (row0 == row1) return BNNP_OK;
  const int64_t rows = row1 - row0;
  xch::unpack_lower_kernel<<<(unsigned)(rows < 148 * 8 ? rows : 148 * 8), 512, 0, as_stream(stream)>>>(
      packed, P, row0, row1, H, prior_vec, prior_scalar);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

extern "C" int bnnp_zero_lower(float* H, int64_t P, void* stream) {
  if (!H || P <= 0) return BNNP_ERR_INVALID;
  xch::zero_lower_kernel<<<(unsigned)(P < 148 * 16 ? P : 148 * 16), 256, 0, as_stream(stream)>>>(H, P);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}
