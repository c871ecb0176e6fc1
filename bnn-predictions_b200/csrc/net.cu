// Layer-factor pass: forward through the layer zoo saving every layer input, then one
// backward sweep that propagates V columns at once (all K output Jacobians, or the V columns of
// the loss-Hessian square root).  Replaces the K x (forward + BackPACK backward) loop of
// preds/gradients.py:24-36 and the BackPACK second-order backward of preds/laplace.py:55-58,73-76.
//
// Linear layers live in two concatenated matrices so that later kernels can synthesise
// Jacobian tiles on chip from one base pointer each:
//   A_cat [N, Dtot]   : input of every Linear (+ a constant-1 column for its bias)
//   G_cat [N*V, Mtot] : d f_v / d (pre-activation) of every Linear
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_tma.cuh"
#include "conv_bwd_tc.cuh"

namespace bnnp {

// ------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------
__global__ void act_fwd_kernel(float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                               int kind) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  float* p = x + (i / cols) * ld + (i % cols);
  float v = *p;
  if (kind == BNNP_OP_RELU) v = v > 0.f ? v : 0.f;
  else if (kind == BNNP_OP_TANH) v = tanhf(v);
  else v = 1.f / (1.f + expf(-v));
  *p = v;
}

// g[(n,v), c] *= act'(out[n, c])      grid: (ceil(cols/256), example slices): the derivative is computed once per example and
// applied to its V rows (a row loop with r / V costs a 64-bit division per element and an activation load per row)
__global__ void act_bwd_kernel(float* __restrict__ g, int64_t gld, const float* __restrict__ out,
                               int64_t old, int64_t R, int V, int64_t cols, int kind) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int64_t N = R / V;
  for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
    const float o = out[n * old + c];
    float d;
    if (kind == BNNP_OP_RELU) d = o > 0.f ? 1.f : 0.f;
    else if (kind == BNNP_OP_TANH) d = 1.f - o * o;
    else d = o * (1.f - o);
    float* p = g + n * V * gld + c;
    for (int v = 0; v < V; ++v, p += gld) *p *= d;
  }
}

__global__ void copy2d_kernel(float* __restrict__ dst, int64_t dld, const float* __restrict__ src,
                              int64_t sld, int64_t rows, int64_t cols) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  dst[(i / cols) * dld + (i % cols)] = src[(i / cols) * sld + (i % cols)];
}

__global__ void fill_col_kernel(float* __restrict__ dst, int64_t ld, int64_t rows, float v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) dst[i * ld] = v;
}

static int copy2d(float* dst, int64_t dld, const float* src, int64_t sld, int64_t rows,
                  int64_t cols, cudaStream_t st) {
  int64_t n = rows * cols;
  if (n == 0) return BNNP_OK;
  copy2d_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(dst, dld, src, sld, rows, cols);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}

// ------------------------------------------------------------------------------------------
// conv / pool (direct form; one thread per output element)
// ------------------------------------------------------------------------------------------
struct ConvGeom {
  int Cin, H, W, Cout, OH, OW, kh, kw, sh, sw, pl, pt;
};
static ConvGeom geom_of(const bnnp_op_t& o) {
  return ConvGeom{o.in_c, o.in_h, o.in_w, o.out_c, o.out_h, o.out_w, o.kh, o.kw,
                  o.stride_h, o.stride_w, o.pad_l, o.pad_t};
}

__global__ void conv_fwd_kernel(const float* __restrict__ in, int64_t ild, float* __restrict__ out,
                                int64_t old, const float* __restrict__ Wt,
                                const float* __restrict__ bias, int64_t N, ConvGeom g) {
  int64_t per = (int64_t)g.Cout * g.OH * g.OW;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * per) return;
  int64_t n = i / per, r = i % per;
  int co = (int)(r / (g.OH * g.OW)), oh = (int)((r / g.OW) % g.OH), ow = (int)(r % g.OW);
  const float* x = in + n * ild;
  const float* w = Wt + (int64_t)co * g.Cin * g.kh * g.kw;
  float acc = bias ? bias[co] : 0.f;
  for (int ci = 0; ci < g.Cin; ++ci)
    for (int a = 0; a < g.kh; ++a) {
      int ih = oh * g.sh - g.pt + a;
      if (ih < 0 || ih >= g.H) continue;
      for (int b = 0; b < g.kw; ++b) {
        int iw = ow * g.sw - g.pl + b;
        if (iw < 0 || iw >= g.W) continue;
        acc = fmaf(w[(ci * g.kh + a) * g.kw + b], x[((int64_t)ci * g.H + ih) * g.W + iw], acc);
      }
    }
  out[n * old + r] = acc;
}

// gin[r, ci, ih, iw] = sum_{co,a,b} W[co,ci,a,b] * gout[r, co, oh, ow],  oh*sh - pt + a = ih
__global__ void conv_bwd_data_kernel(const float* __restrict__ gout, int64_t gold,
                                     float* __restrict__ gin, int64_t gild,
                                     const float* __restrict__ Wt, int64_t R, ConvGeom g) {
  int64_t per = (int64_t)g.Cin * g.H * g.W;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * per) return;
  int64_t r = i / per, e = i % per;
  int ci = (int)(e / (g.H * g.W)), ih = (int)((e / g.W) % g.H), iw = (int)(e % g.W);
  const float* go = gout + r * gold;
  float acc = 0.f;
  for (int a = 0; a < g.kh; ++a) {
    int th = ih + g.pt - a;
    if (th < 0 || th % g.sh) continue;
    int oh = th / g.sh;
    if (oh >= g.OH) continue;
    for (int b = 0; b < g.kw; ++b) {
      int tw = iw + g.pl - b;
      if (tw < 0 || tw % g.sw) continue;
      int ow = tw / g.sw;
      if (ow >= g.OW) continue;
      for (int co = 0; co < g.Cout; ++co)
        acc = fmaf(Wt[(((int64_t)co * g.Cin + ci) * g.kh + a) * g.kw + b],
                   go[((int64_t)co * g.OH + oh) * g.OW + ow], acc);
    }
  }
  gin[r * gild + e] = acc;
}


// ---- convolution as implicit GEMM through the tiled functor GEMM -------------------------------
// forward:   out[n, co, pos] = bias[co] + sum_e W[co, e] * U_n[e, pos]          (batch = example n)
struct ConvWA {            // A(co, e) = W[co, e]
  static constexpr bool K_CONTIG = true;
  const float* W; int E;
  __device__ float operator()(int64_t co, int64_t e, int) const { return W[co * E + e]; }
};
struct ConvIm2colB {       // B(e, pos) = in[n, ci, oh*s - pt + a, ow*s - pl + b]  (0 outside)
  static constexpr bool K_CONTIG = false;
  const float* in; int64_t ild; ConvGeom g;
  __device__ float operator()(int64_t e, int64_t pos, int n) const {
    int b = (int)(e % g.kw), a = (int)((e / g.kw) % g.kh), ci = (int)(e / (g.kw * g.kh));
    int oh = (int)(pos / g.OW), ow = (int)(pos % g.OW);
    int ih = oh * g.sh - g.pt + a, iw = ow * g.sw - g.pl + b;
    if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) return 0.f;
    return in[(int64_t)n * ild + ((int64_t)ci * g.H + ih) * g.W + iw];
  }
};
struct ConvStoreOut {
  float* out; int64_t old; const float* bias; int L;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t co, int64_t pos, float acc, int n) const {
    out[(int64_t)n * old + co * L + pos] = acc + (bias ? bias[co] : 0.f);
  }
};
// backward data: gin[r, ci, ipos] = sum_{(co,a,b)} W[co,ci,a,b] * gout[r, co, oh, ow], oh*s - pt + a = ih
struct ConvWtA {           // A(ci, (co,a,b)) = W[co, ci, a, b]
  static constexpr bool K_CONTIG = false;
  const float* W; int Cin, kk;
  __device__ float operator()(int64_t ci, int64_t q, int) const {
    int64_t co = q / kk, ab = q % kk;
    return W[(co * Cin + ci) * kk + ab];
  }
};
struct ConvGoutB {         // B((co,a,b), ipos) = gout[r, co, oh, ow] if the tap hits a valid output
  static constexpr bool K_CONTIG = false;
  const float* g; int64_t gld; ConvGeom c;
  __device__ float operator()(int64_t q, int64_t ipos, int r) const {
    int kk = c.kh * c.kw;
    int co = (int)(q / kk), a = (int)((q % kk) / c.kw), b = (int)(q % c.kw);
    int ih = (int)(ipos / c.W), iw = (int)(ipos % c.W);
    int th = ih + c.pt - a, tw = iw + c.pl - b;
    if (th < 0 || tw < 0 || th % c.sh || tw % c.sw) return 0.f;
    int oh = th / c.sh, ow = tw / c.sw;
    if (oh >= c.OH || ow >= c.OW) return 0.f;
    return g[(int64_t)r * gld + ((int64_t)co * c.OH + oh) * c.OW + ow];
  }
};
struct ConvStoreGin {
  float* gin; int64_t gild; int HW;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t ci, int64_t ipos, float acc, int r) const {
    gin[(int64_t)r * gild + ci * HW + ipos] = acc;
  }
};

// Max pooling over a zero-padded input (deepobs tfmaxpool2d = ZeroPad2d + MaxPool2d(padding=0)):
// window scanned row-major, strict '>' so the FIRST maximum wins (torch's rule); idx = ih*W+iw
// of the winner or -1 when the winner is a padding zero.
__global__ void maxpool_fwd_kernel(const float* __restrict__ in, int64_t ild,
                                   float* __restrict__ out, int64_t old, int32_t* __restrict__ idx,
                                   int64_t N, ConvGeom g) {
  int64_t per = (int64_t)g.Cout * g.OH * g.OW;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * per) return;
  int64_t n = i / per, r = i % per;
  int c = (int)(r / (g.OH * g.OW)), oh = (int)((r / g.OW) % g.OH), ow = (int)(r % g.OW);
  const float* x = in + n * ild + (int64_t)c * g.H * g.W;
  float best = -INFINITY;
  int bi = -1;
  for (int a = 0; a < g.kh; ++a) {
    int ih = oh * g.sh - g.pt + a;
    for (int b = 0; b < g.kw; ++b) {
      int iw = ow * g.sw - g.pl + b;
      bool inside = ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
      float v = inside ? x[ih * g.W + iw] : 0.f;      // outside = explicit zero padding
      if (v > best || v != v) { best = v; bi = inside ? ih * g.W + iw : -1; }
    }
  }
  out[n * old + r] = best;
  idx[n * per + r] = bi;
}

// grid: (ceil(C*H*W/256), 1, z): the candidate windows of a position are found once and reused for all rows.
// FUSE != 0: the op in front of the pool is an elementwise activation whose (in-place) output `act` is the pool's
// input; its backward  g *= act'(out)  is applied to the value before the single store, which removes one full
// read-modify-write pass over the [R, C*H*W] gradient (the K-column backward of a conv net is HBM-bound).
template <int FUSE>
__global__ void maxpool_bwd_kernel(const float* __restrict__ gout, int64_t gold,
                                   float* __restrict__ gin, int64_t gild,
                                   const int32_t* __restrict__ idx, int64_t N, int V, ConvGeom g,
                                   const float* __restrict__ act, int64_t actld, int act_kind) {
  // one thread per input element (c, ih, iw), flattened over channels so that small feature maps (12 x 12, 6 x 6) fill
  // their blocks (one block row per channel left 44 % / 86 % of the threads idle there)
  const int HW = g.H * g.W;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.Cin * HW) return;
  const int c = j / HW, me = j - c * HW;
  const int ih = me / g.W, iw = me - ih * g.W;
  const int OL = g.OH * g.OW;
  const int64_t operp = (int64_t)g.Cout * OL;
  // the (at most 3 x 3) pooling windows that contain this input position: up to three output rows and three output
  // columns, kept in registers (compile-time indexed) -- a dynamically indexed candidate list lives in local memory and
  // costs a load per use in the row loop
  int ohs[3], ows[3];
  bool vh[3], vw[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int th = ih + g.pt - a, tw = iw + g.pl - a;
    vh[a] = a < g.kh && th >= 0 && th % g.sh == 0 && th / g.sh < g.OH;
    vw[a] = a < g.kw && tw >= 0 && tw % g.sw == 0 && tw / g.sw < g.OW;
    ohs[a] = vh[a] ? th / g.sh : 0;
    ows[a] = vw[a] ? tw / g.sw : 0;
  }
  // examples in the outer loop: the arg-max indices (and the activation derivative) are per example, shared by its V rows
  for (int64_t n = blockIdx.z; n < N; n += gridDim.z) {
    const int32_t* id = idx + n * operp + (int64_t)c * OL;
    unsigned hit = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        if (vh[a] && vw[b] && id[ohs[a] * g.OW + ows[b]] == me) hit |= 1u << (a * 3 + b);
    float d = 1.f;
    if (FUSE) {
      const float o = act[n * actld + (int64_t)c * HW + me];
      d = act_kind == BNNP_OP_RELU ? (o > 0.f ? 1.f : 0.f) : (act_kind == BNNP_OP_TANH ? 1.f - o * o : o * (1.f - o));
    }
    if (hit == 0) {                         // not the arg-max of any window (most positions): the V rows are zeros
      for (int v = 0; v < V; ++v) gin[(n * V + v) * gild + (int64_t)c * HW + me] = 0.f;
      continue;
    }
    for (int v = 0; v < V; ++v) {
      const int64_t r = n * V + v;
      const float* go = gout + r * gold + (int64_t)c * OL;
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
          if (hit & (1u << (a * 3 + b))) acc += go[ohs[a] * g.OW + ows[b]];
      gin[r * gild + (int64_t)c * HW + me] = FUSE ? acc * d : acc;
    }
  }
}

// The same for stride >= 2 and windows <= 3 x 3 (every position lies in at most 2 x 2 windows), four consecutive positions
// per thread and one 16-byte store per row (requires W % 4 == 0 and 16-byte aligned rows).  The general kernel is
// instruction-bound, not memory-bound (ncu on the [5120, 64, 28, 28] gradient of CIFAR10Net's first pool: 1.09e9 warp
// instructions of which 63 % are IMAD / LEA / IADD3, issue slots 63 % busy, DRAM 11 %): with overlapping windows a local maximum
// wins SEVERAL windows, so most threads walk all nine predicated (address, load, add) candidates for each of the V rows.  Here
// the at most four candidate windows of a position are compacted once per thread into register-resident offsets; the row loop
// is sixteen predicated loads on precomputed offsets and one store.
// PLANES != 0: the gradient leaves as the bf16 hi / lo planes [(r, c), pos] of the conv-weight diagonal kernel (conv_diag_tc.cu;
// row pitch H * W, lo plane `plane_elems` elements behind the hi plane, both inside the fp32 gradient buffer they replace)
// instead of fp32 -- the 1 GB gradient of CIFAR10Net's first conv was written here, re-read by the split pass and re-written as
// planes.
template <int FUSE, int PLANES = 0>
__global__ void maxpool_bwd4_kernel(const float* __restrict__ gout, int64_t gold,
                                    float* __restrict__ gin, int64_t gild,
                                    const int32_t* __restrict__ idx, int64_t N, int V, ConvGeom g,
                                    const float* __restrict__ act, int64_t actld, int act_kind, int64_t plane_elems = 0) {
  const int HW = g.H * g.W, W4 = g.W >> 2;
  const int j4 = blockIdx.x * blockDim.x + threadIdx.x;          // one thread per 4 consecutive positions of a row
  if (j4 >= g.Cin * g.H * W4) return;
  const int c = j4 / (g.H * W4), rem = j4 - c * (g.H * W4);
  const int ih = rem / W4, iw0 = (rem - ih * W4) << 2;
  const int me0 = ih * g.W + iw0;
  const int OL = g.OH * g.OW;
  const int64_t operp = (int64_t)g.Cout * OL;
  // candidate output rows of this input row (at most two for stride >= 2), as row offsets into the pooled map
  int orow[2] = {0, 0};
  int na = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int th = ih + g.pt - a;
    if (a < g.kh && th >= 0 && th % g.sh == 0 && th / g.sh < g.OH) {
      if (na == 0) orow[0] = (th / g.sh) * g.OW; else orow[1] = (th / g.sh) * g.OW;
      ++na;
    }
  }
  // candidate windows (i, j) of each of the four positions: offsets into the pooled map, -1 = none
  int off[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    int ocol[2] = {0, 0};
    int nb = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int tw = iw0 + p + g.pl - b;
      if (b < g.kw && tw >= 0 && tw % g.sw == 0 && tw / g.sw < g.OW) {
        if (nb == 0) ocol[0] = tw / g.sw; else ocol[1] = tw / g.sw;
        ++nb;
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) off[p][i * 2 + jj] = (i < na && jj < nb) ? orow[i] + ocol[jj] : -1;
  }
  for (int64_t n = blockIdx.z; n < N; n += gridDim.z) {
    const int32_t* id = idx + n * operp + (int64_t)c * OL;
    unsigned hit = 0;                                              // bit p * 4 + k: position p won its candidate window k
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (off[p][k] >= 0 && __ldg(id + off[p][k]) == me0 + p) hit |= 1u << (p * 4 + k);
    float* gp = gin + n * V * gild + (int64_t)c * HW + me0;
    __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(gin) + n * V * gild + (int64_t)c * HW + me0;
    if (hit == 0u) {                                               // four zeros per row
      if (PLANES) {
        for (int v = 0; v < V; ++v, hp += gild) {
          *reinterpret_cast<uint2*>(hp) = make_uint2(0u, 0u);
          *reinterpret_cast<uint2*>(hp + plane_elems) = make_uint2(0u, 0u);
        }
      } else {
        for (int v = 0; v < V; ++v, gp += gild) *reinterpret_cast<float4*>(gp) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      continue;
    }
    float d[4] = {1.f, 1.f, 1.f, 1.f};
    if (FUSE) {
      const float4 o4 = __ldg(reinterpret_cast<const float4*>(act + n * actld + (int64_t)c * HW + me0));
      const float o[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
        d[p] = act_kind == BNNP_OP_RELU ? (o[p] > 0.f ? 1.f : 0.f) : (act_kind == BNNP_OP_TANH ? 1.f - o[p] * o[p] : o[p] * (1.f - o[p]));
    }
    const float* go = gout + n * V * gold + (int64_t)c * OL;
    for (int v = 0; v < V; ++v, gp += gild, hp += gild, go += gold) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int p = 0; p < 4; ++p) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (hit & (1u << (p * 4 + k))) acc[p] += __ldg(go + off[p][k]);
        if (FUSE) acc[p] *= d[p];
      }
      if (PLANES) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[0], acc[1]), h1 = __floats2bfloat162_rn(acc[2], acc[3]);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(acc[0] - __low2float(h0), acc[1] - __high2float(h0));
        const __nv_bfloat162 l1 = __floats2bfloat162_rn(acc[2] - __low2float(h1), acc[3] - __high2float(h1));
        *reinterpret_cast<uint2*>(hp) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(hp + plane_elems) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
      } else {
        *reinterpret_cast<float4*>(gp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      }
    }
  }
}

// The same with TWO positions per thread and 8-byte stores for maps whose width is even but not a multiple of four (the 6 x 6
// maps of CIFAR10Net's last pool ran on the general nine-candidate kernel: 162 us for 94 MB).
template <int FUSE>
__global__ void maxpool_bwd2_kernel(const float* __restrict__ gout, int64_t gold,
                                    float* __restrict__ gin, int64_t gild,
                                    const int32_t* __restrict__ idx, int64_t N, int V, ConvGeom g,
                                    const float* __restrict__ act, int64_t actld, int act_kind) {
  const int HW = g.H * g.W, W2 = g.W >> 1;
  const int j2 = blockIdx.x * blockDim.x + threadIdx.x;          // one thread per 2 consecutive positions of a row
  if (j2 >= g.Cin * g.H * W2) return;
  const int c = j2 / (g.H * W2), rem = j2 - c * (g.H * W2);
  const int ih = rem / W2, iw0 = (rem - ih * W2) << 1;
  const int me0 = ih * g.W + iw0;
  const int OL = g.OH * g.OW;
  const int64_t operp = (int64_t)g.Cout * OL;
  int orow[2] = {0, 0};
  int na = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int th = ih + g.pt - a;
    if (a < g.kh && th >= 0 && th % g.sh == 0 && th / g.sh < g.OH) {
      if (na == 0) orow[0] = (th / g.sh) * g.OW; else orow[1] = (th / g.sh) * g.OW;
      ++na;
    }
  }
  int off[2][4];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    int ocol[2] = {0, 0};
    int nb = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int tw = iw0 + p + g.pl - b;
      if (b < g.kw && tw >= 0 && tw % g.sw == 0 && tw / g.sw < g.OW) {
        if (nb == 0) ocol[0] = tw / g.sw; else ocol[1] = tw / g.sw;
        ++nb;
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) off[p][i * 2 + jj] = (i < na && jj < nb) ? orow[i] + ocol[jj] : -1;
  }
  for (int64_t n = blockIdx.z; n < N; n += gridDim.z) {
    const int32_t* id = idx + n * operp + (int64_t)c * OL;
    unsigned hit = 0;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (off[p][k] >= 0 && __ldg(id + off[p][k]) == me0 + p) hit |= 1u << (p * 4 + k);
    float* gp = gin + n * V * gild + (int64_t)c * HW + me0;
    if (hit == 0u) {
      for (int v = 0; v < V; ++v, gp += gild) *reinterpret_cast<float2*>(gp) = make_float2(0.f, 0.f);
      continue;
    }
    float d[2] = {1.f, 1.f};
    if (FUSE) {
      const float2 o2 = __ldg(reinterpret_cast<const float2*>(act + n * actld + (int64_t)c * HW + me0));
      const float o[2] = {o2.x, o2.y};
#pragma unroll
      for (int p = 0; p < 2; ++p)
        d[p] = act_kind == BNNP_OP_RELU ? (o[p] > 0.f ? 1.f : 0.f) : (act_kind == BNNP_OP_TANH ? 1.f - o[p] * o[p] : o[p] * (1.f - o[p]));
    }
    const float* go = gout + n * V * gold + (int64_t)c * OL;
    for (int v = 0; v < V; ++v, gp += gild, go += gold) {
      float acc[2] = {0.f, 0.f};
#pragma unroll
      for (int p = 0; p < 2; ++p) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (hit & (1u << (p * 4 + k))) acc[p] += __ldg(go + off[p][k]);
        if (FUSE) acc[p] *= d[p];
      }
      *reinterpret_cast<float2*>(gp) = make_float2(acc[0], acc[1]);
    }
  }
}

// KFAC needs only the POSITION SUMS of the gradient at a conv layer's output (BackPACK KFLR: G = sum (sum_pos g)(sum_pos
// g)^T, preds/laplace.py:73-78).  For the first parameterised layer nothing upstream needs the full gradient either, so
// when it is followed by (activation,) max-pool the [R, C, H, W] gradient (2 MB per example at CIFAR10Net's conv1 and
// the single most expensive pass of the K-column backward) is never materialised: a max-pool routes each pooled
// gradient to exactly one input position (or to a padding zero), hence
//   S[r, c] = sum_pos g_conv[r, c, pos] = sum_windows [winner inside] * act'(act[n, c, winner]) * g_pool[r, c, window].
// One warp per (r, c).
__global__ void maxpool_bwd_possum_kernel(const float* __restrict__ gout, int64_t gold, float* __restrict__ S,
                                          const int32_t* __restrict__ idx, int64_t N, int V, ConvGeom g,
                                          const float* __restrict__ act, int64_t actld, int act_kind) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int C = g.Cout, OL = g.OH * g.OW, HW = g.H * g.W;
  if (w >= N * V * C) return;
  const int c = (int)(w % C);
  const int64_t r = w / C, n = r / V;
  const float* go = gout + r * gold + (int64_t)c * OL;
  const int32_t* id = idx + (n * C + c) * (int64_t)OL;
  const float* a = act ? act + n * actld + (int64_t)c * HW : nullptr;
  float acc = 0.f;
  for (int i = lane; i < OL; i += 32) {
    const int win = id[i];
    if (win >= 0) {
      float d = 1.f;
      if (a) {
        const float o = a[win];
        d = act_kind == BNNP_OP_RELU ? (o > 0.f ? 1.f : 0.f) : (act_kind == BNNP_OP_TANH ? 1.f - o * o : o * (1.f - o));
      }
      acc = fmaf(go[i], d, acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) S[r * C + c] = acc;
}

// The same with one warp per (example, channel) for pooled maps of at most 256 positions: the winner index and the activation
// derivative of a window depend on the example only, so they are gathered ONCE into registers and the V rows of the example are
// streamed against them (the per-row kernel repeats the dependent index -> activation gathers for each of the V rows).
__global__ void maxpool_bwd_possum_rows_kernel(const float* __restrict__ gout, int64_t gold, float* __restrict__ S,
                                               const int32_t* __restrict__ idx, int64_t N, int V, ConvGeom g,
                                               const float* __restrict__ act, int64_t actld, int act_kind) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int C = g.Cout, OL = g.OH * g.OW, HW = g.H * g.W;
  if (w >= N * C) return;
  const int c = (int)(w % C);
  const int64_t n = w / C;
  const int32_t* id = idx + (n * C + c) * (int64_t)OL;
  const float* a = act ? act + n * actld + (int64_t)c * HW : nullptr;
  float d[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = lane + 32 * k;
    const int win = i < OL ? __ldg(id + i) : -1;
    d[k] = 0.f;
    if (win >= 0) {
      d[k] = 1.f;
      if (a) {
        const float o = __ldg(a + win);
        d[k] = act_kind == BNNP_OP_RELU ? (o > 0.f ? 1.f : 0.f) : (act_kind == BNNP_OP_TANH ? 1.f - o * o : o * (1.f - o));
      }
    }
  }
  for (int v = 0; v < V; ++v) {
    const int64_t r = n * V + v;
    const float* go = gout + r * gold + (int64_t)c * OL;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = lane + 32 * k;
      if (i < OL) acc = fmaf(__ldg(go + i), d[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) S[r * C + c] = acc;
  }
}

__global__ void avgpool_fwd_kernel(const float* __restrict__ in, int64_t ild,
                                   float* __restrict__ out, int64_t old, int64_t N, ConvGeom g) {
  int64_t per = (int64_t)g.Cout * g.OH * g.OW;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * per) return;
  int64_t n = i / per, r = i % per;
  int c = (int)(r / (g.OH * g.OW)), oh = (int)((r / g.OW) % g.OH), ow = (int)(r % g.OW);
  const float* x = in + n * ild + (int64_t)c * g.H * g.W;
  float acc = 0.f;
  for (int a = 0; a < g.kh; ++a)
    for (int b = 0; b < g.kw; ++b) {
      int ih = oh * g.sh - g.pt + a, iw = ow * g.sw - g.pl + b;
      if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) acc += x[ih * g.W + iw];
    }
  out[n * old + r] = acc / (float)(g.kh * g.kw);
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ gout, int64_t gold,
                                   float* __restrict__ gin, int64_t gild, int64_t R, ConvGeom g) {
  int64_t per = (int64_t)g.Cin * g.H * g.W;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * per) return;
  int64_t r = i / per, e = i % per;
  int c = (int)(e / (g.H * g.W)), ih = (int)((e / g.W) % g.H), iw = (int)(e % g.W);
  const float* go = gout + r * gold + (int64_t)c * g.OH * g.OW;
  float acc = 0.f;
  for (int a = 0; a < g.kh; ++a) {
    int th = ih + g.pt - a;
    if (th < 0 || th % g.sh) continue;
    int oh = th / g.sh;
    if (oh >= g.OH) continue;
    for (int b = 0; b < g.kw; ++b) {
      int tw = iw + g.pl - b;
      if (tw < 0 || tw % g.sw) continue;
      int ow = tw / g.sw;
      if (ow >= g.OW) continue;
      acc += go[oh * g.OW + ow];
    }
  }
  gin[r * gild + e] = acc / (float)(g.kh * g.kw);
}

// Linear forward epilogue: out = acc + bias
struct StoreBias {
  float* p; int64_t ld; const float* bias;
  __device__ bool skip(int64_t, int64_t, int, int) const { return false; }
  __device__ void operator()(int64_t m, int64_t n, float acc, int) const {
    p[m * ld + n] = acc + (bias ? bias[n] : 0.f);
  }
};

// ---- convolution backward-data on tensor cores ----------------------------------------------------
//   dU[(r,pos), e] = sum_co gout[r, co, pos] * W[co, e]          (plain GEMM, K = Cout)
//   gin[r, ci, ih, iw] = sum_{a,b} dU[(r, pos(ih,iw,a,b)), (ci,a,b)]   (col2im as a gather)
// dU is stored transposed, [r, e, pos], so that both its write and the gather are coalesced.
struct ConvGoutRowsA {     // A(m = (r,pos), k = co) = gout[r, co, pos]
  static constexpr bool K_CONTIG = false;
  using Row = const float*;
  const float* g; long long gld; int L;
  __device__ __forceinline__ Row row(int64_t m, int) const { return g + (m / L) * gld + (m % L); }
  __device__ __forceinline__ void load8(const Row& rw, int64_t k0, int64_t K, float* out) const {
#pragma unroll
    for (int e = 0; e < 8; ++e) out[e] = (k0 + e < K) ? __ldg(rw + (k0 + e) * L) : 0.f;
  }
};
struct ConvStoreDUt : tcg::TcNoState {    // dUt[(r*E + e)*L + pos] = acc(m = (r,pos), e)
  float* d; int E, L;
  __host__ __device__ ConvStoreDUt(float* d_, int E_, int L_) : d(d_), E(E_), L(L_) {}
  __device__ __forceinline__ void row(State&, int64_t m, int64_t e0, const float* v, int nvalid, int, int) const {
    float* q = d + ((m / L) * E + e0) * L + (m % L);
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e < nvalid) q[(long long)e * L] = v[e];
  }
};
// One thread per input element (ci, ih, iw) -- flattened, so that small feature maps (6 x 6, 14 x 14) fill their blocks --
// with the (at most KH x KW) contributing taps resolved once into registers (compile-time indexed: a dynamically indexed
// offset list lives in local memory and costs a load per tap and row); rows r in the outer loop.
template <int KH, int KW>
__global__ void col2im_gather_kernel(const float* __restrict__ dUt, float* __restrict__ gin, int64_t gild,
                                     int64_t R, ConvGeom g) {
  const int HW = g.H * g.W;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;       // grid: (ceil(Cin*H*W/256), row slices)
  if (j >= g.Cin * HW) return;
  const int ci = j / HW, me = j - ci * HW;
  const int ih = me / g.W, iw = me - ih * g.W;
  const int L = g.OH * g.OW;
  int off[KH * KW];
  bool ok[KH * KW];
#pragma unroll
  for (int a = 0; a < KH; ++a) {
    const int th = ih + g.pt - a;
    const bool vh = a < g.kh && th >= 0 && th % g.sh == 0 && th / g.sh < g.OH;
#pragma unroll
    for (int b = 0; b < KW; ++b) {
      const int tw = iw + g.pl - b;
      const bool vw = b < g.kw && tw >= 0 && tw % g.sw == 0 && tw / g.sw < g.OW;
      ok[a * KW + b] = vh && vw;
      off[a * KW + b] = vh && vw ? (a * g.kw + b) * L + (th / g.sh) * g.OW + tw / g.sw : 0;
    }
  }
  const int64_t per_row = (int64_t)g.Cin * g.kh * g.kw * L;
  const float* base0 = dUt + (int64_t)ci * g.kh * g.kw * L;
  for (int64_t r = blockIdx.y; r < R; r += gridDim.y) {
    const float* base = base0 + r * per_row;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < KH * KW; ++t)
      if (ok[t]) acc += base[off[t]];
    gin[r * gild + j] = acc;
  }
}

// ---- convolution forward on tensor cores: out[(n,pos), co] = sum_e U[(n,pos), e] W[co, e] + bias ----
// grid (ceil(L / 16), N), 256 threads along e: (ci, a, b) of e are decomposed once per thread, the 16 positions of the block
// walked incrementally (a flat one-thread-per-element index needs eight integer divisions per element, three of them
// 64-bit: the pass was compute-bound at 0.5 TB/s of stores); stores are e-contiguous per position
__global__ void im2col_fwd_kernel(const float* __restrict__ in, int64_t ild, int64_t N, ConvGeom g, int E, int Epad,
                                  float* __restrict__ U) {
  const int L = g.OH * g.OW;
  const int64_t n = blockIdx.y;
  const int pos0 = blockIdx.x * 16;
  const int npos = L - pos0 < 16 ? L - pos0 : 16;
  const int oh0 = pos0 / g.OW, ow0 = pos0 - oh0 * g.OW;
  const float* xn = in + n * ild;
  float* Un = U + (n * L + pos0) * (int64_t)Epad;
  for (int e = threadIdx.x; e < Epad; e += blockDim.x) {
    const int b = e % g.kw, t = e / g.kw, a = t % g.kh, ci = t / g.kh;
    const float* xc = xn + (int64_t)ci * g.H * g.W;
    int oh = oh0, ow = ow0;
    for (int p = 0; p < npos; ++p) {
      float v = 0.f;
      if (e < E) {
        const int ih = oh * g.sh - g.pt + a, iw = ow * g.sw - g.pl + b;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) v = __ldg(xc + ih * g.W + iw);
      }
      Un[(int64_t)p * Epad + e] = v;
      if (++ow == g.OW) { ow = 0; ++oh; }
    }
  }
}
struct ConvStoreOutT : tcg::TcNoState {   // out[n*old + co*L + pos] = acc(m = (n,pos), co) + bias[co]
  float* out; long long old; const float* bias; int L;
  __host__ __device__ ConvStoreOutT(float* o_, long long old_, const float* b_, int L_) : out(o_), old(old_), bias(b_), L(L_) {}
  __device__ __forceinline__ void row(State&, int64_t m, int64_t c0, const float* v, int nvalid, int, int) const {
    float* q = out + (m / L) * old + c0 * L + (m % L);
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e < nvalid) q[(long long)e * L] = v[e] + (bias ? __ldg(bias + c0 + e) : 0.f);
  }
};

// tensor-core epilogue of the Linear forward: out = acc + bias
struct TcStoreBias : tcg::TcNoState {
  float* p; long long ld; const float* bias;
  __host__ __device__ TcStoreBias(float* p_, long long ld_, const float* b_) : p(p_), ld(ld_), bias(b_) {}
  __device__ __forceinline__ void row(State&, int64_t m, int64_t n0, const float* v, int nvalid, int, int) const {
    float* q = p + m * ld + n0;
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e < nvalid) q[e] = v[e] + (bias ? __ldg(bias + n0 + e) : 0.f);
  }
};
// contractions below this many MACs stay on the SIMT path (a 128 x 256 tensor-core tile would be mostly padding)
constexpr long long kTcMinMacs = 1LL << 25;

}  // namespace bnnp

using namespace bnnp;

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
static int next_param_op(const bnnp_op_t* ops, int n_ops, int i) {
  for (int j = i + 1; j < n_ops; ++j) {
    if (is_elementwise_op(ops[j]) || ops[j].kind == BNNP_OP_FLATTEN) continue;
    return j;
  }
  return -1;
}

extern "C" int bnnp_net_plan(const bnnp_op_t* ops, int n_ops, int64_t N, int V, bnnp_plan_t* plan) {
  if (!ops || !plan || n_ops <= 0 || n_ops > BNNP_MAX_OPS || N <= 0 || V <= 0) return BNNP_ERR_INVALID;
  bnnp_plan_t P = {};
  P.first_linear = -1;
  int64_t nparam = 0;
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    if (o.kind < BNNP_OP_LINEAR || o.kind > BNNP_OP_SIGMOID) return BNNP_ERR_UNSUPPORTED;
    if (i > 0 && op_in_size(o) != op_out_size(ops[i - 1])) return BNNP_ERR_INVALID;
    if ((is_elementwise_op(o) || o.kind == BNNP_OP_FLATTEN)) {
      if (i == 0) return BNNP_ERR_UNSUPPORTED;
      if (op_in_size(o) != op_out_size(o)) return BNNP_ERR_INVALID;
    }
    if (is_param_op(o)) {
      if (!o.weight) return BNNP_ERR_INVALID;
      int64_t wn = (int64_t)o.out_c * o.in_c * o.kh * o.kw;
      nparam += wn + (o.has_bias ? o.out_c : 0);
    }
    if (o.kind == BNNP_OP_LINEAR) {
      if (o.in_h != 1 || o.in_w != 1 || o.kh != 1 || o.kw != 1) return BNNP_ERR_INVALID;
      if (P.first_linear < 0) P.first_linear = i;
      // every layer slice starts on a 16-byte boundary (vectorised operand loads); padding columns are zero
      P.op[i].lin_col = P.Dtot;
      P.op[i].lin_unit = P.Mtot;
      P.Dtot += (o.in_c + 1 + 3) / 4 * 4;
      P.Mtot += (o.out_c + 3) / 4 * 4;
      P.n_linear++;
    } else {
      P.op[i].lin_col = P.op[i].lin_unit = -1;
    }
  }
  P.P = nparam;
  P.K = (int32_t)op_out_size(ops[n_ops - 1]);
  int64_t cur = 0;
  auto take = [&](int64_t n) { int64_t o = cur; cur += (n + 31) / 32 * 32; return o; };
  P.acat_off = take(N * (int64_t)P.Dtot);
  P.gcat_off = take(N * V * (int64_t)P.Mtot);
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    bnnp_op_plan_t& q = P.op[i];
    q.idx_off = -1;
    if (is_elementwise_op(o) || o.kind == BNNP_OP_FLATTEN) {
      q.act_off = P.op[i - 1].act_off; q.act_ld = P.op[i - 1].act_ld;
      q.grad_off = P.op[i - 1].grad_off; q.grad_ld = P.op[i - 1].grad_ld;
      continue;
    }
    int64_t osz = op_out_size(o);
    int nx = next_param_op(ops, n_ops, i);
    if (o.kind == BNNP_OP_LINEAR && nx >= 0 && ops[nx].kind == BNNP_OP_LINEAR) {
      q.act_off = P.acat_off + P.op[nx].lin_col; q.act_ld = P.Dtot;
    } else {
      q.act_off = take(N * osz); q.act_ld = osz;
    }
    if (o.kind == BNNP_OP_LINEAR) {
      q.grad_off = P.gcat_off + q.lin_unit; q.grad_ld = P.Mtot;
    } else {
      q.grad_off = take(N * V * osz); q.grad_ld = osz;
    }
    if (o.kind == BNNP_OP_MAXPOOL2D) q.idx_off = take(N * osz);
  }
  // scratch of the tensor-core conv paths: the unfolded input of the forward GEMM and the [N*V, E, L] buffer
  // of the backward-data GEMM (every conv but the first parameterised op propagates a gradient)
  P.conv_scratch_off = -1;
  {
    int first_param = -1;
    for (int i = 0; i < n_ops; ++i) if (is_param_op(ops[i])) { first_param = i; break; }
    int64_t need = 0;
    for (int i = 0; i < n_ops; ++i)
      if (ops[i].kind == BNNP_OP_CONV2D) {
        const int64_t E = (int64_t)ops[i].in_c * ops[i].kh * ops[i].kw, L = (int64_t)ops[i].out_h * ops[i].out_w;
        int64_t v = N * L * ((E + 3) / 4 * 4);                    // forward: unfolded input [N*L, Epad]
        // backward-data: dU^T [N*V, E, L] + the transposed bf16 hi/lo gradients and weights of the TMA-fed product
        const int64_t bwd = (N * V * E * L + 63) / 64 * 64 + conv_bwd_tma_scratch_floats(N * V, ops[i].out_c, L, E);
        if (i > first_param && bwd > v) v = bwd;
        const int64_t bwi = conv_bwd_implicit_scratch_floats(N * V, ops[i].out_c, (int)L, ops[i].in_c, ops[i].kh * ops[i].kw) + 64;
        if (i > first_param && bwi > v) v = bwi;
        // forward as the same implicit GEMM: planes of the layer input [N, in_h * in_w, C_in] and of the mirrored weights
        const int64_t fwi = conv_bwd_implicit_scratch_floats(N, ops[i].in_c, ops[i].in_h * ops[i].in_w, ops[i].out_c,
                                                             ops[i].kh * ops[i].kw) + 64;
        if (fwi > v) v = fwi;
        need = need > v ? need : v;
      }
    // bf16 hi/lo operand copies of the TMA-fed Linear-layer products (forward over N rows, backward over N*V rows)
    for (int i = 0; i < n_ops; ++i)
      if (ops[i].kind == BNNP_OP_LINEAR) {
        int64_t f = linear_tma_scratch_floats(N, ops[i].in_c, ops[i].out_c);
        int64_t b = linear_tma_scratch_floats(N * V, ops[i].out_c, ops[i].in_c);
        need = need > f ? need : f;
        if (i > first_param) need = need > b ? need : b;
      }
    if (need > 0) P.conv_scratch_off = take(need);
  }
  P.total_floats = cur;
  *plan = P;
  return BNNP_OK;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
extern "C" int bnnp_net_forward(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                const float* X, int64_t N, float* ws, float* logits, void* stream) {
  if (!ops || !plan || !X || !ws || N <= 0) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  if (plan->Dtot > 0)
    BNNP_CUDA_CHECK(cudaMemsetAsync(ws + plan->acat_off, 0, sizeof(float) * N * plan->Dtot, st));
  const float* in = X;
  int64_t ild = op_in_size(ops[0]);
  for (int i = 0; i < n_ops; ++i) {
    const bnnp_op_t& o = ops[i];
    const bnnp_op_plan_t& q = plan->op[i];
    float* out = ws + q.act_off;
    int64_t osz = op_out_size(o);
    unsigned blocks = (unsigned)ceil_div64(N * osz, 256);
    switch (o.kind) {
      case BNNP_OP_LINEAR: {
        float* slice = ws + plan->acat_off + q.lin_col;
        if (in != slice) {
          int rc = copy2d(slice, plan->Dtot, in, ild, N, o.in_c, st);
          if (rc) return rc;
        }
        fill_col_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, st>>>(slice + o.in_c, plan->Dtot, N, 1.f);
        BNNP_LAUNCH_CHECK();
        int rc;
        if (!g_bnnp_opt.net_simt && (long long)N * o.out_c * o.in_c >= kTcMinMacs && plan->conv_scratch_off >= 0 &&
            o.in_c <= 4096 && g_bnnp_opt.gemm_generated == 0)
          // TMA-fed tensor-core product on bf16 hi/lo copies of the activations and the weights (gemm_tma.cu)
          rc = linear_tma(slice, plan->Dtot, N, o.in_c, o.weight, o.in_c, o.out_c, 0, o.has_bias ? o.bias : nullptr, out,
                          q.act_ld, ws + plan->conv_scratch_off, st);
        else if (!g_bnnp_opt.net_simt && (long long)N * o.out_c * o.in_c >= kTcMinMacs)
          rc = tcg::gemm_tc(N, o.out_c, o.in_c, 1, tcg::TcRowMajor{slice, plan->Dtot, 0},
                            tcg::TcRowMajor{o.weight, o.in_c, 0},
                            TcStoreBias(out, q.act_ld, o.has_bias ? o.bias : nullptr), 1, st);
        else
          rc = gemm_simt(N, o.out_c, o.in_c, 1, LoadRowMajorA{slice, plan->Dtot, 0},
                         LoadColMajorB{o.weight, o.in_c, 0},
                         StoreBias{out, q.act_ld, o.has_bias ? o.bias : nullptr}, st);
        if (rc) return rc;
        break;
      }
      case BNNP_OP_CONV2D: {
        const int E = o.in_c * o.kh * o.kw, L = o.out_h * o.out_w;
        if (plan->conv_scratch_off >= 0 && !g_bnnp_opt.net_simt && (long long)N * L * E * o.out_c >= kTcMinMacs &&
            g_bnnp_opt.gemm_generated == 0 && g_bnnp_opt.conv_bwd_explicit == 0 && o.in_c >= 16 &&
            conv_bwd_implicit_ok(o.out_c, o.in_c, o.out_h, o.out_w, o.in_h, o.in_w, o.kh, o.kw, o.stride_h, o.stride_w, N)) {
          // implicit GEMM on 4-D TMA boxes of the NHWC bf16 hi / lo input (conv_bwd_tc.cu with the channel roles swapped and the
          // taps mirrored): no unfolded input.  Not for a few input channels (an RGB first layer): the kernel's reduction axis is the
          // channel axis in chunks of 32 per tap, so C_in = 3 would run at 10 % useful MMA work (measured slower than the unfold).
          int rc = conv_bwd_implicit(in, ild, N, o.out_c, o.in_c, o.out_h, o.out_w, o.in_h, o.in_w, o.kh, o.kw, o.kw - 1 - o.pad_l,
                                     o.kh - 1 - o.pad_t, o.weight, out, q.act_ld, ws + plan->conv_scratch_off, st, 1,
                                     o.has_bias ? o.bias : nullptr);
          if (rc) return rc;
        } else if (plan->conv_scratch_off >= 0 && !g_bnnp_opt.net_simt && (long long)N * L * E * o.out_c >= kTcMinMacs && N <= 65535) {
          const int Epad = (E + 3) / 4 * 4;
          float* U = ws + plan->conv_scratch_off;
          im2col_fwd_kernel<<<dim3((unsigned)ceil_div64(L, 16), (unsigned)N), 256, 0, st>>>(in, ild, N, geom_of(o), E, Epad, U);
          BNNP_LAUNCH_CHECK();
          int rc = tcg::gemm_tc(N * L, o.out_c, E, 1, tcg::TcRowMajor{U, Epad, 0}, tcg::TcRowMajor{o.weight, E, 0},
                                ConvStoreOutT(out, q.act_ld, o.has_bias ? o.bias : nullptr, L), 1, st);
          if (rc) return rc;
        } else if (N <= 65535) {
          int rc = gemm_simt(o.out_c, L, E, (int)N, ConvWA{o.weight, E}, ConvIm2colB{in, ild, geom_of(o)},
                             ConvStoreOut{out, q.act_ld, o.has_bias ? o.bias : nullptr, L}, st);
          if (rc) return rc;
        } else {
          conv_fwd_kernel<<<blocks, 256, 0, st>>>(in, ild, out, q.act_ld, o.weight,
                                                  o.has_bias ? o.bias : nullptr, N, geom_of(o));
          BNNP_LAUNCH_CHECK();
        }
        break;
      }
      case BNNP_OP_MAXPOOL2D:
        maxpool_fwd_kernel<<<blocks, 256, 0, st>>>(in, ild, out, q.act_ld,
                                                   reinterpret_cast<int32_t*>(ws) + q.idx_off, N, geom_of(o));
        BNNP_LAUNCH_CHECK();
        break;
      case BNNP_OP_AVGPOOL2D:
        avgpool_fwd_kernel<<<blocks, 256, 0, st>>>(in, ild, out, q.act_ld, N, geom_of(o));
        BNNP_LAUNCH_CHECK();
        break;
      case BNNP_OP_RELU: case BNNP_OP_TANH: case BNNP_OP_SIGMOID:
        act_fwd_kernel<<<blocks, 256, 0, st>>>(out, N, osz, q.act_ld, o.kind);
        BNNP_LAUNCH_CHECK();
        break;
      case BNNP_OP_FLATTEN:
        break;
      default:
        return BNNP_ERR_UNSUPPORTED;
    }
    in = out;
    ild = q.act_ld;
  }
  if (logits) {
    int rc = copy2d(logits, plan->K, in, ild, N, plan->K, st);
    if (rc) return rc;
  }
  return BNNP_OK;
}

// ------------------------------------------------------------------------------------------
// backward (V columns)
// ------------------------------------------------------------------------------------------
extern "C" int bnnp_net_backward(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                 const float* seed, int64_t N, int V, float* ws, void* stream) {
  return bnnp_net_backward_ex(ops, n_ops, plan, seed, N, V, ws, 0, stream);
}

// 1 when BNNP_BWD_FIRST_CONV_POSSUM applies to this network: first parameterised op is a conv followed by
// (an optional elementwise activation and) a max-pool
extern "C" int bnnp_net_first_conv_possum_ok(const bnnp_op_t* ops, int n_ops) {
  if (!ops) return 0;
  int fp = -1;
  for (int i = 0; i < n_ops; ++i) if (is_param_op(ops[i])) { fp = i; break; }
  if (fp < 0 || ops[fp].kind != BNNP_OP_CONV2D) return 0;
  int j = fp + 1;
  if (j < n_ops && is_elementwise_op(ops[j])) ++j;
  return j < n_ops && ops[j].kind == BNNP_OP_MAXPOOL2D && ops[j].kh <= 3 && ops[j].kw <= 3 ? 1 : 0;
}
// index of the max-pool behind the first conv (conv [-> activation] -> max-pool), its activation flag
static int first_conv_pool(const bnnp_op_t* ops, int n_ops, int* first_param, bool* act_before) {
  int fp = -1;
  for (int i = 0; i < n_ops; ++i) if (is_param_op(ops[i])) { fp = i; break; }
  if (fp < 0 || ops[fp].kind != BNNP_OP_CONV2D) return -1;
  int j = fp + 1;
  const bool act = j < n_ops && is_elementwise_op(ops[j]);
  if (act) ++j;
  if (j >= n_ops || ops[j].kind != BNNP_OP_MAXPOOL2D || ops[j].kh > 3 || ops[j].kw > 3) return -1;
  *first_param = fp; *act_before = act;
  return j;
}
// S[r, c] = sum_pos of the first conv's output gradient, straight from the pooled gradient (maxpool_bwd_possum kernels)
int net_first_conv_possum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, int V, const float* ws,
                          float* S, cudaStream_t st) {
  int fp; bool act_before;
  const int j = first_conv_pool(ops, n_ops, &fp, &act_before);
  if (j < 0) return BNNP_ERR_INVALID;
  const bnnp_op_t& o = ops[j];
  const bnnp_op_plan_t& q = plan->op[j];
  const bnnp_op_plan_t& qp = plan->op[j - 1];
  const int64_t R = N * V;
  const int32_t* idxp = reinterpret_cast<const int32_t*>(ws) + q.idx_off;
  const float* gout = ws + q.grad_off;
  if (o.out_h * o.out_w <= 256)
    maxpool_bwd_possum_rows_kernel<<<(unsigned)ceil_div64(N * o.in_c * 32, 256), 256, 0, st>>>(
        gout, q.grad_ld, S, idxp, N, V, geom_of(o), act_before ? ws + qp.act_off : nullptr, qp.act_ld,
        act_before ? ops[j - 1].kind : 0);
  else
    maxpool_bwd_possum_kernel<<<(unsigned)ceil_div64(R * o.in_c * 32, 256), 256, 0, st>>>(
        gout, q.grad_ld, S, idxp, N, V, geom_of(o), act_before ? ws + qp.act_off : nullptr, qp.act_ld,
        act_before ? ops[j - 1].kind : 0);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}
// structural half of bnnp_net_first_conv_planes_ok (ggn.cu adds the conditions of the diag kernel)
bool net_first_conv_planes_structure_ok(const bnnp_op_t* ops, int n_ops) {
  int fp; bool act_before;
  const int j = first_conv_pool(ops, n_ops, &fp, &act_before);
  if (j < 0) return false;
  const bnnp_op_t& o = ops[j];
  const int64_t isz = (int64_t)o.in_c * o.in_h * o.in_w;
  return o.in_w % 4 == 0 && o.stride_h >= 2 && o.stride_w >= 2 && isz % 4 == 0 && (o.in_h * o.in_w) % 8 == 0;
}

extern "C" int bnnp_net_backward_ex(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                    const float* seed, int64_t N, int V, float* ws, int flags, void* stream) {
  if (!ops || !plan || !seed || !ws || N <= 0 || V <= 0) return BNNP_ERR_INVALID;
  if ((flags & BNNP_BWD_FIRST_CONV_POSSUM) && !bnnp_net_first_conv_possum_ok(ops, n_ops)) return BNNP_ERR_INVALID;
  if ((flags & BNNP_BWD_FIRST_CONV_PLANES) &&
      ((flags & BNNP_BWD_FIRST_CONV_POSSUM) || !net_first_conv_planes_structure_ok(ops, n_ops)))
    return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  const int K = plan->K;
  const int64_t R = N * V;
  int first_param = -1;
  for (int i = 0; i < n_ops; ++i) if (is_param_op(ops[i])) { first_param = i; break; }
  if (first_param < 0) return BNNP_ERR_INVALID;
  if (plan->Mtot > 0)
    BNNP_CUDA_CHECK(cudaMemsetAsync(ws + plan->gcat_off, 0, sizeof(float) * R * plan->Mtot, st));
  {
    const bnnp_op_plan_t& q = plan->op[n_ops - 1];
    int rc = copy2d(ws + q.grad_off, q.grad_ld, seed, K, R, K, st);
    if (rc) return rc;
  }
  bool fused_prev = false;
  for (int i = n_ops - 1; i > first_param; --i) {
    if (fused_prev) { fused_prev = false; continue; }      // this activation's backward was folded into the op after it
    const bnnp_op_t& o = ops[i];
    const bnnp_op_plan_t& q = plan->op[i];
    const bnnp_op_plan_t& qp = plan->op[i - 1];
    float* gout = ws + q.grad_off;
    float* gin = ws + qp.grad_off;
    int64_t isz = op_in_size(o);
    unsigned blocks = (unsigned)ceil_div64(R * isz, 256);
    switch (o.kind) {
      case BNNP_OP_LINEAR: {
        int rc;
        if (!g_bnnp_opt.net_simt && (long long)R * o.in_c * o.out_c >= kTcMinMacs && plan->conv_scratch_off >= 0 &&
            o.out_c <= 4096 && g_bnnp_opt.gemm_generated == 0)
          rc = linear_tma(gout, q.grad_ld, R, o.out_c, o.weight, o.in_c, o.in_c, 1, nullptr, gin, qp.grad_ld,
                          ws + plan->conv_scratch_off, st);
        else if (!g_bnnp_opt.net_simt && (long long)R * o.in_c * o.out_c >= kTcMinMacs)
          rc = tcg::gemm_tc(R, o.in_c, o.out_c, 1, tcg::TcRowMajor{gout, q.grad_ld, 0},
                            tcg::TcColMajor{o.weight, o.in_c, 0},
                            tcg::TcStoreAxpby(gin, qp.grad_ld, 0, 1.f, 0.f), 1, st);
        else
          rc = gemm_simt(R, o.in_c, o.out_c, 1, LoadRowMajorA{gout, q.grad_ld, 0},
                         LoadRowMajorB{o.weight, o.in_c, 0},
                         StoreAxpby{gin, qp.grad_ld, 0, 1.f, 0.f}, st);
        if (rc) return rc;
        break;
      }
      case BNNP_OP_CONV2D:
        if (plan->conv_scratch_off >= 0 && !g_bnnp_opt.net_simt && (long long)R * o.out_c * o.out_h * o.out_w * o.in_c * o.kh * o.kw >= kTcMinMacs) {
          const int E = o.in_c * o.kh * o.kw, L = o.out_h * o.out_w;
          float* dUt = ws + plan->conv_scratch_off;
          int rc;
          if (g_bnnp_opt.gemm_generated == 0 && g_bnnp_opt.conv_bwd_explicit == 0 &&
              conv_bwd_implicit_ok(o.in_c, o.out_c, o.in_h, o.in_w, o.out_h, o.out_w, o.kh, o.kw, o.stride_h, o.stride_w, R)) {
            // implicit GEMM: the gradient planes are fetched per filter tap by 4-D TMA boxes (out-of-range = padding); neither
            // the [R, E, L] product nor the col2im gather exists
            rc = conv_bwd_implicit(gout, q.grad_ld, R, o.in_c, o.out_c, o.in_h, o.in_w, o.out_h, o.out_w, o.kh, o.kw, o.pad_l,
                                   o.pad_t, o.weight, gin, qp.grad_ld, dUt, st);
            if (rc) return rc;
            break;
          }
          if (g_bnnp_opt.gemm_generated == 0 && R <= 65535)
            rc = conv_bwd_tma(gout, q.grad_ld, R, o.out_c, L, o.weight, E, dUt, dUt + ((int64_t)R * E * L + 63) / 64 * 64, st);
          else
            rc = tcg::gemm_tc(R * L, E, o.out_c, 1, ConvGoutRowsA{gout, q.grad_ld, L}, tcg::TcColMajor{o.weight, E, 0},
                              ConvStoreDUt(dUt, E, L), 1, st);
          if (rc) return rc;
          if (o.kh > 5 || o.kw > 5) return BNNP_ERR_UNSUPPORTED;
          {
            // few row slices: the tap offsets depend on (channel, position) only and are reused across rows
            const int64_t bx = ceil_div64((int64_t)o.in_c * o.in_h * o.in_w, 256);
            int64_t gy = ceil_div64(8192, bx);
            gy = gy < 1 ? 1 : (gy > R ? R : (gy > 65535 ? 65535 : gy));
            const dim3 grid((unsigned)bx, (unsigned)gy);
            if (o.kh <= 3 && o.kw <= 3)
              col2im_gather_kernel<3, 3><<<grid, 256, 0, st>>>(dUt, gin, qp.grad_ld, R, geom_of(o));
            else
              col2im_gather_kernel<5, 5><<<grid, 256, 0, st>>>(dUt, gin, qp.grad_ld, R, geom_of(o));
          }
          BNNP_LAUNCH_CHECK();
        } else if (R <= 65535) {
          int rc = gemm_simt(o.in_c, (int64_t)o.in_h * o.in_w, (int64_t)o.out_c * o.kh * o.kw, (int)R,
                             ConvWtA{o.weight, o.in_c, o.kh * o.kw}, ConvGoutB{gout, q.grad_ld, geom_of(o)},
                             ConvStoreGin{gin, qp.grad_ld, o.in_h * o.in_w}, st);
          if (rc) return rc;
        } else {
          conv_bwd_data_kernel<<<blocks, 256, 0, st>>>(gout, q.grad_ld, gin, qp.grad_ld, o.weight, R, geom_of(o));
          BNNP_LAUNCH_CHECK();
        }
        break;
      case BNNP_OP_MAXPOOL2D:
        if (o.kh > 3 || o.kw > 3) return BNNP_ERR_UNSUPPORTED;
        {
          // few z-blocks: the candidate windows depend on (channel, position) only and are reused across examples
          const int64_t bxy = ceil_div64((int64_t)o.in_h * o.in_w * o.in_c, 256);
          int64_t gz = ceil_div64(4096, bxy);
          gz = gz < 1 ? 1 : (gz > N ? N : gz);
          const dim3 grid((unsigned)bxy, 1, (unsigned)gz);
          const int32_t* idxp = reinterpret_cast<const int32_t*>(ws) + q.idx_off;
          // position sums only for the first conv (KFAC): see maxpool_bwd_possum_kernel
          const bool act_before = i - 1 > first_param && is_elementwise_op(ops[i - 1]);
          if ((flags & BNNP_BWD_FIRST_CONV_POSSUM) && (act_before ? i - 2 : i - 1) == first_param) {
            int rc = net_first_conv_possum(ops, n_ops, plan, N, V, ws, ws + plan->op[first_param].grad_off, st);
            if (rc) return rc;
            return BNNP_OK;                  // nothing upstream of the first parameterised layer
          }
          // an elementwise activation right in front of the pool (conv -> relu -> pool): its backward is folded into
          // this kernel's store and the op is skipped below
          const bool fuse_act = i - 1 > first_param && is_elementwise_op(ops[i - 1]);
          const float* actp = fuse_act ? ws + qp.act_off : nullptr;
          // stride >= 2: four positions per thread, compacted candidate windows, 16-byte stores (aligned rows)
          const bool vec4 = o.in_w % 4 == 0 && o.stride_h >= 2 && o.stride_w >= 2 && qp.grad_ld % 4 == 0 &&
                            (reinterpret_cast<uintptr_t>(gin) & 15) == 0 &&
                            (!fuse_act || (qp.act_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(actp) & 15) == 0));
          if ((flags & BNNP_BWD_FIRST_CONV_PLANES) && (fuse_act ? i - 2 : i - 1) == first_param && !vec4) return BNNP_ERR_INVALID;
          if (vec4) {
            const int64_t b4 = ceil_div64((int64_t)o.in_h * (o.in_w / 4) * o.in_c, 128);
            int64_t gz4 = ceil_div64(8192, b4);
            gz4 = gz4 < 1 ? 1 : (gz4 > N ? N : gz4);
            const dim3 grid4((unsigned)b4, 1, (unsigned)gz4);
            const bool planes = (flags & BNNP_BWD_FIRST_CONV_PLANES) && (fuse_act ? i - 2 : i - 1) == first_param;
            if (planes && fuse_act)
              maxpool_bwd4_kernel<1, 1><<<grid4, 128, 0, st>>>(gout, q.grad_ld, gin, qp.grad_ld, idxp, N, V, geom_of(o), actp,
                                                             qp.act_ld, ops[i - 1].kind, R * qp.grad_ld);
            else if (planes)
              maxpool_bwd4_kernel<0, 1><<<grid4, 128, 0, st>>>(gout, q.grad_ld, gin, qp.grad_ld, idxp, N, V, geom_of(o), nullptr, 0, 0,
                                                             R * qp.grad_ld);
            else if (fuse_act)
              maxpool_bwd4_kernel<1><<<grid4, 128, 0, st>>>(gout, q.grad_ld, gin, qp.grad_ld, idxp, N, V, geom_of(o), actp,
                                                          qp.act_ld, ops[i - 1].kind);
            else
              maxpool_bwd4_kernel<0><<<grid4, 128, 0, st>>>(gout, q.grad_ld, gin, qp.grad_ld, idxp, N, V, geom_of(o), nullptr, 0, 0);
          } else if (o.in_w % 2 == 0 && o.stride_h >= 2 && o.stride_w >= 2 && qp.grad_ld % 2 == 0 &&
                     (reinterpret_cast<uintptr_t>(gin) & 7) == 0 &&
                     (!fuse_act || (qp.act_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(actp) & 7) == 0))) {
            // even width: two positions per thread, 8-byte stores
            const int64_t b2 = ceil_div64((int64_t)o.in_h * (o.in_w / 2) * o.in_c, 128);
            int64_t gz2 = ceil_div64(8192, b2);
            gz2 = gz2 < 1 ? 1 : (gz2 > N ? N : gz2);
            const dim3 grid2((unsigned)b2, 1, (unsigned)gz2);
            if (fuse_act)
              maxpool_bwd2_kernel<1><<<grid2, 128, 0, st>>>(gout, q.grad_ld, gin, qp.grad_ld, idxp, N, V, geom_of(o), actp,
                                                          qp.act_ld, ops[i - 1].kind);
            else
              maxpool_bwd2_kernel<0><<<grid2, 128, 0, st>>>(gout, q.grad_ld, gin, qp.grad_ld, idxp, N, V, geom_of(o), nullptr, 0, 0);
          } else if (fuse_act) {
            maxpool_bwd_kernel<1><<<grid, 256, 0, st>>>(gout, q.grad_ld, gin, qp.grad_ld, idxp, N, V, geom_of(o),
                                                        actp, qp.act_ld, ops[i - 1].kind);
          } else {
            maxpool_bwd_kernel<0><<<grid, 256, 0, st>>>(gout, q.grad_ld, gin, qp.grad_ld, idxp, N, V, geom_of(o),
                                                        nullptr, 0, 0);
          }
          fused_prev = fuse_act;
        }
        BNNP_LAUNCH_CHECK();
        break;
      case BNNP_OP_AVGPOOL2D:
        avgpool_bwd_kernel<<<blocks, 256, 0, st>>>(gout, q.grad_ld, gin, qp.grad_ld, R, geom_of(o));
        BNNP_LAUNCH_CHECK();
        break;
      case BNNP_OP_RELU: case BNNP_OP_TANH: case BNNP_OP_SIGMOID:
        {
          int64_t gy = ceil_div64(8192, ceil_div64(isz, 256));
          gy = gy < 1 ? 1 : (gy > R / V ? R / V : gy);
          if (gy < 1) gy = 1;
          act_bwd_kernel<<<dim3((unsigned)ceil_div64(isz, 256), (unsigned)gy), 256, 0, st>>>(
              gout, q.grad_ld, ws + q.act_off, q.act_ld, R, V, isz, o.kind);
        }
        BNNP_LAUNCH_CHECK();
        break;
      case BNNP_OP_FLATTEN:
        break;
      default:
        return BNNP_ERR_UNSUPPORTED;
    }
  }
  return BNNP_OK;
}
