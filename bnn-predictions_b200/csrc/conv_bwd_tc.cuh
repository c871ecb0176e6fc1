// Backward-data of a stride-1 convolution as an implicit GEMM on TMA-fed tensor cores (conv_bwd_tc.cu).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace bnnp {
// dX[r * dld + ci * H * W + y * W + x] = sum_{a,b,co} gout[r * gld + co * OH * OW + (y + pt - a) * OW + (x + pl - b)] * Wt[co, ci, a, b]
// for R back-propagated rows; stride 1, Cin <= 256, W <= 128.  scratch: conv_bwd_implicit_scratch_floats().
bool conv_bwd_implicit_ok(int Cin, int Cout, int H, int W, int OH, int OW, int kh, int kw, int sh, int sw, int64_t R);
int64_t conv_bwd_implicit_scratch_floats(int64_t R, int Cout, int L, int Cin, int taps);
int conv_bwd_implicit(const float* gout, int64_t gld, int64_t R, int Cin, int Cout, int H, int W, int OH, int OW, int kh, int kw,
                      int pl, int pt, const float* Wt, float* dX, int64_t dld, float* scratch, cudaStream_t st, int fwd = 0,
                      const float* bias = nullptr);
// fwd = 1: the FORWARD convolution out[n, co, y, x] = bias[co] + sum_{a,b,ci} in[n, ci, y - pt_c + a, x - pl_c + b] * W[co, ci, a, b]
// through the same kernel -- call with gout = in, (Cin, Cout) = (C_out, C_in) of the layer, (H, W) = output map, (OH, OW) = input
// map, pl = kw - 1 - pl_c, pt = kh - 1 - pt_c, Wt = the layer's weight [C_out, C_in, kh, kw] (mirrored and transposed while it
// is split), dX = out.
}  // namespace bnnp
