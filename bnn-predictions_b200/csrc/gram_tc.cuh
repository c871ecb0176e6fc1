// tcgen05 / TMEM weighted-Gram engine for the full GGN (see gram_tc.cu).
#pragma once
#include "common.cuh"

namespace bnnp {
int64_t gram_tc_scratch_floats(const bnnp_plan_t* plan, int64_t N);
bool gram_tc_preferred(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N);
int gram_tc_full_accum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                       const float* ws, const float* Lambda, float* H, float* scratch, int64_t row_begin,
                       int64_t row_end, int prepared, cudaStream_t st);
int gram_row_chunks(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int n_chunks, int64_t* bounds);
}  // namespace bnnp
