#include "gram_tc.cuh"
namespace bnnp {
int64_t gram_tc_scratch_floats(const bnnp_plan_t*, int64_t) { return 0; }
bool gram_tc_preferred(const bnnp_op_t*, int, const bnnp_plan_t*, int64_t) { return false; }
int gram_tc_full_accum(const bnnp_op_t*, int, const bnnp_plan_t*, int64_t, const float*, const float*,
                       float*, float*, cudaStream_t) { return BNNP_ERR_UNSUPPORTED; }
}  // namespace bnnp
