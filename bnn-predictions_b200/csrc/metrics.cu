// Metric reductions that follow every predictive in the drivers:
// nll_cls / acc / macc / ece of preds/utils.py:6-52, one pass over the [N,K] mean predictive.
// Integer work (arg-max, label compare, bin membership, counts) is exact; the floating sums are
// accumulated in fp64 so that their order does not matter at fp32 resolution.
#include <float.h>
#include "common.cuh"

namespace bnnp {

constexpr int kMaxBins = 64;

// out[0] = sum_n log p(y_n), out[1] = #correct, out[2] = N (as counted), then per bin b:
// out[3+3b] = count, out[4+3b] = sum of confidences, out[5+3b] = #correct in the bin.
__global__ void __launch_bounds__(256)
metric_cls_kernel(int lik, const float* __restrict__ probs, const void* __restrict__ y, int64_t N, int K,
                  const float* __restrict__ bounds, int bins, double* __restrict__ out) {
  __shared__ double acc[3 + 3 * kMaxBins];
  const int nacc = 3 + 3 * bins;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) {
    float logp, conf;
    int64_t label, pred;
    bool correct;
    if (lik == BNNP_LIK_CATEGORICAL) {
      const float* p = probs + n * K;
      label = reinterpret_cast<const int64_t*>(y)[n];
      float sum = 0.f, best = p[0];
      pred = 0;
      for (int k = 0; k < K; ++k) {
        sum += p[k];
        if (p[k] > best) { best = p[k]; pred = k; }       // first maximum wins (torch.max / argmax on ties)
      }
      // Categorical(probs=p): normalise, clamp to [eps, 1-eps], log  (torch.distributions.utils.probs_to_logits)
      const float pn = (label >= 0 && label < K) ? p[label] / sum : 0.f;
      logp = logf(fminf(fmaxf(pn, FLT_EPSILON), 1.f - FLT_EPSILON));
      conf = best;
      correct = pred == label;
    } else {                                               // Bernoulli: probs [N], labels float {0,1}
      const float p = probs[n];
      const float yy = reinterpret_cast<const float*>(y)[n];
      const float pc = fminf(fmaxf(p, FLT_EPSILON), 1.f - FLT_EPSILON);
      const float x = logf(pc) - log1pf(-pc);              // logits of Bernoulli(probs=p)
      logp = -((1.f - yy) * x + fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x))));
      // acc(): (g >= 0.5) == y;  ece(): probs = [1-p, p], max with the first index on ties
      correct = ((p >= 0.5f) ? 1.f : 0.f) == yy;
      const float q = 1.f - p;
      conf = (p > q) ? p : q;
      pred = (p > q) ? 1 : 0;
      label = (int64_t)yy;
    }
    atomicAdd(&acc[0], (double)logp);
    if (correct) atomicAdd(&acc[1], 1.0);
    atomicAdd(&acc[2], 1.0);
    const bool hit = pred == label;                        // ece's own accuracy (predictions.eq(labels.long()))
    for (int b = 0; b < bins; ++b)
      if (conf > bounds[b] && conf <= bounds[b + 1]) {
        atomicAdd(&acc[3 + 3 * b], 1.0);
        atomicAdd(&acc[4 + 3 * b], (double)conf);
        if (hit) atomicAdd(&acc[5 + 3 * b], 1.0);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nacc; i += blockDim.x)
    if (acc[i] != 0.0) atomicAdd(out + i, acc[i]);
}

__global__ void zero_double_kernel(double* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}

}  // namespace bnnp
using namespace bnnp;

extern "C" int bnnp_metric_cls(int lik, const float* probs, const void* y, int64_t N, int K, const float* bounds,
                               int bins, double* out, void* stream) {
  if (!probs || !y || !bounds || !out || N <= 0 || K <= 0 || bins < 0 || bins > kMaxBins) return BNNP_ERR_INVALID;
  if (lik != BNNP_LIK_CATEGORICAL && lik != BNNP_LIK_BERNOULLI) return BNNP_ERR_INVALID;
  if (lik == BNNP_LIK_BERNOULLI && K != 1) return BNNP_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  zero_double_kernel<<<1, 256, 0, st>>>(out, 3 + 3 * bins);
  BNNP_LAUNCH_CHECK();
  metric_cls_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, st>>>(lik, probs, y, N, K, bounds, bins, out);
  BNNP_LAUNCH_CHECK();
  return BNNP_OK;
}
