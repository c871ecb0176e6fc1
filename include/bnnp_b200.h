/*
 * bnnp_b200.h -- C ABI of libbnnp_b200.so: the B200 (sm_100a) implementation of the Laplace-GGN
 * hot path of aleximmer/BNN-predictions.
 *
 * The reference has no FFI of its own (it is pure Python on torch + BackPACK), so the drop-in
 * boundary is the `preds` Python API (bnn-predictions_b200/preds_b200 mirrors it) and this C ABI
 * sits one level below: each entry point replaces one library call site of the reference, cited
 * as  <file>:<line>  relative to the reference root.  See INTEGRATION.md for the ctypes stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless marked [host];
 *   - the caller allocates and owns every buffer (torch does, in our glue); nothing here
 *     allocates or frees device memory; the only process-wide state is the launch counter, the
 *     profiling hooks and the engine-selection options at the end of this header;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); every call is
 *     asynchronous on that stream;
 *   - float32 everywhere; matrices are row-major;
 *   - return value: 0 = BNNP_OK, negative = error (bnnp_strerror()).  The Python glue turns
 *     BNNP_ERR_INVALID into ValueError and everything else into RuntimeError.
 */
#ifndef BNNP_B200_H
#define BNNP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNNP_OK                0
#define BNNP_ERR_INVALID      -1   /* bad argument / shape */
#define BNNP_ERR_UNSUPPORTED  -2   /* layer type or size outside what the kernels cover */
#define BNNP_ERR_CUDA         -3   /* a CUDA runtime call or launch failed */
#define BNNP_ERR_NO_DEVICE    -4

#define BNNP_MAX_OPS  64
#define BNNP_MAX_K    32           /* max #outputs handled by the register-resident K x K code */
#define BNNP_MAX_K_WIDE 128        /* likelihood terms and the GLM sampler beyond that: warp / block per example (wide_k.cu) */

const char* bnnp_strerror(int code);
int bnnp_version(void);
/* 1 if a CUDA device of compute capability 10.x is visible, else 0 (never throws). */
int bnnp_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * Network descriptor: the layer zoo the Jacobians are taken over (preds/models.py:91-122 MLPS,
 * :239-285 CIFAR10Net, :182-236 CIFAR100Net).  One entry per leaf module in execution order.
 * A ZeroPad2d in front of a conv / pool (deepobs tf-"same" padding) is folded into pad_*.
 * ---------------------------------------------------------------------------------------- */
enum {
  BNNP_OP_LINEAR = 1, BNNP_OP_CONV2D = 2, BNNP_OP_RELU = 3, BNNP_OP_TANH = 4,
  BNNP_OP_MAXPOOL2D = 5, BNNP_OP_AVGPOOL2D = 6, BNNP_OP_FLATTEN = 7, BNNP_OP_SIGMOID = 8
};

typedef struct {
  int32_t kind;
  int32_t in_c, in_h, in_w;          /* per-example input shape; Linear: in_c = features, h = w = 1 */
  int32_t out_c, out_h, out_w;
  int32_t kh, kw, stride_h, stride_w;
  int32_t pad_l, pad_r, pad_t, pad_b;
  int32_t has_bias;
  const float* weight;               /* [out_c, in_c*kh*kw]  (Linear: [out, in]) */
  const float* bias;                 /* [out_c] or NULL */
  int64_t theta_off_w, theta_off_b;  /* offsets of weight / bias in the flat parameter vector */
} bnnp_op_t;

/* Where each op's saved tensors live inside the caller-provided float workspace. */
typedef struct {
  int64_t act_off, act_ld;           /* output activation of example n at ws + act_off + n*act_ld */
  int64_t grad_off, grad_ld;         /* d f_v / d output for row (n*V+v) at ws + grad_off + row*grad_ld */
  int64_t idx_off;                   /* maxpool argmax (int32 view of ws) or -1 */
  int32_t lin_col, lin_unit;         /* Linear: first column in A_cat / first unit column in G_cat */
} bnnp_op_plan_t;

typedef struct {
  int64_t total_floats;              /* workspace size to allocate, in floats */
  int64_t acat_off, gcat_off;        /* A_cat [N, Dtot] (layer inputs + a ones column per Linear),
                                        G_cat [N*V, Mtot] (d f_v / d pre-activation of every Linear);
                                        each layer's slice is padded to a multiple of 4 columns (zeros) */
  int32_t Dtot, Mtot;
  int32_t n_linear, first_linear;    /* #Linear ops and index of the first one */
  int64_t P;                         /* total #parameters */
  int32_t K;                         /* network outputs */
  int32_t reserved;
  int64_t conv_scratch_off;          /* scratch of the tensor-core conv GEMMs: unfolded input [N*L, E] /
                                        backward-data [N*V, E, L] (or -1 without conv layers) */
  bnnp_op_plan_t op[BNNP_MAX_OPS];
} bnnp_plan_t;

/* [host] Lay out the workspace for batches of N examples and V back-propagated columns. */
int bnnp_net_plan(const bnnp_op_t* ops, int n_ops, int64_t N, int V, bnnp_plan_t* plan);

/* Forward pass, saving every layer input (replaces model(data) inside preds/gradients.py:26 and
 * preds/laplace.py:56,73).  X: [N, in_size]; logits: [N, K] (also written to the workspace). */
int bnnp_net_forward(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                     const float* X, int64_t N, float* ws, float* logits, void* stream);

/* Back-propagate V columns at once.  seed: [N, V, K], row (n,v) is the vector multiplied from the
 * left onto d f/d(.).  seed = I gives output Jacobians (the K backward passes of
 * preds/gradients.py:24-36); seed = S_n^T gives BackPACK's sqrt-GGN backprop (Appendix A). */
int bnnp_net_backward(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                      const float* seed, int64_t N, int V, float* ws, void* stream);
/* The same with flags.  BNNP_BWD_FIRST_CONV_POSSUM (KFAC only): when the first parameterised layer is a conv followed
 * by (activation,) max-pool, its [N*V, Cout, L] output gradient is never materialised; the head of its gradient
 * buffer receives the position sums [N*V, Cout] that KFLR needs (preds/laplace.py:73-78), to be consumed by
 * bnnp_kfac_accum_ex with the same flag.  bnnp_net_first_conv_possum_ok() says whether the network qualifies. */
#define BNNP_BWD_FIRST_CONV_POSSUM 1
int bnnp_net_first_conv_possum_ok(const bnnp_op_t* ops, int n_ops);
/* BNNP_BWD_FIRST_CONV_PLANES (diagonal GGN only; not together with ..._POSSUM): when the first parameterised layer is a conv
 * followed by ([activation,] a stride >= 2 max-pool), the pool backward leaves that conv's output gradient as the bf16 hi / lo
 * planes [(r, c), pos] the conv-weight diagonal kernel streams (inside the fp32 gradient buffer they replace) instead of fp32;
 * consumed by bnnp_ggn_diag_accum_ex with the same flag.  bnnp_net_first_conv_planes_ok() says whether network and batch
 * qualify (preds/laplace.py:52-59 is the loop this serves). */
#define BNNP_BWD_FIRST_CONV_PLANES 2
int bnnp_net_first_conv_planes_ok(const bnnp_op_t* ops, int n_ops, int64_t N, int V);
int bnnp_net_backward_ex(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                         const float* seed, int64_t N, int V, float* ws, int flags, void* stream);

/* ------------------------------------------------------------------------------------------
 * Likelihood terms (preds/likelihoods.py).  lik: 0 categorical, 1 gaussian, 2 bernoulli.
 * ---------------------------------------------------------------------------------------- */
enum { BNNP_LIK_CATEGORICAL = 0, BNNP_LIK_GAUSSIAN = 1, BNNP_LIK_BERNOULLI = 2 };

/* Lambda [N,K,K]: categorical diag(p)-pp^T with p = clamp(softmax f, 1e-7, 1-1e-7)
 * (preds/likelihoods.py:90-93); gaussian 1/sigma2 (:50-52); bernoulli p(1-p) clamped (:67-69). */
int bnnp_lik_hessian(int lik, const float* f, int64_t N, int K, float sigma2, float* Lambda,
                     void* stream);
/* seed [N,V=K,K] for the BackPACK loss-Hessian square root (unclamped softmax; MSE: sqrt(2) I). */
int bnnp_lik_sqrt_seed(int lik, const float* f, int64_t N, int K, float* seed, void* stream);
/* seed [N,K,K] = identity (Jacobian back-propagation). */
int bnnp_identity_seed(int64_t N, int K, float* seed, void* stream);
/* residual [N,K]: onehot(y) - softmax f (:84-88), (y-f)/sigma2 (:47-48), y - sigmoid f (:25-26).
 * y is int64 labels for categorical, float [N,K] otherwise. */
int bnnp_lik_residual(int lik, const float* f, const void* y, int64_t N, int K, float sigma2,
                      float* rs, void* stream);
/* out = inv_link(f) in place-capable: softmax over the last dim / identity / sigmoid. */
int bnnp_lik_inv_link(int lik, const float* f, int64_t rows, int K, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Jacobians and GGN accumulation
 * ---------------------------------------------------------------------------------------- */
/* Materialise Js [N,P,K] (K last) from the saved factors -- API compatibility with
 * preds/gradients.py:20-46; small shapes only.  Requires backward with V = K, seed = I. */
int bnnp_jacobians(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                   const float* X, const float* ws, float* Js, void* stream);
/* The same after a backward pass with V columns and an arbitrary seed [N,V,K]: Js [N,P,V] = seed-weighted output
 * Jacobians.  V = 1 with a one-hot seed gives the [N,P] Jacobian of one output, the operand of the function-space kernels
 * J_c J_c^T (preds/gradients.py:49-77 Jacobians_gp, preds/gps.py:30-53). */
int bnnp_jacobians_cols(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, int V,
                        const float* X, const float* ws, float* Js, void* stream);

/* H[P,P] += sum_n J_n^T Lambda_n J_n  (preds/laplace.py:42, preds/optimizers.py:94), Linear-only
 * networks.  Only tiles on/below the diagonal are touched; call bnnp_symmetrize_lower() once
 * after the last batch.  engine: 0 auto, 1 SIMT fp32, 2 tcgen05 split-bf16 tensor cores.
 * scratch: caller buffer of bnnp_ggn_full_scratch_floats() floats. */
int64_t bnnp_ggn_full_scratch_floats(const bnnp_plan_t* plan, int64_t N, int engine);
int bnnp_ggn_full_accum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                        const float* ws, const float* Lambda, float* H, float* scratch,
                        int engine, void* stream);
/* The same restricted to rows [row_begin, row_end) of H (row_end < 0: all rows), so that finished row chunks can be
 * exchanged between GPUs while later chunks still accumulate (data-parallel mode).  Chunk boundaries come from
 * bnnp_ggn_full_row_chunks(): bounds[0..n] with bounds[0] = 0, bounds[n] = P, chunks of roughly equal triangle
 * area; returns n (<= n_chunks; 1 when the engine that would run is not row-chunked).  prepared != 0: the per-batch
 * operands in `scratch` were built by an earlier call for the same batch (same ws / Lambda / N) and are reused. */
int bnnp_ggn_full_row_chunks(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, int engine,
                             int n_chunks, int64_t* bounds /* [host], n_chunks + 1 */);
int bnnp_ggn_full_accum_rows(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                             const float* ws, const float* Lambda, float* H, float* scratch, int engine,
                             int64_t row_begin, int64_t row_end, int prepared, void* stream);
/* H[p, 0..p] = 0 for all rows: a fresh accumulator for bnnp_ggn_full_accum (which adds into the lower triangle only;
 * whatever the upper triangle holds is overwritten by bnnp_symmetrize_lower) at half the bytes of a full memset. */
int bnnp_zero_lower(float* H, int64_t P, void* stream);
/* H[p,q] = H[q,p] for p < q; _rows: only the source rows q in [row0,row1) (their mirror images H[p,q], p < q). */
int bnnp_symmetrize_lower(float* H, int64_t P, void* stream);
int bnnp_symmetrize_lower_rows(float* H, int64_t P, int64_t row0, int64_t row1, void* stream);

/* ------------------------------------------------------------------------------------------
 * The one exchange step of the data-parallel full-covariance path (the reference loop
 * preds/laplace.py:39-42 sharded over examples): in-place all-reduce(sum) of rows [row0,row1) of the LOWER TRIANGLE
 * of H over peer memory, with the prior precision added to the diagonal by the owner of each row.  H must live at
 * the same offset of a symmetric allocation on every rank: peers[w] = rank w's H mapped into this process [host
 * array of `world` device pointers]; multicast = NVLS multicast mapping of H or NULL.  With a multicast mapping the
 * sum is formed inside NVSwitch (multimem.ld_reduce) and broadcast with multimem.st; without, the owner of a row
 * loads it from every peer in rank order and stores the sum to every peer.  Every rank ends with identical bits.
 * The caller orders the launch against the other ranks (a cross-rank barrier on the stream before and after).
 * prior_vec [P] (device) or NULL -> prior_scalar.  n_ctas <= 0: default. */
int bnnp_exchange_lower_allreduce(float* const* peers /* [host] */, float* multicast, int rank, int world, int64_t P,
                                  int64_t row0, int64_t row1, const float* prior_vec, float prior_scalar,
                                  int n_ctas, void* stream);
/* Packed lower triangle of rows [row0,row1) (row p contributes p+1 floats) for a library all-reduce of half the bytes:
 * pack -> collective -> unpack (+ prior on the diagonal). */
int64_t bnnp_packed_lower_floats(int64_t row0, int64_t row1);
int bnnp_pack_lower(const float* H, int64_t P, int64_t row0, int64_t row1, float* packed, void* stream);
int bnnp_unpack_lower(const float* packed, int64_t P, int64_t row0, int64_t row1, float* H,
                      const float* prior_vec, float prior_scalar, void* stream);

/* diag[P] += factor * diag(sum_n J_n^T Lambda_n J_n) from a sqrt-seeded backward
 * (BackPACK DiagGGNExact as read by preds/laplace.py:52-59). */
int bnnp_ggn_diag_accum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                        int V, const float* X, const float* ws, float factor, float* diag,
                        float* scratch, void* stream);
/* The same with the flags the backward ran with (0 or BNNP_BWD_FIRST_CONV_PLANES). */
int bnnp_ggn_diag_accum_ex(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                           int V, const float* X, const float* ws, float factor, float* diag,
                           float* scratch, int flags, void* stream);
int64_t bnnp_ggn_diag_scratch_floats(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                     int64_t N, int V);

/* KFLR factors (preds/laplace.py:73-78 + preds/kron.py:58-69): for every parameterised op,
 * G += factor * sum_{n,v} s s^T  and  A += batch_factor * (1/N) sum_n a a^T.
 * factors: [host] array of 2*n_ops device pointers, {G_i, A_i} per op (NULL for other ops). */
int bnnp_kfac_accum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, int V,
                    const float* X, const float* ws, float factor, float batch_factor,
                    float* const* factors, float* scratch, void* stream);
int bnnp_kfac_accum_ex(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N, int V,
                       const float* X, const float* ws, float factor, float batch_factor,
                       float* const* factors, float* scratch, int flags, void* stream);
int64_t bnnp_kfac_scratch_floats(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                 int64_t N, int V);

/* ------------------------------------------------------------------------------------------
 * GLM predictive (preds/predictives.py:50-76)
 * ---------------------------------------------------------------------------------------- */
/* f_mu[N,K] = f + J (mu - theta) (:58); delta = mu - theta [P] (may be NULL -> f_mu = f). */
int bnnp_glm_fmu(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                 const float* X, const float* ws, const float* f, const float* delta, float* f_mu,
                 void* stream);
/* f_var[N,K,K] = J Sigma J^T, Sigma [P,P] symmetric (:63).  Linear-only networks.
 * engine: 0 auto, 1 SIMT (T = J Sigma as written), 2 tcgen05 structured
 * (T[n,u,u'] = a^T Sigma_uu' a' per unit pair: P^2 instead of K P^2 MACs per input). */
int64_t bnnp_glm_fvar_full_scratch_floats(const bnnp_plan_t* plan, int64_t N, int engine);
int bnnp_glm_fvar_full(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                       const float* ws, const float* Sigma, float* f_var, float* scratch,
                       int engine, void* stream);
/* The same on TMA-fed tensor cores: Sigma_hi / Sigma_lo are the bf16 hi/lo split of Sigma with a
 * 16-byte aligned row pitch of roundup8(P) elements, made once per posterior by bnnp_split_bf16()
 * (dst pitch = roundup8(cols)).  No generator warps: activations and Sigma stream through TMA. */
int bnnp_split_bf16(const float* src, int64_t ld_src, int64_t rows, int64_t cols, void* hi, void* lo,
                    void* stream);
int64_t bnnp_glm_fvar_full_tma_scratch_floats(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                              int64_t N);
int bnnp_glm_fvar_full_tma(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                           const float* ws, const float* Sigma, const void* Sigma_hi, const void* Sigma_lo,
                           float* f_var, float* scratch, void* stream);
/* f_var = J diag(sigma2) J^T (:65). */
int bnnp_glm_fvar_diag(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                       const float* X, const float* ws, const float* sigma2, float* f_var,
                       float* scratch, void* stream);
int64_t bnnp_glm_fvar_diag_scratch_floats(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                          int64_t N);
/* Kronecker functional variance (preds/kron.py:112-143) using the rank structure of the layer
 * Jacobians.  eig: [host] 4*n_ops device pointers {W_G, lam_G, W_A, lam_A} per op;
 * delta_w / delta_b: [host] prior precision per op for weight / bias group; dampen as :130-134. */
int bnnp_glm_fvar_kron(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                       const float* X, const float* ws, const float* const* eig,
                       const float* delta_w, const float* delta_b, int dampen, float* f_var,
                       float* scratch, void* stream);
int64_t bnnp_glm_fvar_kron_scratch_floats(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan,
                                          int64_t N);
/* out[S,N,K] = inv_link(f_mu + chol(f_var) eps) (:69-72, MultivariateNormal.sample).  eps [S,N,K]
 * injected draws, or NULL to draw in-kernel (Philox, `seed`).  status[N] (int32) = 0 ok,
 * 1 = f_var[n] not positive definite (caller takes the reference's fallback, :73-76). */
int bnnp_glm_sample(int lik, const float* f_mu, const float* f_var, const float* eps,
                    uint64_t seed, int64_t S, int64_t N, int K, float* out, int32_t* status,
                    void* stream);

/* One parameter group of Kron.functional_variance on a MATERIALISED Jacobian block
 * Jb[R, m*n] (row r = (b,k), row stride ldj) -- API compatibility with preds/kron.py:112-143
 * (Kron.functional_variance(Js)).  W2 == NULL selects the 1-factor (bias) form on Jb[R, m].
 * f_var[R/K, K, K] is accumulated into.  scratch: 2*R*m*n + m*n floats (1-factor: R*m + m). */
int bnnp_kron_fvar_block(const float* Jb, int64_t ldj, int K, int m, int n, const float* W1,
                         const float* l1, const float* W2, const float* l2, float delta, int dampen,
                         float* f_var, float* scratch, int64_t R, void* stream);
/* The reference's fallback when a K x K Cholesky fails (preds/predictives.py:73-76):
 * out = inv_link(f_mu + clamp(diag f_var, 1e-5) * eps). */
int bnnp_glm_sample_diag_fallback(int lik, const float* f_mu, const float* f_var, const float* eps,
                                  uint64_t seed, int64_t S, int64_t N, int K, float* out,
                                  void* stream);

/* ------------------------------------------------------------------------------------------
 * BNN predictive (preds/predictives.py:12-28) and Kron sampling (preds/kron.py:85-110)
 * ---------------------------------------------------------------------------------------- */
/* theta_s[S,P] = mu + (L eps)^T for L [P,P] lower (full) or sigma [P] (diag); eps [P,S]. */
int bnnp_sample_theta_full(const float* L, const float* mu, const float* eps, int64_t P,
                           int64_t S, float* theta_s, void* stream);
int bnnp_sample_theta_diag(const float* sigma, const float* mu, const float* eps, int64_t P,
                           int64_t S, float* theta_s, void* stream);
/* One parameter group of Kron.sample: out[s, off + (i*n+j)] = mu + W1((W1^T eps_s W2) * D)W2^T
 * (2-factor, eps [S,m,n]) or W(sqrt(1/(l+delta)) * W^T eps) (1-factor, eps [m,S]).
 * mu may be NULL (zero-mean draws as returned by Kron.sample).  scratch: 2*S*m*n + m*n floats. */
int bnnp_kron_sample_group(const float* W1, const float* l1, const float* W2, const float* l2,
                           int m, int n, float delta, int dampen, const float* eps, int64_t S,
                           const float* mu, float* theta_s, int64_t P, int64_t off,
                           float* scratch, void* stream);
/* out[S,N,K] = inv_link(f(X; theta_s)) for all S weight samples (Linear-only networks run as
 * batched-weight GEMMs; others loop over s).  ws: S-independent scratch sized by the call below. */
int64_t bnnp_bnn_forward_scratch_floats(const bnnp_op_t* ops, int n_ops, int64_t N, int64_t S);
int bnnp_bnn_forward(const bnnp_op_t* ops, int n_ops, int lik, const float* X, int64_t N,
                     const float* theta_s, int64_t S, int64_t P, float* out, float* scratch,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * Callers either side of the GGN kernel, Linear-only networks.  Both read the
 * layer factors left in `ws` by bnnp_net_forward + the identity-seeded bnnp_net_backward (V = K).
 * ---------------------------------------------------------------------------------------- */
/* grad[P] += sum_n J_n^T r_n -- replaces `G += einsum('mpk,mk->p', Js, rs)` in
 * LaplaceGGN.post_process (preds/optimizers.py:96).  rs [N,K]; scratch: bnnp_jtr_scratch_floats(). */
int64_t bnnp_jtr_scratch_floats(const bnnp_plan_t* plan, int64_t N);
int bnnp_jtr_accum(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                   const float* ws, const float* rs, float* grad, float* scratch, void* stream);
/* diag[P] += diag(sum_n J_n^T Lambda_n J_n) for an arbitrary per-example Lambda [N,K,K] -- the diagonal GGN that
 * vi_diag_refine re-evaluates every iteration on a fixed J (preds/refine.py:107-112).  scratch: bnnp_jtr_scratch_floats(). */
int bnnp_ggn_diag_lambda(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                         const float* ws, const float* Lambda, float* diag, float* scratch, void* stream);
/* out[s,n,k] = f[n,k] + sum_p J[(n,k),p] dtheta[s,p] -- replaces `offset + Js @ samples[i]` of
 * linear_sampling_predictive (preds/predictives.py:31-47) for all S samples at once; dtheta [S,P] =
 * theta_s - theta*; out [S,N,K] holds logits (apply bnnp_lik_inv_link for the link).  S <= 65535.
 * scratch: bnnp_glm_linear_samples_scratch_floats() floats. */
int64_t bnnp_glm_linear_samples_scratch_floats(const bnnp_plan_t* plan, int64_t N, int64_t S);
int bnnp_glm_linear_samples(const bnnp_op_t* ops, int n_ops, const bnnp_plan_t* plan, int64_t N,
                            const float* ws, const float* f, const float* dtheta, int64_t S,
                            float* out, float* scratch, void* stream);

/* ------------------------------------------------------------------------------------------
 * The steps right after the predictive in every driver
 * ---------------------------------------------------------------------------------------- */
/* mean over the S samples of bnnp_glm_sample without materialising them: mean [N,K] -- the `.mean(0)` of
 * experiments/imginference.py:391 / classification.py:36 fused into the sampling epilogue.  Same eps / seed /
 * status contract as bnnp_glm_sample (on a non-zero status redo through the materialising fallback). */
int bnnp_glm_sample_mean(int lik, const float* f_mu, const float* f_var, const float* eps,
                         uint64_t seed, int64_t S, int64_t N, int K, float* mean, int32_t* status,
                         void* stream);
/* linear_regression_predictive epilogue (preds/predictives.py:79-94): Lam = Hessian(f)*sigma2,
 * mu_star [N,K] = inv_link(f) + Lam (f_mu - f), var_f [N,K,K] = Lam f_var Lam, var_noise [N,K,K] = Lam*sigma2. */
int bnnp_lik_link_moments(int lik, const float* f, const float* f_mu, const float* f_var, float sigma2,
                          int64_t N, int K, float* mu_star, float* var_f, float* var_noise, void* stream);
/* nll_cls / acc / macc / ece (preds/utils.py:6-52) in one pass over probs ([N,K] categorical with int64 labels,
 * [N] Bernoulli with float labels).  bounds [bins+1] = torch.linspace(0,1,bins+1); out (double[3+3*bins]):
 * sum log p(y), #correct, N, then per bin (count, sum of confidences, #correct).  bins <= 64. */
int bnnp_metric_cls(int lik, const float* probs, const void* y, int64_t N, int K, const float* bounds,
                    int bins, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small dense helpers used by the posterior code (preds/laplace.py:111, preds/kron.py:58-69)
 * ---------------------------------------------------------------------------------------- */
/* C[M,N] = alpha * op(A) op(B) + beta * C, fp32 SIMT; ta/tb: 0 = as stored, 1 = transposed. */
int bnnp_sgemm(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
               int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
               void* stream);
/* Same contract on the tensor cores (tcgen05, split-bf16 = fp32-grade, operands synthesised on chip).
 * scratch (may be NULL): bnnp_sgemm_tc_scratch_floats() floats enabling deterministic split-K. */
int64_t bnnp_sgemm_tc_scratch_floats(int64_t M, int64_t N, int64_t K);
int bnnp_sgemm_tc(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                  int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
                  float* scratch, void* stream);
/* TMA-fed tensor-core GEMM on operands that were split ONCE into bf16 hi / lo copies (fp32-grade: three MMAs per product):
 * C[M,N] = alpha * sum_k A[m,k] B[n,k] + beta * C; A [M, lda], B [N, ldb] k-contiguous, lda / ldb multiples of 8 elements,
 * 16-byte aligned.  bnnp_split_bf16 makes such copies of a row-major matrix, bnnp_split_bf16_t of its transpose
 * (hi/lo[c, r] = src[r, c], row pitch `pitch` >= rows).  scratch: bnnp_gemm_nt_scratch_floats() floats (split-K
 * partials, one TMEM accumulation never spans more than 4096 k); may be NULL for K <= 4096.  Used by the Linear-layer
 * products and the KFAC Grams (preds/kron.py:58-69, preds/laplace.py:69-78). */
int64_t bnnp_gemm_nt_scratch_floats(int64_t M, int64_t N, int64_t K);
int bnnp_gemm_nt_split(int64_t M, int64_t N, int64_t K, float alpha, const void* Ah, const void* Al, int64_t lda,
                       const void* Bh, const void* Bl, int64_t ldb, float beta, float* C, int64_t ldc,
                       float* scratch, void* stream);
int bnnp_split_bf16_t(const float* src, int64_t ld_src, int64_t rows, int64_t cols, int64_t pitch, void* hi, void* lo,
                      void* stream);
/* bnnp_gemm_nt_split (K <= 4096) whose product is multiplied elementwise by had[m, n] (row pitch ldh) before it enters C:
 * C = alpha * (A B^T) o had + beta * C.  One layer's term of the function-space kernel of a Linear network,
 * K_c += (G_c G_c^T) o (A~ A~^T)  (preds/gps.py:56-88 contracts the materialised Jacobians instead). */
int bnnp_gemm_nt_split_had(int64_t M, int64_t N, int64_t K, float alpha, const void* Ah, const void* Al, int64_t lda,
                           const void* Bh, const void* Bl, int64_t ldb, float beta, float* C, int64_t ldc,
                           const float* had, int64_t ldh, void* stream);
/* bnnp_split_bf16 with an explicit destination pitch (>= cols, multiple of 8; columns cols..pitch-1 are zeroed): rows that
 * start on 128-byte lines keep the 64-byte rows of the TMA boxes inside one sector pair (function-space kernels,
 * preds/gps.py:56-88). */
int bnnp_split_bf16_rows(const float* src, int64_t ld_src, int64_t rows, int64_t cols, int64_t pitch, void* hi, void* lo,
                         void* stream);
/* Tensor-core GEMM of the one-off factorisation (preds/laplace.py:43-45: torch.cholesky, torch.cholesky_inverse,
 * torch.cholesky on [P,P]) with the accuracy of an fp32 library GEMM: operands split into THREE bf16 planes (hi, mid, lo;
 * `plane_stride` elements apart, rows k-contiguous with pitch lda / ldb, multiples of 8 elements, 16-byte aligned), six MMAs
 * per K-step, hi*hi accumulated in 128-k chunks that are summed in registers (round to nearest), the five correction terms in
 * their own accumulator.  C[M,N] = alpha * sum_k A[m,k] B[n,k] + beta * C, K <= 4096 per call (longer K: segments with
 * beta = 1).  lower != 0: only the tiles on / below the diagonal (M == N).  ktri = 1 / 2: the K loop of a 128 x 128 tile
 * starts at its column / row origin + koff (triangular operands); a tile without work leaves C untouched when beta == 1.
 * bnnp_split3_bf16 writes the planes of scale * src (transposed != 0: of its transpose, planes[c, r] = src[r, c]). */
int bnnp_split3_bf16(const float* src, int64_t ld_src, int64_t rows, int64_t cols, void* planes, int64_t pitch,
                     int64_t plane_stride, int transposed, float scale, void* stream);
int bnnp_gemm3_nt(int64_t M, int64_t N, int64_t K, float alpha, const void* A, int64_t lda, int64_t a_plane_stride,
                  const void* B, int64_t ldb, int64_t b_plane_stride, float beta, float* C, int64_t ldc, int lower,
                  int ktri, int64_t koff, void* stream);
/* The same with the 128 x 128 tiles of C dealt out between the `own_mod` processes of a data-parallel job: only the tiles
 * whose global column (own_dim = 0) or row (own_dim = 1) tile index -- local index + own_off -- is congruent to own_rank
 * modulo own_mod are computed, every other tile of C is left untouched.  Ownership is fixed in global coordinates, so across
 * the panels of a blocked factorisation a rank's copy stays complete on its own tiles; the caller gathers what the next
 * step needs (preds_b200/factor.py: the one-off factorisation shared between the ranks, every rank ends with the bits of
 * the single-GPU chain). */
int bnnp_gemm3_nt_owned(int64_t M, int64_t N, int64_t K, float alpha, const void* A, int64_t lda, int64_t a_plane_stride,
                        const void* B, int64_t ldb, int64_t b_plane_stride, float beta, float* C, int64_t ldc, int lower,
                        int ktri, int64_t koff, int own_mod, int own_rank, int own_dim, int64_t own_off, void* stream);
/* y += alpha * x  (Kron.update axpy, preds/kron.py:62-66). */
int bnnp_axpy(int64_t n, float alpha, const float* x, float* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): number of kernel launches issued by the library so far, and
 * CUDA-event timing of the tensor-core gram launches (recorded on the launching stream).
 * ---------------------------------------------------------------------------------------- */
uint64_t bnnp_launch_count(void);
/* Engine-selection options for tests and A/B measurements (process-wide, default 0; never read from the
 * environment; none of them makes a kernel skip work):  "gram_kernel" 0 auto / 1 single-CTA kernel,
 * "glm_single" 1 = single-CTA full-covariance GLM kernel, "glm_simt" 1 = kron / diag GLM contractions on the fp32
 * SIMT kernels, "debug_sync" 1 = synchronise after every GLM TMA launch,
 * "net_simt" 1 = forward / backward / diag / KFAC contractions on the fp32 SIMT kernels whatever their size,
 * "gemm_generated" 1 = Linear-layer products / KFAC Grams on the generated-operand GEMM instead of the TMA-fed one;
 * "conv_bwd_explicit" 1 = stride-1 conv backward-data as product + col2im gather and conv forward as unfold + GEMM instead of
 * the implicit-GEMM kernel (csrc/conv_bwd_tc.cu), and bnnp_net_first_conv_planes_ok() answers 0.  Unknown name: BNNP_ERR_INVALID. */
int bnnp_set_option(const char* name, int value);
int bnnp_get_option(const char* name);
int bnnp_prof_enable(int on);
int bnnp_prof_read(float* ms, double* executed_flops, double* useful_flops, int cap);
/* the same for the full-covariance GLM f_var launches (bnnp_glm_fvar_full_tma) since bnnp_prof_enable(1) */
int bnnp_glm_prof_read(float* ms, double* useful_flops, int cap);

#ifdef __cplusplus
}
#endif
#endif /* BNNP_B200_H */
