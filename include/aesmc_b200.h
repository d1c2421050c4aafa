/*
 * aesmc_b200.h -- C ABI of libaesmc_b200.so: the sm_100a implementation of aesmc's SMC hot path.
 *
 * The reference (tuananhle7/aesmc) has no FFI; its boundary is the Python API of
 * aesmc/inference.py, aesmc/state.py, aesmc/math.py, aesmc/statistics.py.  Each
 * entry point below is what a binding for that path calls, and cites the reference lines it
 * replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; all tensors are dense,
 *     row-major: particle tables are [B, K] (row = independent SMC problem, column = particle),
 *     particle state is [B, K, D]
 *   - sizes are int64_t; `stream` is a cudaStream_t passed as void*; functions only enqueue work on
 *     that stream, never synchronise, own no memory and keep no state between calls (re-entrant)
 *   - return value: AESMC_OK or an AESMC_ERR_* code; aesmc_last_error_string() describes the last
 *     failure on the calling thread
 *   - data-dependent conditions (NaN weights, degenerate rows) are reported asynchronously by
 *     OR-ing AESMC_FLAG_* bits into the caller-provided int32 `flags` word
 */
#ifndef AESMC_B200_H
#define AESMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AESMC_OK 0
#define AESMC_ERR_BAD_ARG 1     /* null/negative/misaligned argument              -> ValueError   */
#define AESMC_ERR_LAUNCH 2      /* cudaGetLastError() after launch                -> RuntimeError */
#define AESMC_ERR_UNSUPPORTED 3 /* shape outside what this build handles          -> RuntimeError */

/* bits OR-ed into *flags by kernels */
#define AESMC_FLAG_NAN 1        /* a log-weight is NaN: inference.py:244-245 FloatingPointError   */
#define AESMC_FLAG_DEGENERATE 2 /* a row's normaliser is not finite and positive (all -inf, +inf):
                                   the reference silently emits index K there         */
#define AESMC_FLAG_INDEX_RANGE 4 /* an ancestor index outside [0, K) was passed to a gather       */

/* resampling arithmetic */
#define AESMC_MODE_EXACT 0 /* reference-order arithmetic: numpy float32 exp, scipy logsumexp's
                              pairwise sums, sequential float32 cumulative sum, float64 comparison;
                              ancestor indices are bit-identical to the reference's              */
#define AESMC_MODE_FAST 1  /* warp-shuffle reductions, ex2.approx, parallel block scan; indices may
                              differ from the reference where a position is within rounding of a
                              CDF boundary                                                        */

int aesmc_version(void);
const char *aesmc_last_error_string(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t aesmc_launch_count(void);
/* largest K the single-CTA-per-row step kernel accepts (shared-memory bound) */
int64_t aesmc_max_particles_single_cta(void);

/*
 * One SMC time step over B independent rows of K particles.
 * Replaces: log-weight formation           inference.py:97-98, 125-126   (log_w = (a + b) - c)
 *           per-step log-evidence term     inference.py:130              (lse[b] = logsumexp_k log_w)
 *           sample_ancestral_index         inference.py:234-269, math.py:6-51
 *           state.resample of the newest latent   state.py:158-183 via inference.py:102-104
 *
 *   lp_a [B,K]; lp_b, lp_c [B,K] or NULL (treated as 0)
 *   u    [B] float64 uniforms in [0,1), one per row (inference.py:250); NULL iff idx == NULL
 *   log_w [B,K] out (must not alias the inputs); lse [B] out or NULL
 *   idx  [B,K] int32 out, or NULL to skip resampling (last time step / importance sampling)
 *   x_in [B,K,D] -> x_out [B,K,D] = x_in[b, idx[b,k], :] (fused ancestral gather); both NULL to skip
 *   flags: int32[1], OR-ed with AESMC_FLAG_*
 */
int aesmc_smc_step_f32(const float *lp_a, const float *lp_b, const float *lp_c, const double *u,
                       int64_t B, int64_t K, float *log_w, float *lse, int32_t *idx,
                       const float *x_in, float *x_out, int64_t D, int32_t *flags, int mode,
                       void *stream);

/*
 * Rows beyond the register-blocked single-CTA kernel (K > 16 384, or 8 192 < K with K % 4 != 0) run a
 * multi-CTA pipeline that needs caller-allocated device scratch: aesmc_smc_step_workspace_bytes(B, K) bytes
 * (0 when not needed), 256-byte aligned, passed to aesmc_smc_step_ws_f32 (identical to aesmc_smc_step_f32
 * otherwise).  Without a workspace, rows up to aesmc_max_particles_single_cta() still run (a slower
 * shared-memory kernel, 3-4x behind the multi-CTA path); longer rows return AESMC_ERR_BAD_ARG.
 */
int64_t aesmc_smc_step_workspace_bytes(int64_t B, int64_t K);
int aesmc_smc_step_ws_f32(const float *lp_a, const float *lp_b, const float *lp_c, const double *u,
                          int64_t B, int64_t K, float *log_w, float *lse, int32_t *idx,
                          const float *x_in, float *x_out, int64_t D, int32_t *flags, int mode,
                          void *workspace, int64_t workspace_bytes, void *stream);

/*
 * One SMC time step with the user model fused in, for scalar linear-Gaussian state-space models
 * (the sampling and the three log-densities that inference.py:108-126 obtains from torch
 * kernels are evaluated inside the step kernel, operation for operation as torch.distributions.Normal does
 * in float32):   x ~ N(q.mult*x_prev + q_off[b], q.scale^2)
 *                log_w = (log N(x; t.mult*x_prev + t.off, t.scale^2) + log N(y[b]; e.mult*x + e.off, e.scale^2))
 *                        - log N(x; q.mult*x_prev + q_off[b], q.scale^2)
 * then lse / systematic ancestors / gather of x exactly as aesmc_smc_step_f32.
 *   x_prev [B,K] resampled latents of the previous step (NULL at t = 0: treated as 0)
 *   y [B]; noise [B,K] injected standard normals or NULL (Philox4x32-10 keyed by seed, stream_offset;
 *   seed_dev non-NULL: the key is read from device memory at run time, for CUDA-graph replays)
 *   q_off [B] per-row proposal offset or NULL (then the proposal's scalar offset)
 *   params_host: HOST pointer to 15 floats = (mult, off, scale, 2*scale^2, log scale) for t, e, q
 *   x_new, log_w: optional outputs (proposed latents, log-weights); x_out NULL skips resampling (last step);
 *   idx may be NULL on its own when the ancestors are not wanted (filtering for the evidence only)
 */
int aesmc_smc_step_lg_f32(const float *x_prev, const float *y, const float *noise, const float *q_off,
                          const float *params_host, float half_log_2pi, uint64_t seed, const uint64_t *seed_dev,
                          uint64_t stream_offset, int64_t B, int64_t K, const double *u, float *x_new, float *log_w,
                          float *lse, int32_t *idx, float *x_out, int32_t *flags, int mode, void *stream);

/* The same step with the 15 model parameters read from DEVICE memory (training: the parameters change every
 * optimiser step and never visit the host, so the whole step can also be captured in a CUDA graph). */
int aesmc_smc_step_lg_dev_f32(const float *x_prev, const float *y, const float *noise, const float *q_off,
                              const float *params_dev, float half_log_2pi, uint64_t seed, const uint64_t *seed_dev,
                              uint64_t stream_offset, int64_t B, int64_t K, const double *u, float *x_new, float *log_w,
                              float *lse, int32_t *idx, float *x_out, int32_t *flags, int mode, void *stream);

/*
 * Vector latents (BASELINE config 3): proposal sampling and the three log-densities of a D-dimensional linear-
 * Gaussian state-space model with diagonal noise, for all B*K particles in one launch -- what the user model's
 * torch callables compute per time step in inference.py:108-126 (proposal -> state.sample -> three state.log_prob
 * -> (transition + emission) - proposal):
 *     x_t | x_{t-1} ~ N(A x_{t-1} + b, diag sx^2)    y_t | x_t ~ N(C x_t + d, diag sy^2)
 *     q(x_t | x_{t-1}, y_t) = N(Wx x_{t-1} + q_row, diag sq^2)       (bootstrap != 0: q is the transition itself)
 *   x_prev [B,K,D] resampled latents (NULL at t = 0: then b, sx are the initial loc / scale and Wx is ignored)
 *   y [B,Dy]; noise [B,K,D] injected standard normals or NULL (Philox4x32-10 keyed by seed / stream_offset)
 *   q_row [B,D]: the part of the proposal mean that depends on the row only (Wy y_t + bias), NULL iff bootstrap
 *   params_host (HOST memory, row-major): A [D*D] | b [D] | sx [D] | C [Dy*D] | d [Dy] | sy [Dy] | Wx [D*D] | sq [D]
 *   out: x_new [B,K,D] proposed latents, log_w [B,K] = (log p(x|x_prev) + log p(y|x)) - log q(x|x_prev,y)
 * 1 <= D, Dy <= 16.  The log-weights then go through aesmc_smc_step_f32 (lp_a = log_w, x_in = x_new).
 */
int aesmc_lgv_propose_f32(const float *x_prev, const float *y, const float *noise, const float *q_row,
                          const float *params_host, int64_t D, int64_t Dy, int bootstrap, uint64_t seed,
                          uint64_t stream_offset, int64_t B, int64_t K, float *x_new, float *log_w, void *stream);

/*
 * Backward of one fused linear-Gaussian step: replaces torch autograd of losses.get_loss(..., 'aesmc')
 * (losses.py:5-65 -> inference.py:99-134) for that model family -- the logsumexp gradient, the ancestral
 * gather's scatter-add over descendants (state.py:158-183 backward) and the analytic Normal.log_prob /
 * rsample gradients in one launch per time step.
 *   x, x_prev [B,K]: the step's proposed latents and its (resampled) inputs (x_prev NULL at t = 0, then g_x_prev NULL)
 *   lse, g_lse [B]: the step's log-normaliser and dL/dlse;  params_dev: the 15 floats of the forward step
 *   g_next [B,K], idx [B,K]: dL/d(resampled latents) of the NEXT step and this step's ancestors (both NULL at t = T-1)
 *   g_x_prev [B,K] out: dL/dx_prev;  g_params [B,6] out: per-row sums for
 *   (t.mult, t.off, e.mult, e.off, q.mult, q_off[b]); scales are treated as constants.  K <= 16384.
 */
int aesmc_lg_step_bwd_f32(const float *x, const float *x_prev, const float *y, const float *q_off,
                          const float *params_dev, const float *lse, const float *g_lse, const float *g_next,
                          const int32_t *idx, int64_t B, int64_t K, float *g_x_prev, float *g_params, void *stream);

/*
 * Resampling entered at a later stage (used by the staged parity tests, and useful on their own):
 *   from normalised weights w [B,K]: cumulative sum (inference.py:257), renormalisation by the last
 *   entry (:260-261), search (:263-264);  from a normalised CDF [B,K]: the search alone.
 */
int aesmc_resample_from_weights_f32(const float *w, const double *u, int64_t B, int64_t K, int32_t *idx,
                                    int32_t *flags, int mode, void *stream);
int aesmc_resample_from_cdf_f32(const float *cdf, const double *u, int64_t B, int64_t K, int32_t *idx,
                                int32_t *flags, void *stream);

/* Importance-sampling accumulation (inference.py:156-157): acc += (a + b) - c elementwise over n
 * floats; log_w (nullable) receives the per-step term. */
int aesmc_is_accumulate_f32(const float *lp_a, const float *lp_b, const float *lp_c, float *acc,
                            float *log_w, int64_t n, int first, void *stream);

/* Row-wise logsumexp over particles (inference.py:130,158; statistics.py:90): lse[b]. */
int aesmc_logsumexp_f32(const float *log_w, int64_t B, int64_t K, float *lse, int32_t *flags, void *stream);
int aesmc_logsumexp_f64(const double *log_w, int64_t B, int64_t K, double *lse, int32_t *flags, void *stream);

/* Autograd of the step's log-weight/log-evidence outputs:
 *   g[b,k] = g_log_w[b,k] (nullable) + g_lse[b] (nullable) * exp(log_w[b,k] - lse[b])
 * written to g_pos (gradient of lp_a and lp_b) and, if non-NULL, -g to g_neg (gradient of lp_c). */
int aesmc_step_bwd_f32(const float *log_w, const float *lse, const float *g_log_w, const float *g_lse,
                       int64_t B, int64_t K, float *g_pos, float *g_neg, void *stream);

/* lognormexp / exponentiate_and_normalize over the last axis (math.py:6-51, torch branch):
 * out[b,k] = log_w[b,k] - lse[b]   (exponentiate == 0)   or   exp(log_w[b,k] - lse[b]). */
int aesmc_lognormexp_f32(const float *log_w, int64_t B, int64_t K, float *out, int exponentiate, void *stream);

/* Ancestral gather, any element type (state.py:158-183): dst[b,k,:] = src[b, idx[b,k], :] where a
 * particle is `row_bytes` contiguous bytes.  idx is int32 (idx_is_i64 == 0) or int64. */
int aesmc_gather_bytes(const void *src, const void *idx, int idx_is_i64, int64_t B, int64_t K,
                       int64_t row_bytes, void *dst, int32_t *flags, void *stream);

/* Backward of the gather (autograd of torch.gather = scatter_add, state.py:179):
 * gsrc[b,j,:] = sum_{k : idx[b,k]==j} gdst[b,k,:].  gsrc is fully written (zeros where no child).
 * sorted != 0: the caller guarantees every idx row is non-decreasing (true for indices produced by
 * aesmc_smc_step_f32); children are then summed run by run in k order, deterministically and
 * without atomics.  sorted == 0: atomic scatter-add, any index pattern. */
int aesmc_gather_bwd_f32(const float *gdst, const void *idx, int idx_is_i64, int64_t B, int64_t K,
                         int64_t D, float *gsrc, int sorted, void *stream);
int aesmc_gather_bwd_f64(const double *gdst, const void *idx, int idx_is_i64, int64_t B, int64_t K,
                         int64_t D, double *gsrc, int sorted, void *stream);

/* Genealogy composition (inference.py:226-229): out[b,k] = prev[b, cur[b,k]]; int32. */
int aesmc_compose_index_i32(const int32_t *prev, const int32_t *cur, int64_t B, int64_t K,
                            int32_t *out, void *stream);
/* out[b,k] = k (inference.py:215-220) */
int aesmc_iota_index_i32(int64_t B, int64_t K, int32_t *out, void *stream);
/* int32 <-> int64 index conversion for the LongTensor API surface (inference.py:266-269) */
int aesmc_index_widen(const int32_t *in, int64_t *out, int64_t n, void *stream);
int aesmc_index_narrow(const int64_t *in, int32_t *out, int64_t n, void *stream);

/* torch.distributions.Normal.log_prob over a [B, K] particle table in one pass, bit-identical to torch's six
 * elementwise kernels (the log-density calls of state.py:114-155 / inference.py:88-120).  value / loc kinds:
 * 0 = [B*K] elements, 1 = [B] (one per row), 2 = one device element; loc == NULL: loc_host.  scale is a
 * scalar: scale_dev (device element: IEEE division by 2*scale^2, logf) or, if NULL, the host-computed
 * float32 reciprocal inv_two_var_host and log_scale_host (torch multiplies by the reciprocal of a CPU scalar). */
int aesmc_normal_log_prob_f32(const float *value, int value_kind, const float *loc, int loc_kind, float loc_host,
                              const float *scale_dev, float inv_two_var_host, float log_scale_host, float half_log_2pi,
                              int64_t B, int64_t K, float *out, void *stream);
/* Its backward: per-particle gradient terms g_value = -g d/var, g_loc = g d/var, g_scale = g (d^2/scale^3 - 1/scale)
 * (each output nullable; reductions over broadcast operands are the caller's). */
int aesmc_normal_log_prob_bwd_f32(const float *value, int value_kind, const float *loc, int loc_kind, float loc_host,
                                  const float *scale_dev, float scale_host, const float *g, int64_t B, int64_t K,
                                  float *g_value, float *g_loc, float *g_scale, void *stream);

/* Device self-test: compares the hot path's specialised float32 exp (non-positive arguments, custom
 * correctly-rounded division) with the general reference-order exp for EVERY float in [-104, -0] and
 * -inf.  out2: device uint64[2] = {number of mismatching inputs, bit pattern of one of them}. */
int aesmc_selftest_expf(uint64_t *out2, void *stream);

/* Test hook: force the exact row kernel's rare paths on every row of the following launches (0 = off, the default).
 * bit 0: the exact scan's verification fails (the row is redone with the plain sequential chain); bit 1: a boundary is
 * reported as needing the reference's float64 comparison (the row's run marks are redone by the general loop).  Results
 * must not change (tests/test_step_parity_gpu.py).  Process-wide, not thread-safe; returns the previous value. */
int aesmc_debug_force_rare_paths(int bits);

/* statistics.log_ess (statistics.py:79-91): 2*lse(lw) - lse(2*lw) per row. */
int aesmc_log_ess_f32(const float *log_w, int64_t B, int64_t K, float *out, void *stream);
int aesmc_log_ess_f64(const double *log_w, int64_t B, int64_t K, double *out, void *stream);

/* statistics.empirical_mean / empirical_variance fast path (statistics.py:7-76 with f = x, x^2):
 * mean[b,d] = sum_k w[b,k] x[b,k,d], second[b,d] = sum_k w[b,k] x[b,k,d]^2, w = softmax_k(log_w). */
int aesmc_weighted_moments_f32(const float *x, const float *log_w, int64_t B, int64_t K, int64_t D,
                               float *mean, float *second, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AESMC_B200_H */
