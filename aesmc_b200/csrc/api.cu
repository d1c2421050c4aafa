// api.cu -- the extern "C" surface declared in include/aesmc_b200.h: argument validation, error
// strings, launch accounting.  No torch types, no allocation, no synchronisation.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace aesmc {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return AESMC_ERR_LAUNCH;
    }
    return AESMC_OK;
}

// kernels (smc_step.cu, reduce.cu, gather.cu)
int64_t max_particles_single_cta();
int launch_smc_step(const float *, const float *, const float *, const double *, int64_t, int64_t, float *, float *,
                    int32_t *, const float *, float *, int64_t, int32_t *, int, int, void *, int64_t, cudaStream_t);
int64_t step_workspace_bytes(int64_t B, int64_t K);
bool smc_step_lg_supported(int64_t K);
int launch_smc_step_lg(const float *, const float *, const float *, const float *, const float *, const float *, float,
                       unsigned long long, const unsigned long long *, unsigned long long, int64_t, int64_t,
                       const double *, float *, float *, float *, int32_t *, float *, int32_t *, int, cudaStream_t);
int launch_lg_step_bwd(const float *, const float *, const float *, const float *, const float *, const float *, const float *,
                       const float *, const int32_t *, int64_t, int64_t, float *, float *, cudaStream_t);
int launch_lgv_propose(const float *, const float *, const float *, const float *, const float *, int64_t, int64_t, int,
                       unsigned long long, unsigned long long, int64_t, int64_t, float *, float *, cudaStream_t);
int launch_logsumexp_f32(const float *, int64_t, int64_t, float *, int32_t *, cudaStream_t);
int launch_logsumexp_f64(const double *, int64_t, int64_t, double *, int32_t *, cudaStream_t);
int launch_lognormexp_f32(const float *, int64_t, int64_t, float *, int, cudaStream_t);
int launch_log_ess_f32(const float *, int64_t, int64_t, float *, cudaStream_t);
int launch_log_ess_f64(const double *, int64_t, int64_t, double *, cudaStream_t);
int launch_step_bwd_f32(const float *, const float *, const float *, const float *, int64_t, int64_t, float *,
                        float *, cudaStream_t);
int launch_is_accumulate_f32(const float *, const float *, const float *, float *, float *, int64_t, int, cudaStream_t);
int launch_weighted_moments_f32(const float *, const float *, int64_t, int64_t, int64_t, float *, float *, cudaStream_t);
int launch_gather_bytes(const void *, const void *, int, int64_t, int64_t, int64_t, void *, int32_t *, cudaStream_t);
int launch_gather_bwd_f32(const float *, const void *, int, int64_t, int64_t, int64_t, float *, int, cudaStream_t);
int launch_gather_bwd_f64(const double *, const void *, int, int64_t, int64_t, int64_t, double *, int, cudaStream_t);
int launch_compose_index(const int32_t *, const int32_t *, int64_t, int64_t, int32_t *, cudaStream_t);
int launch_iota_index(int64_t, int64_t, int32_t *, cudaStream_t);
int launch_index_widen(const int32_t *, int64_t *, int64_t, cudaStream_t);
int launch_index_narrow(const int64_t *, int32_t *, int64_t, cudaStream_t);
int launch_selftest_expf(unsigned long long *, cudaStream_t);
void smc_step_x_set_debug_force(int bits);
int normal_log_prob_f32(const float *, int, const float *, int, float, const float *, float, float, float, int64_t, int64_t,
                        float *, cudaStream_t);
int normal_log_prob_bwd_f32(const float *, int, const float *, int, float, const float *, float, const float *, int64_t,
                            int64_t, float *, float *, float *, cudaStream_t);

} // namespace aesmc

using namespace aesmc;

#define REQUIRE(cond, fn)                                                            \
    do {                                                                             \
        if (!(cond)) {                                                               \
            set_error("%s: invalid argument: %s", fn, #cond);                        \
            return AESMC_ERR_BAD_ARG;                                                \
        }                                                                            \
    } while (0)

static inline cudaStream_t S(void *s) { return static_cast<cudaStream_t>(s); }
static const int64_t kMaxDim = 2147483647LL;

extern "C" {

int aesmc_version(void) { return 100; }
const char *aesmc_last_error_string(void) { return g_err; }
int64_t aesmc_launch_count(void) { return (int64_t)g_launches.load(); }
int64_t aesmc_max_particles_single_cta(void) { return max_particles_single_cta(); }

int64_t aesmc_smc_step_workspace_bytes(int64_t B, int64_t K)
{
    if (B <= 0 || K <= 0) return 0;
    return step_workspace_bytes(B, K);
}

int aesmc_smc_step_f32(const float *lp_a, const float *lp_b, const float *lp_c, const double *u, int64_t B,
                       int64_t K, float *log_w, float *lse, int32_t *idx, const float *x_in, float *x_out,
                       int64_t D, int32_t *flags, int mode, void *stream)
{
    return aesmc_smc_step_ws_f32(lp_a, lp_b, lp_c, u, B, K, log_w, lse, idx, x_in, x_out, D, flags, mode, nullptr, 0,
                                 stream);
}

int aesmc_smc_step_ws_f32(const float *lp_a, const float *lp_b, const float *lp_c, const double *u, int64_t B,
                          int64_t K, float *log_w, float *lse, int32_t *idx, const float *x_in, float *x_out,
                          int64_t D, int32_t *flags, int mode, void *workspace, int64_t workspace_bytes, void *stream)
{
    const char *fn = "aesmc_smc_step_f32";
    REQUIRE(lp_a && log_w && flags, fn);
    REQUIRE(B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    REQUIRE(mode == AESMC_MODE_EXACT || mode == AESMC_MODE_FAST, fn);
    REQUIRE((idx == nullptr) || (u != nullptr), fn);
    REQUIRE((x_in == nullptr) == (x_out == nullptr), fn);
    REQUIRE(x_in == nullptr || (idx != nullptr && D >= 1 && K * D <= kMaxDim), fn);
    REQUIRE(log_w != lp_a && log_w != lp_b && log_w != lp_c, fn);
    if (B == 0) return AESMC_OK;
    return launch_smc_step(lp_a, lp_b, lp_c, u, B, K, log_w, lse, idx, x_in, x_out, x_in ? D : 1, flags, mode, 0, workspace,
                           workspace_bytes, S(stream));
}

static int smc_step_lg_common(const char *fn, const float *x_prev, const float *y, const float *noise, const float *q_off,
                              const float *params_host, const float *params_dev, float half_log_2pi, uint64_t seed,
                              const uint64_t *seed_dev, uint64_t stream_offset, int64_t B, int64_t K, const double *u,
                              float *x_new, float *log_w, float *lse, int32_t *idx, float *x_out, int32_t *flags, int mode,
                              void *stream)
{
    REQUIRE(y && (params_host || params_dev) && flags, fn);
    REQUIRE(B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    REQUIRE(mode == AESMC_MODE_EXACT || mode == AESMC_MODE_FAST, fn);
    REQUIRE(idx == nullptr || x_out != nullptr, fn); // ancestors may be dropped (idx NULL), the resampled latents not
    REQUIRE(x_out == nullptr || u != nullptr, fn);
    if (!smc_step_lg_supported(K)) {
        set_error("%s: K=%lld not supported by the fused model kernel (64 <= K <= 16384, K %% 4 == 0)", fn, (long long)K);
        return AESMC_ERR_UNSUPPORTED;
    }
    const uintptr_t bits = reinterpret_cast<uintptr_t>(x_prev) | reinterpret_cast<uintptr_t>(noise) |
                           reinterpret_cast<uintptr_t>(x_new) | reinterpret_cast<uintptr_t>(log_w) |
                           reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(x_out);
    REQUIRE((bits & 15) == 0, fn);
    if (B == 0) return AESMC_OK;
    return launch_smc_step_lg(x_prev, y, noise, q_off, params_host, params_dev, half_log_2pi, seed,
                              reinterpret_cast<const unsigned long long *>(seed_dev), stream_offset, B, K, u, x_new, log_w,
                              lse, idx, x_out, flags, mode, S(stream));
}

int aesmc_smc_step_lg_f32(const float *x_prev, const float *y, const float *noise, const float *q_off,
                          const float *params_host, float half_log_2pi, uint64_t seed, const uint64_t *seed_dev,
                          uint64_t stream_offset, int64_t B, int64_t K, const double *u, float *x_new, float *log_w,
                          float *lse, int32_t *idx, float *x_out, int32_t *flags, int mode, void *stream)
{
    REQUIRE(params_host, "aesmc_smc_step_lg_f32");
    return smc_step_lg_common("aesmc_smc_step_lg_f32", x_prev, y, noise, q_off, params_host, nullptr, half_log_2pi, seed,
                              seed_dev, stream_offset, B, K, u, x_new, log_w, lse, idx, x_out, flags, mode, stream);
}

int aesmc_smc_step_lg_dev_f32(const float *x_prev, const float *y, const float *noise, const float *q_off,
                              const float *params_dev, float half_log_2pi, uint64_t seed, const uint64_t *seed_dev,
                              uint64_t stream_offset, int64_t B, int64_t K, const double *u, float *x_new, float *log_w,
                              float *lse, int32_t *idx, float *x_out, int32_t *flags, int mode, void *stream)
{
    REQUIRE(params_dev, "aesmc_smc_step_lg_dev_f32");
    return smc_step_lg_common("aesmc_smc_step_lg_dev_f32", x_prev, y, noise, q_off, nullptr, params_dev, half_log_2pi, seed,
                              seed_dev, stream_offset, B, K, u, x_new, log_w, lse, idx, x_out, flags, mode, stream);
}

int aesmc_lgv_propose_f32(const float *x_prev, const float *y, const float *noise, const float *q_row,
                          const float *params_host, int64_t D, int64_t Dy, int bootstrap, uint64_t seed,
                          uint64_t stream_offset, int64_t B, int64_t K, float *x_new, float *log_w, void *stream)
{
    const char *fn = "aesmc_lgv_propose_f32";
    REQUIRE(y && params_host && x_new && log_w, fn);
    REQUIRE(D >= 1 && D <= 16 && Dy >= 1 && Dy <= 16, fn);
    REQUIRE(bootstrap || q_row, fn);
    REQUIRE(B >= 0 && K >= 1 && B <= kMaxDim && K * D <= kMaxDim, fn);
    REQUIRE(x_new != x_prev, fn);
    if (B == 0) return AESMC_OK;
    return launch_lgv_propose(x_prev, y, noise, q_row, params_host, D, Dy, bootstrap ? 1 : 0, seed, stream_offset, B, K,
                              x_new, log_w, S(stream));
}

int aesmc_lg_step_bwd_f32(const float *x, const float *x_prev, const float *y, const float *q_off, const float *params_dev,
                          const float *lse, const float *g_lse, const float *g_next, const int32_t *idx, int64_t B,
                          int64_t K, float *g_x_prev, float *g_params, void *stream)
{
    const char *fn = "aesmc_lg_step_bwd_f32";
    REQUIRE(x && y && params_dev && lse && g_lse && g_params, fn);
    REQUIRE((g_next == nullptr) == (idx == nullptr), fn);
    REQUIRE((x_prev == nullptr) == (g_x_prev == nullptr), fn);
    REQUIRE(B >= 0 && K >= 1 && B <= kMaxDim && K <= 16384, fn);
    if (B == 0) return AESMC_OK;
    return launch_lg_step_bwd(x, x_prev, y, q_off, params_dev, lse, g_lse, g_next, idx, B, K, g_x_prev, g_params, S(stream));
}

int aesmc_resample_from_weights_f32(const float *w, const double *u, int64_t B, int64_t K, int32_t *idx,
                                    int32_t *flags, int mode, void *stream)
{
    const char *fn = "aesmc_resample_from_weights_f32";
    REQUIRE(w && u && idx && flags, fn);
    REQUIRE(B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    REQUIRE(mode == AESMC_MODE_EXACT || mode == AESMC_MODE_FAST, fn);
    if (B == 0) return AESMC_OK;
    return launch_smc_step(w, nullptr, nullptr, u, B, K, nullptr, nullptr, idx, nullptr, nullptr, 1, flags, mode, 1, nullptr, 0,
                           S(stream));
}

int aesmc_resample_from_cdf_f32(const float *cdf, const double *u, int64_t B, int64_t K, int32_t *idx,
                                int32_t *flags, void *stream)
{
    const char *fn = "aesmc_resample_from_cdf_f32";
    REQUIRE(cdf && u && idx && flags, fn);
    REQUIRE(B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_smc_step(cdf, nullptr, nullptr, u, B, K, nullptr, nullptr, idx, nullptr, nullptr, 1, flags,
                           AESMC_MODE_EXACT, 2, nullptr, 0, S(stream));
}

int aesmc_is_accumulate_f32(const float *lp_a, const float *lp_b, const float *lp_c, float *acc, float *log_w,
                            int64_t n, int first, void *stream)
{
    const char *fn = "aesmc_is_accumulate_f32";
    REQUIRE(lp_a && acc && n >= 0, fn);
    if (n == 0) return AESMC_OK;
    return launch_is_accumulate_f32(lp_a, lp_b, lp_c, acc, log_w, n, first, S(stream));
}

int aesmc_logsumexp_f32(const float *log_w, int64_t B, int64_t K, float *lse, int32_t *flags, void *stream)
{
    const char *fn = "aesmc_logsumexp_f32";
    REQUIRE(log_w && lse && B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_logsumexp_f32(log_w, B, K, lse, flags, S(stream));
}
int aesmc_logsumexp_f64(const double *log_w, int64_t B, int64_t K, double *lse, int32_t *flags, void *stream)
{
    const char *fn = "aesmc_logsumexp_f64";
    REQUIRE(log_w && lse && B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_logsumexp_f64(log_w, B, K, lse, flags, S(stream));
}

int aesmc_step_bwd_f32(const float *log_w, const float *lse, const float *g_log_w, const float *g_lse, int64_t B,
                       int64_t K, float *g_pos, float *g_neg, void *stream)
{
    const char *fn = "aesmc_step_bwd_f32";
    REQUIRE(g_pos && B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    REQUIRE(g_lse == nullptr || (log_w && lse), fn);
    if (B == 0) return AESMC_OK;
    return launch_step_bwd_f32(log_w, lse, g_log_w, g_lse, B, K, g_pos, g_neg, S(stream));
}

int aesmc_lognormexp_f32(const float *log_w, int64_t B, int64_t K, float *out, int exponentiate, void *stream)
{
    const char *fn = "aesmc_lognormexp_f32";
    REQUIRE(log_w && out && B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_lognormexp_f32(log_w, B, K, out, exponentiate, S(stream));
}

int aesmc_gather_bytes(const void *src, const void *idx, int idx_is_i64, int64_t B, int64_t K, int64_t row_bytes,
                       void *dst, int32_t *flags, void *stream)
{
    const char *fn = "aesmc_gather_bytes";
    REQUIRE(src && idx && dst && src != dst, fn);
    REQUIRE(B >= 0 && K >= 1 && row_bytes >= 1 && B <= kMaxDim && K * row_bytes <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_gather_bytes(src, idx, idx_is_i64, B, K, row_bytes, dst, flags, S(stream));
}

int aesmc_gather_bwd_f32(const float *gdst, const void *idx, int idx_is_i64, int64_t B, int64_t K, int64_t D,
                         float *gsrc, int sorted, void *stream)
{
    const char *fn = "aesmc_gather_bwd_f32";
    REQUIRE(gdst && idx && gsrc && B >= 0 && K >= 1 && D >= 1 && B <= kMaxDim && K * D <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_gather_bwd_f32(gdst, idx, idx_is_i64, B, K, D, gsrc, sorted, S(stream));
}
int aesmc_gather_bwd_f64(const double *gdst, const void *idx, int idx_is_i64, int64_t B, int64_t K, int64_t D,
                         double *gsrc, int sorted, void *stream)
{
    const char *fn = "aesmc_gather_bwd_f64";
    REQUIRE(gdst && idx && gsrc && B >= 0 && K >= 1 && D >= 1 && B <= kMaxDim && K * D <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_gather_bwd_f64(gdst, idx, idx_is_i64, B, K, D, gsrc, sorted, S(stream));
}

int aesmc_compose_index_i32(const int32_t *prev, const int32_t *cur, int64_t B, int64_t K, int32_t *out, void *stream)
{
    const char *fn = "aesmc_compose_index_i32";
    REQUIRE(prev && cur && out && out != prev && B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_compose_index(prev, cur, B, K, out, S(stream));
}
int aesmc_iota_index_i32(int64_t B, int64_t K, int32_t *out, void *stream)
{
    const char *fn = "aesmc_iota_index_i32";
    REQUIRE(out && B >= 0 && K >= 1 && K <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_iota_index(B, K, out, S(stream));
}
int aesmc_index_widen(const int32_t *in, int64_t *out, int64_t n, void *stream)
{
    const char *fn = "aesmc_index_widen";
    REQUIRE(in && out && n >= 0, fn);
    if (n == 0) return AESMC_OK;
    return launch_index_widen(in, out, n, S(stream));
}
int aesmc_index_narrow(const int64_t *in, int32_t *out, int64_t n, void *stream)
{
    const char *fn = "aesmc_index_narrow";
    REQUIRE(in && out && n >= 0, fn);
    if (n == 0) return AESMC_OK;
    return launch_index_narrow(in, out, n, S(stream));
}

int aesmc_normal_log_prob_f32(const float *value, int value_kind, const float *loc, int loc_kind, float loc_host,
                              const float *scale_dev, float inv_two_var_host, float log_scale_host, float half_log_2pi,
                              int64_t B, int64_t K, float *out, void *stream)
{
    const char *fn = "aesmc_normal_log_prob_f32";
    REQUIRE(value && out && B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    REQUIRE(value_kind >= 0 && value_kind <= 2 && loc_kind >= 0 && loc_kind <= 2, fn);
    if (B == 0) return AESMC_OK;
    return normal_log_prob_f32(value, value_kind, loc, loc_kind, loc_host, scale_dev, inv_two_var_host, log_scale_host,
                               half_log_2pi, B, K, out, S(stream));
}

int aesmc_normal_log_prob_bwd_f32(const float *value, int value_kind, const float *loc, int loc_kind, float loc_host,
                                  const float *scale_dev, float scale_host, const float *g, int64_t B, int64_t K,
                                  float *g_value, float *g_loc, float *g_scale, void *stream)
{
    const char *fn = "aesmc_normal_log_prob_bwd_f32";
    REQUIRE(value && g && B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    REQUIRE(value_kind >= 0 && value_kind <= 2 && loc_kind >= 0 && loc_kind <= 2, fn);
    if (B == 0) return AESMC_OK;
    return normal_log_prob_bwd_f32(value, value_kind, loc, loc_kind, loc_host, scale_dev, scale_host, g, B, K, g_value,
                                   g_loc, g_scale, S(stream));
}

int aesmc_selftest_expf(uint64_t *out2, void *stream)
{
    const char *fn = "aesmc_selftest_expf";
    REQUIRE(out2 != nullptr, fn);
    return launch_selftest_expf(reinterpret_cast<unsigned long long *>(out2), S(stream));
}

int aesmc_debug_force_rare_paths(int bits)
{
    static int current = 0;
    const int prev = current;
    current = bits;
    smc_step_x_set_debug_force(bits);
    return prev;
}

int aesmc_log_ess_f32(const float *log_w, int64_t B, int64_t K, float *out, void *stream)
{
    const char *fn = "aesmc_log_ess_f32";
    REQUIRE(log_w && out && B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_log_ess_f32(log_w, B, K, out, S(stream));
}
int aesmc_log_ess_f64(const double *log_w, int64_t B, int64_t K, double *out, void *stream)
{
    const char *fn = "aesmc_log_ess_f64";
    REQUIRE(log_w && out && B >= 0 && K >= 1 && B <= kMaxDim && K <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_log_ess_f64(log_w, B, K, out, S(stream));
}

int aesmc_weighted_moments_f32(const float *x, const float *log_w, int64_t B, int64_t K, int64_t D, float *mean,
                               float *second, void *stream)
{
    const char *fn = "aesmc_weighted_moments_f32";
    REQUIRE(x && log_w && mean && B >= 0 && K >= 1 && D >= 1 && B <= kMaxDim && K * D <= kMaxDim, fn);
    if (B == 0) return AESMC_OK;
    return launch_weighted_moments_f32(x, log_w, B, K, D, mean, second, S(stream));
}

} // extern "C"
