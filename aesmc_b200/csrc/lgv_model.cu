// lgv_model.cu -- proposal sampling and the three log-densities of a D-dimensional linear-Gaussian state-space
// model with diagonal noise (BASELINE config 3), one thread per particle:
//
//     x_0 ~ N(m0, diag s0^2)      x_t | x_{t-1} ~ N(A x_{t-1} + b, diag sx^2)      y_t | x_t ~ N(C x_t + d, diag sy^2)
//     q(x_t | x_{t-1}, y_t) = N(Wx x_{t-1} + q_row(y_t), diag sq^2)               (or the prior dynamics: bootstrap)
//
// This is what the user model's torch callables do per time step in inference.py:108-126 of the reference
// (nn.Linear / matmul on [B, K, D], Independent(Normal).rsample, three log_prob's, (transition + emission) -
// proposal): ~40 torch kernels and ~2 KB of HBM traffic per particle-step, here one launch and 8 D + 4 bytes.  The
// matrices travel as kernel parameters, i.e. they sit in the constant bank and feed the FFMAs as immediate
// operands; everything is zero-padded to a multiple of four dimensions at compile time.  The log-weights then go
// through the unchanged step kernel (lse, systematic ancestors, D-float gather).  The D x D mat-vecs run on the FMA
// pipe in registers: at ~500 flops per 84 bytes the kernel is still HBM-bound, tensor cores would not help.
#include "common.cuh"
#include "lg_model.cuh"

namespace aesmc {

template <int DP, int DYP> struct LgvParams {
    float A[DP][DP], Wx[DP][DP], C[DYP][DP];
    float b[DP], sq[DP];
    float t_i2v[DP], t_ln[DP], q_i2v[DP], q_ln[DP]; // 1 / (2 sigma^2) and log sigma + log sqrt(2 pi), per dimension
    float d[DYP], e_i2v[DYP], e_ln[DYP];
};

template <int DP, int DYP>
__global__ void __launch_bounds__(128) lgv_propose_kernel(const float *__restrict__ x_prev, const float *__restrict__ y,
                                                          const float *__restrict__ noise, const float *__restrict__ q_row,
                                                          const __grid_constant__ LgvParams<DP, DYP> P, int D, int Dy,
                                                          int bootstrap, int pairs, unsigned long long seed, unsigned long long stream_offset,
                                                          int K, int tiles_per_row, float *__restrict__ x_new,
                                                          float *__restrict__ log_w)
{
    const int row = blockIdx.x / tiles_per_row, k = (blockIdx.x - row * tiles_per_row) * 128 + threadIdx.x;
    if (k >= K) return;
    const size_t particle = (size_t)row * K + k;
    float xp[DP], eps[DP], x[DP];
#pragma unroll
    for (int i = 0; i < DP; ++i) xp[i] = eps[i] = 0.f;
    if (x_prev) {
        if (pairs) {
            const float2 *s = reinterpret_cast<const float2 *>(x_prev + particle * D);
#pragma unroll
            for (int i = 0; i < DP; i += 2)
                if (i < D) { const float2 v = __ldg(s + (i >> 1)); xp[i] = v.x; xp[i + 1] = v.y; }
        } else {
#pragma unroll
            for (int i = 0; i < DP; ++i)
                if (i < D) xp[i] = __ldg(x_prev + particle * D + i);
        }
    }
    if (noise) {
        if (pairs) {
            const float2 *s = reinterpret_cast<const float2 *>(noise + particle * D);
#pragma unroll
            for (int i = 0; i < DP; i += 2)
                if (i < D) { const float2 v = __ldg(s + (i >> 1)); eps[i] = v.x; eps[i + 1] = v.y; }
        } else {
#pragma unroll
            for (int i = 0; i < DP; ++i)
                if (i < D) eps[i] = __ldg(noise + particle * D + i);
        }
    } else {
#pragma unroll
        for (int i = 0; i < DP; i += 4)
            if (i < D) {
                const float4 v = philox_normal4(seed, stream_offset, particle * (DP / 4) + (i >> 2));
                eps[i] = v.x; eps[i + 1] = v.y; eps[i + 2] = v.z; eps[i + 3] = v.w;
            }
    }
    float lt = 0.f, lq = 0.f;
#pragma unroll
    for (int i = 0; i < DP; ++i) {
        x[i] = 0.f;
        if (i < D) {
            float mt = P.b[i], mq = bootstrap ? 0.f : __ldg(q_row + (size_t)row * D + i);
#pragma unroll
            for (int j = 0; j < DP; ++j) {
                mt = fmaf(P.A[i][j], xp[j], mt);
                mq = fmaf(P.Wx[i][j], xp[j], mq);
            }
            if (bootstrap) mq = mt;
            const float xi = fmaf(eps[i], P.sq[i], mq); // Normal.rsample
            x[i] = xi;
            const float rq = xi - mq, rt = xi - mt;
            lq += -(rq * rq) * P.q_i2v[i] - P.q_ln[i];
            lt += -(rt * rt) * P.t_i2v[i] - P.t_ln[i];
        }
    }
    float le = 0.f;
#pragma unroll
    for (int m = 0; m < DYP; ++m) {
        if (m < Dy) {
            float me = P.d[m];
#pragma unroll
            for (int j = 0; j < DP; ++j) me = fmaf(P.C[m][j], x[j], me);
            const float r = __ldg(y + (size_t)row * Dy + m) - me;
            le += -(r * r) * P.e_i2v[m] - P.e_ln[m];
        }
    }
    log_w[particle] = (lt + le) - lq; // inference.py:125-126
    if (pairs) {
        float2 *o = reinterpret_cast<float2 *>(x_new + particle * D);
#pragma unroll
        for (int i = 0; i < DP; i += 2)
            if (i < D) o[i >> 1] = make_float2(x[i], x[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < DP; ++i)
            if (i < D) x_new[particle * D + i] = x[i];
    }
}

// params_host: A [D*D] | b [D] | sx [D] | C [Dy*D] | d [Dy] | sy [Dy] | Wx [D*D] | sq [D]   (row-major, HOST memory)
template <int DP, int DYP>
static int launch_lgv(const float *x_prev, const float *y, const float *noise, const float *q_row, const float *h, int D,
                      int Dy, int bootstrap, unsigned long long seed, unsigned long long stream_offset, int64_t B, int64_t K,
                      float *x_new, float *log_w, cudaStream_t stream)
{
    LgvParams<DP, DYP> P = {};
    const float *A = h, *b = A + D * D, *sx = b + D, *C = sx + D, *d = C + Dy * D, *sy = d + Dy, *Wx = sy + Dy, *sq = Wx + D * D;
    const float c = 0.91893853320467274178f; // log sqrt(2 pi)
    for (int i = 0; i < D; ++i) {
        for (int j = 0; j < D; ++j) { P.A[i][j] = A[i * D + j]; P.Wx[i][j] = bootstrap ? 0.f : Wx[i * D + j]; }
        P.b[i] = b[i];
        const float s_t = sx[i], s_q = bootstrap ? sx[i] : sq[i];
        P.sq[i] = s_q;
        P.t_i2v[i] = 1.0f / (2.0f * s_t * s_t); P.t_ln[i] = logf(s_t) + c;
        P.q_i2v[i] = 1.0f / (2.0f * s_q * s_q); P.q_ln[i] = logf(s_q) + c;
    }
    for (int m = 0; m < Dy; ++m) {
        for (int j = 0; j < D; ++j) P.C[m][j] = C[m * D + j];
        P.d[m] = d[m];
        P.e_i2v[m] = 1.0f / (2.0f * sy[m] * sy[m]); P.e_ln[m] = logf(sy[m]) + c;
    }
    // float2 traffic when every particle's D floats start on an 8-byte boundary
    const int pairs = (D % 2 == 0) && (((uintptr_t)x_prev | (uintptr_t)noise | (uintptr_t)x_new) % 8 == 0);
    const int tiles = (int)((K + 127) / 128);
    const long long grid = (long long)tiles * B;
    if (grid > 2147483647LL) { set_error("aesmc_lgv_propose_f32: B * K too large"); return AESMC_ERR_BAD_ARG; }
    lgv_propose_kernel<DP, DYP><<<(unsigned)grid, 128, 0, stream>>>(x_prev, y, noise, q_row, P, D, Dy, bootstrap, pairs, seed,
                                                                     stream_offset, (int)K, tiles, x_new, log_w);
    count_launch();
    return check_launch("lgv_propose_kernel");
}

template <int DP>
static int dispatch_lgv_dy(const float *x_prev, const float *y, const float *noise, const float *q_row, const float *h, int D,
                           int Dy, int bootstrap, unsigned long long seed, unsigned long long so, int64_t B, int64_t K,
                           float *x_new, float *log_w, cudaStream_t stream)
{
    if (Dy <= 4) return launch_lgv<DP, 4>(x_prev, y, noise, q_row, h, D, Dy, bootstrap, seed, so, B, K, x_new, log_w, stream);
    if (Dy <= 8) return launch_lgv<DP, 8>(x_prev, y, noise, q_row, h, D, Dy, bootstrap, seed, so, B, K, x_new, log_w, stream);
    if (Dy <= 12) return launch_lgv<DP, 12>(x_prev, y, noise, q_row, h, D, Dy, bootstrap, seed, so, B, K, x_new, log_w, stream);
    return launch_lgv<DP, 16>(x_prev, y, noise, q_row, h, D, Dy, bootstrap, seed, so, B, K, x_new, log_w, stream);
}

int launch_lgv_propose(const float *x_prev, const float *y, const float *noise, const float *q_row, const float *params_host,
                       int64_t D, int64_t Dy, int bootstrap, unsigned long long seed, unsigned long long stream_offset,
                       int64_t B, int64_t K, float *x_new, float *log_w, cudaStream_t stream)
{
    const int d = (int)D, dy = (int)Dy;
    if (d <= 4) return dispatch_lgv_dy<4>(x_prev, y, noise, q_row, params_host, d, dy, bootstrap, seed, stream_offset, B, K, x_new, log_w, stream);
    if (d <= 8) return dispatch_lgv_dy<8>(x_prev, y, noise, q_row, params_host, d, dy, bootstrap, seed, stream_offset, B, K, x_new, log_w, stream);
    if (d <= 12) return dispatch_lgv_dy<12>(x_prev, y, noise, q_row, params_host, d, dy, bootstrap, seed, stream_offset, B, K, x_new, log_w, stream);
    return dispatch_lgv_dy<16>(x_prev, y, noise, q_row, params_host, d, dy, bootstrap, seed, stream_offset, B, K, x_new, log_w, stream);
}

} // namespace aesmc
