// lg_bwd.cu -- backward of ONE time step of the fused scalar linear-Gaussian SMC step (aesmc_smc_step_lg_*):
// what torch autograd does, through ~60 elementwise kernels and a scatter_add per step, for
//     losses.get_loss(..., 'aesmc')  (losses.py:5-65 -> inference.py:99-134) on the reference's trainable LGSSM
//     (test/models/lgssm.py:19-72: learnable transition / emission multipliers, affine proposal in (x_prev, y)).
//
// Forward of step t (per row b, particle k), with the proposal reparameterised (state.sample -> rsample):
//     x   = mu_q + eps * s_q,                  mu_q = q.mult * xp + q_off[b]         (xp = resampled x of step t-1)
//     lw  = log N(x; mu_t, s_t) + log N(y[b]; mu_e, s_e) - log N(x; mu_q, s_q),   mu_t = t.mult * xp + t.off,  mu_e = e.mult * x + e.off
//     lse = logsumexp_k lw                      -> log-evidence term;   xp' = x[idx]  -> next step
// Backward, given  gl = dL/dlse[b]  and  G[k] = dL/dxp'[k]  of the NEXT step (resampling indices carry no gradient,
// inference.py:254 detaches the weights):
//     S[j]  = sum over the children k of particle j (idx[k] == j) of G[k]            (the gather's scatter-add)
//     w     = gl * exp(lw - lse)                                                      (the logsumexp gradient)
//     r_t   = (x - mu_t) / s_t^2,   r_e = (y - mu_e) / s_e^2
//     gx    = S + w * (-r_t + e.mult * r_e)          dL/dx; the proposal density drops out: along x = mu_q + eps s_q
//                                                    its x- and mu_q-derivatives cancel exactly (fixed scales)
//     dL/dxp   = gx * q.mult + w * r_t * t.mult
//     per-row parameter sums: t.mult: w r_t xp | t.off: w r_t | e.mult: w r_e x | e.off: w r_e | q.mult: gx xp | q_off[b]: gx
// One CTA per row: the ancestors of a row are sorted, so a warp first folds equal neighbours (segmented shuffle scan)
// and only run tails touch the shared-memory accumulator; then one pass over (x, xp).  20 bytes per particle-step.
#include "common.cuh"
#include "lg_model.cuh"

namespace aesmc {

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) lg_step_bwd_kernel(
    const float *__restrict__ x, const float *__restrict__ xp, const float *__restrict__ y, const float *__restrict__ q_off,
    const float *__restrict__ params, const float *__restrict__ lse, const float *__restrict__ g_lse,
    const float *__restrict__ g_next, const int32_t *__restrict__ idx, int B, int K, float *__restrict__ g_xp,
    float *__restrict__ g_params)
{
    extern __shared__ __align__(16) float S[]; // [K] gradient arriving at each particle from its children
    __shared__ float red[6][kThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const LgAffine mt = lg_load_affine(params), me = lg_load_affine(params + 5), mq = lg_load_affine(params + 10);
    const bool bootstrap = (q_off == nullptr) && lg_same(mq, mt); // proposal == transition: log p(x | xp) - log q = 0
    const float it = 2.0f / mt.two_var, ie = 2.0f / me.two_var, iq = 2.0f / mq.two_var; // 1 / s^2

    for (int row = blockIdx.x; row < B; row += gridDim.x) {
        const size_t off = (size_t)row * K;
        for (int k = tid; k < K; k += kThreads) S[k] = 0.f;
        __syncthreads();
        if (g_next != nullptr) {
            for (int k0 = 0; k0 < K; k0 += kThreads) {
                const int k = k0 + tid;
                const bool valid = k < K;
                const int j = valid ? idx[off + k] : -1 - lane; // invalid lanes never match a neighbour
                float g = valid ? g_next[off + k] : 0.f;
                const int jp = __shfl_up_sync(kFull, j, 1), jn = __shfl_down_sync(kFull, j, 1);
                int head = (lane == 0) || (j != jp);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float gv = __shfl_up_sync(kFull, g, o);
                    const int hv = __shfl_up_sync(kFull, head, o);
                    if (lane >= o && !head) { g += gv; head = hv; }
                }
                if (valid && ((lane == 31) || (j != jn))) atomicAdd(&S[min(max(j, 0), K - 1)], g);
            }
            __syncthreads();
        }
        const float yv = y[row], qo = q_off ? q_off[row] : mq.off, l = lse[row], gl = g_lse[row];
        float sa = 0.f, sb = 0.f, sc = 0.f, sd = 0.f, sqx = 0.f, sq1 = 0.f;
        for (int k = tid; k < K; k += kThreads) {
            const float xv = x[off + k], xpv = xp ? xp[off + k] : 0.f;
            const float dt = xv - fmaf(mt.mult, xpv, mt.off);
            const float de = yv - fmaf(me.mult, xv, me.off);
            const float dq = xv - fmaf(mq.mult, xpv, qo);
            float lw = -0.5f * ie * de * de;
            if (!bootstrap) lw += 0.5f * (iq * dq * dq - it * dt * dt) + (mq.log_scale - mt.log_scale);
            lw -= me.log_scale + 0.9189385332046727f;
            const float w = gl * expf(lw - l);
            const float rt = bootstrap ? 0.f : dt * it, re = de * ie;
            const float gx = S[k] + w * fmaf(me.mult, re, -rt);
            if (g_xp) g_xp[off + k] = fmaf(gx, mq.mult, w * rt * mt.mult);
            sa = fmaf(w * rt, xpv, sa);
            sb = fmaf(w, rt, sb);
            sc = fmaf(w * re, xv, sc);
            sd = fmaf(w, re, sd);
            sqx = fmaf(gx, xpv, sqx);
            sq1 += gx;
        }
        float v[6] = {sa, sb, sc, sd, sqx, sq1};
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const float r = warp_sum(v[i]);
            if (lane == 0) red[i][warp] = r;
        }
        __syncthreads();
        if (tid < 6) {
            float r = 0.f;
#pragma unroll
            for (int wq = 0; wq < kThreads / 32; ++wq) r += red[tid][wq];
            g_params[(size_t)row * 6 + tid] = r;
        }
        __syncthreads(); // S and red are reused by the next row
    }
}

} // namespace

int launch_lg_step_bwd(const float *x, const float *xp, const float *y, const float *q_off, const float *params,
                       const float *lse, const float *g_lse, const float *g_next, const int32_t *idx, int64_t B, int64_t K,
                       float *g_xp, float *g_params, cudaStream_t st)
{
    const size_t smem = (size_t)K * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(lg_step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
        configured = true;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lg_step_bwd_kernel, kThreads, smem);
    long long grid = (long long)sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > B) grid = B;
    lg_step_bwd_kernel<<<(unsigned)grid, kThreads, smem, st>>>(x, xp, y, q_off, params, lse, g_lse, g_next, idx, (int)B, (int)K,
                                                              g_xp, g_params);
    count_launch();
    return check_launch("lg_step_bwd_kernel");
}

} // namespace aesmc
