// row_stats.cu -- single-pass float32 row statistics, ONE WARP PER ROW (HBM-bound streaming kernels):
//   logsumexp over particles        inference.py:130,158 / statistics.py:90-91
//   log effective sample size       statistics.py:79-91      2 lse(lw) - lse(2 lw)
//   weighted first / second moment  statistics.py:7-76 with f = x, x^2 (scalar latents)
//
// The CTA-per-row kernels in reduce.cu read a row twice (max, then sum) with two block reductions per pass and
// reached 0.42-0.46 of the measured HBM bandwidth at B = K = 4096.  Here a row never leaves its warp: every lane
// streams 16-byte chunks (512 contiguous bytes per warp instruction, eight of them in flight per lane) through
// an online (max, sum) pair that is rescaled whenever the running maximum grows -- one read of the data, no shared
// memory, no barrier -- and the 32 partial pairs are merged by shuffles at the end of the row.
// Used when there are enough rows to fill the machine with warps; otherwise reduce.cu's kernels serve.
#include "common.cuh"

namespace aesmc {

namespace {


__device__ __forceinline__ float4 ld_cs(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

enum { kLse = 0, kLogEss = 1, kMoments = 2 };
// 16-byte loads per lane and pipeline stage (two stages: up to 8 loads in flight per lane, one table x 4 x 2 or two
// tables x 2 x 2) -- with ~28 warps per SM at B = 4096 that keeps ~110 KB per SM in flight, what the HBM
// latency-bandwidth product asks for
template <int MODE> struct Unroll { static constexpr int value = MODE == kMoments ? 2 : 4; }; // per pipeline stage

// Online state of one lane: running maximum m and sums of e = exp(v - m) (s0), and mode-dependent companions
// (s1, s2): e^2 for log-ESS; e x and e x^2 for the moments.
template <int MODE> struct Partial {
    float m, s0, s1, s2;
    int bad;
    __device__ __forceinline__ void init() { m = -INFINITY; s0 = s1 = s2 = 0.f; bad = 0; }
    __device__ __forceinline__ void rescale(float m_new)
    {
        if (m_new > m) { // exp(-inf - finite) = 0 wipes the (empty) sums of an all -inf prefix
            const float f = expf(m - m_new);
            s0 *= f;
            if (MODE == kLogEss) s1 *= f * f;
            if (MODE == kMoments) { s1 *= f; s2 *= f; }
            m = m_new;
        }
    }
    __device__ __forceinline__ void add(float v, float x)
    {
        const float e = expf(v - m);
        s0 += e;
        if (MODE == kLogEss) s1 = fmaf(e, e, s1);
        if (MODE == kMoments) { s1 = fmaf(e, x, s1); s2 = fmaf(e * x, x, s2); }
    }
};

// WPR warps share a row (chunks interleaved warp by warp): with one warp per row B = 4096 rows fill only 43 % of the
// 64 warp slots of each SM, in a single wave whose warps each wait for four dependent rounds of loads; two warps
// per row double the bytes in flight and halve the rounds.  The WPR partial (max, sums) pairs meet in shared memory.
template <int MODE, int WPR>
__global__ void __launch_bounds__(256) row_stats_warp_kernel(const float *__restrict__ lw, const float *__restrict__ x,
                                                             int B, int K, float *__restrict__ out0,
                                                             float *__restrict__ out1, int32_t *flags)
{
    constexpr int kWarps = 8, kRows = kWarps / WPR; // rows per CTA and pass
    __shared__ float part[kWarps][4];
    __shared__ int part_bad[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = warp % WPR, rl = warp / WPR;
    const int nch = K >> 2;
    constexpr int kUnroll = Unroll<MODE>::value;
    for (int row0 = blockIdx.x * kRows; row0 < B; row0 += gridDim.x * kRows) { // CTA-uniform trip count
        const int row = row0 + rl;
        Partial<MODE> a;
        a.init();
        if (row < B) {
            const float4 *__restrict__ r4 = reinterpret_cast<const float4 *>(lw + (size_t)row * K);
            const float4 *__restrict__ x4 = MODE == kMoments ? reinterpret_cast<const float4 *>(x + (size_t)row * K) : nullptr;
            // software pipeline: the loads of the next group are in flight while this one is reduced, so the warp
            // never sits through a full HBM round trip with nothing outstanding
            constexpr int kStep = 32 * WPR * kUnroll;
            float4 v[kUnroll], xv[kUnroll], nv[kUnroll], nxv[kUnroll];
            auto fetch = [&](int c0, float4 (&dst)[kUnroll], float4 (&xdst)[kUnroll]) {
#pragma unroll
                for (int q = 0; q < kUnroll; ++q) {
                    const int c = c0 + 32 * WPR * q;
                    dst[q] = c < nch ? ld_cs(r4 + c) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                    if (MODE == kMoments) xdst[q] = c < nch ? ld_cs(x4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            fetch(lane + 32 * sub, nv, nxv);
            for (int c0 = lane + 32 * sub; c0 < nch; c0 += kStep) {
#pragma unroll
                for (int q = 0; q < kUnroll; ++q) { v[q] = nv[q]; if (MODE == kMoments) xv[q] = nxv[q]; }
                if (c0 + kStep < nch) fetch(c0 + kStep, nv, nxv);
                float cm = -INFINITY;
#pragma unroll
                for (int q = 0; q < kUnroll; ++q) {
                    a.bad |= (v[q].x != v[q].x) | (v[q].y != v[q].y) | (v[q].z != v[q].z) | (v[q].w != v[q].w);
                    cm = fmaxf(fmaxf(cm, fmaxf(v[q].x, v[q].y)), fmaxf(v[q].z, v[q].w));
                }
                a.rescale(cm);
                if (a.m > -INFINITY && a.m < INFINITY) {
#pragma unroll
                    for (int q = 0; q < kUnroll; ++q) {
                        a.add(v[q].x, MODE == kMoments ? xv[q].x : 0.f);
                        a.add(v[q].y, MODE == kMoments ? xv[q].y : 0.f);
                        a.add(v[q].z, MODE == kMoments ? xv[q].z : 0.f);
                        a.add(v[q].w, MODE == kMoments ? xv[q].w : 0.f);
                    }
                }
            }
        }
        // merge the 32 lanes: common maximum, rescale, sum
        float m = warp_max(a.m);
        int bad = __any_sync(kFull, a.bad);
        float f = (a.m > -INFINITY && m < INFINITY) ? expf(a.m - m) : 0.f;
        float s0 = warp_sum(a.s0 * f);
        float s1 = 0.f, s2 = 0.f;
        if (MODE == kLogEss) s1 = warp_sum(a.s1 * f * f);
        if (MODE == kMoments) { s1 = warp_sum(a.s1 * f); s2 = warp_sum(a.s2 * f); }
        if (WPR > 1) { // merge the row's warps
            if (lane == 0) { part[warp][0] = m; part[warp][1] = s0; part[warp][2] = s1; part[warp][3] = s2; part_bad[warp] = bad; }
            __syncthreads();
            if (sub == 0 && lane == 0) {
                float M = m;
#pragma unroll
                for (int j = 1; j < WPR; ++j) M = fmaxf(M, part[warp + j][0]);
                float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int j = 0; j < WPR; ++j) {
                    const float mj = part[warp + j][0];
                    const float g = (mj > -INFINITY && M < INFINITY) ? expf(mj - M) : 0.f;
                    t0 += part[warp + j][1] * g;
                    t1 += part[warp + j][2] * (MODE == kLogEss ? g * g : g);
                    t2 += part[warp + j][3] * g;
                    bad |= part_bad[warp + j];
                }
                m = M; s0 = t0; s1 = t1; s2 = t2;
            }
            __syncthreads();
        }
        if (sub == 0 && lane == 0 && row < B) {
            if (MODE == kLse) {
                float o;
                if (bad) {
                    o = NAN;
                    if (flags) atomicOr(flags, AESMC_FLAG_NAN);
                } else if (!(fabsf(m) < INFINITY)) {
                    o = m; // all -inf -> -inf ; +inf present -> +inf (torch.logsumexp convention)
                } else {
                    o = m + logf(s0);
                }
                out0[row] = o;
            } else if (MODE == kLogEss) {
                out0[row] = bad ? NAN : 2.f * logf(s0) - logf(s1); // 2 (m + log s0) - (2 m + log s1)
            } else {
                const float inv = 1.f / s0; // softmax weights sum to one: divide once per row
                out0[row] = bad ? NAN : s1 * inv;
                if (out1) out1[row] = bad ? NAN : s2 * inv;
            }
        }
    }
}

int sm_count_rs()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int MODE, int WPR> int ctas_per_sm()
{
    static int n = 0;
    if (n == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, row_stats_warp_kernel<MODE, WPR>, 256, 0);
        if (n < 1) n = 1;
    }
    return n;
}

template <int MODE, int WPR>
int launch_wpr(const float *lw, const float *x, int64_t B, int64_t K, float *out0, float *out1, int32_t *flags,
               cudaStream_t st, const char *what)
{
    const int rows_per_cta = 8 / WPR;
    long long grid = (B + rows_per_cta - 1) / rows_per_cta;
    const long long cap = (long long)sm_count_rs() * ctas_per_sm<MODE, WPR>();
    if (grid > cap) grid = cap;
    row_stats_warp_kernel<MODE, WPR><<<(unsigned)grid, 256, 0, st>>>(lw, x, (int)B, (int)K, out0, out1, flags);
    count_launch();
    return check_launch(what);
}

// as many warps per row (1, 2 or 4) as still fit the resident warp slots in one wave
template <int MODE>
int launch(const float *lw, const float *x, int64_t B, int64_t K, float *out0, float *out1, int32_t *flags,
           cudaStream_t st, const char *what)
{
    const long long sms = sm_count_rs();
    if (K >= 4096 && B * 4 <= sms * ctas_per_sm<MODE, 4>() * 8) return launch_wpr<MODE, 4>(lw, x, B, K, out0, out1, flags, st, what);
    if (K >= 2048 && B * 2 <= sms * ctas_per_sm<MODE, 2>() * 8) return launch_wpr<MODE, 2>(lw, x, B, K, out0, out1, flags, st, what);
    return launch_wpr<MODE, 1>(lw, x, B, K, out0, out1, flags, st, what);
}

} // namespace

// enough rows to give every SM a few warps, rows short enough that one warp streams them quickly, 16-byte access
bool row_stats_warp_supported(const void *lw, const void *x, int64_t B, int64_t K)
{
    if ((K & 3) != 0 || K > (1 << 20)) return false;
    if (((reinterpret_cast<uintptr_t>(lw) | reinterpret_cast<uintptr_t>(x)) & 15) != 0) return false;
    if (K > 65536) return false;
    return B >= 4LL * sm_count_rs() || (K >= 4096 && B >= sm_count_rs()); // (long rows: up to four warps each)
}
int launch_logsumexp_warp_f32(const float *lw, int64_t B, int64_t K, float *lse, int32_t *flags, cudaStream_t st)
{
    return launch<kLse>(lw, nullptr, B, K, lse, nullptr, flags, st, "row_stats_warp_kernel<lse>");
}
int launch_log_ess_warp_f32(const float *lw, int64_t B, int64_t K, float *out, cudaStream_t st)
{
    return launch<kLogEss>(lw, nullptr, B, K, out, nullptr, nullptr, st, "row_stats_warp_kernel<log_ess>");
}
int launch_weighted_moments_warp_f32(const float *x, const float *lw, int64_t B, int64_t K, float *mean, float *second,
                                     cudaStream_t st)
{
    return launch<kMoments>(lw, x, B, K, mean, second, nullptr, st, "row_stats_warp_kernel<moments>");
}

} // namespace aesmc
