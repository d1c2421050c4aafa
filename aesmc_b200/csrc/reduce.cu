// reduce.cu -- row reductions and elementwise companions of the SMC step (HBM-bound, no tensor cores):
//   logsumexp over particles      inference.py:130,158 / statistics.py:90-91
//   step backward                  autograd of torch.logsumexp + the (a + b) - c sum
//   lognormexp / exponentiate_and_normalize   math.py:6-51 (torch branch)
//   log_ess                        statistics.py:79-91
//   weighted moments               statistics.py:7-76 with f = x, x^2
//   importance-sampling accumulate inference.py:156-157
#include "common.cuh"
#include <initializer_list>

namespace aesmc {

template <typename T> struct Acc;
template <> struct Acc<float> {
    using OpMax = OpMaxF; using OpSum = OpSumF;
    __device__ static float ninf() { return -INFINITY; }
    __device__ static float ex(float x) { return expf(x); }
    __device__ static float lg(float x) { return logf(x); }
    __device__ static float mx(float a, float b) { return fmaxf(a, b); }
};
template <> struct Acc<double> {
    using OpMax = OpMaxD; using OpSum = OpSumD;
    __device__ static double ninf() { return -INFINITY; }
    __device__ static double ex(double x) { return exp(x); }
    __device__ static double lg(double x) { return log(x); }
    __device__ static double mx(double a, double b) { return fmax(a, b); }
};

// Row sweep: f(v) for every element, 16-byte loads when the row allows (float32, K % 4 == 0, aligned).
template <typename T, typename F>
__device__ __forceinline__ void for_each_in_row(const T *__restrict__ row, int K, F f)
{
    if (sizeof(T) == 4 && (K & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        const float4 *r4 = reinterpret_cast<const float4 *>(row);
        for (int c = threadIdx.x; c < (K >> 2); c += blockDim.x) {
            const float4 v = r4[c];
            f((T)v.x); f((T)v.y); f((T)v.z); f((T)v.w);
        }
    } else {
        for (int k = threadIdx.x; k < K; k += blockDim.x) f(row[k]);
    }
}

// Row max and NaN flag; every thread of the CTA gets the result.
template <typename T>
__device__ __forceinline__ T row_max(const T *__restrict__ row, int K, T *scratch, int *has_nan)
{
    T m = Acc<T>::ninf();
    int bad = 0;
    for_each_in_row(row, K, [&](T v) {
        bad |= (v != v);
        m = Acc<T>::mx(m, v);
    });
    m = block_allreduce(m, Acc<T>::ninf(), typename Acc<T>::OpMax(), scratch);
    *has_nan = __syncthreads_or(bad);
    return m;
}

template <typename T>
__global__ void logsumexp_rows_kernel(const T *__restrict__ lw, int B, int K, T *__restrict__ lse, int32_t *flags)
{
    __shared__ T scratch[32];
    for (int row = blockIdx.x; row < B; row += gridDim.x) {
        const T *r = lw + (size_t)row * K;
        int bad;
        const T m = row_max(r, K, scratch, &bad);
        T out;
        if (bad) {
            out = m + (T)NAN;
            if (threadIdx.x == 0 && flags) atomicOr(flags, AESMC_FLAG_NAN);
        } else if (!(fabs((double)m) < INFINITY)) {
            out = m; // all -inf -> -inf ; +inf present -> +inf (torch.logsumexp convention)
        } else {
            T s = 0;
            for_each_in_row(r, K, [&](T v) { s += Acc<T>::ex(v - m); });
            s = block_allreduce(s, (T)0, typename Acc<T>::OpSum(), scratch);
            out = m + Acc<T>::lg(s);
        }
        if (threadIdx.x == 0) lse[row] = out;
    }
}

// out = lw - lse (exponentiate=0) or exp(lw - lse) = softmax (math.py:6-51).  The normaliser and the final
// exp / subtraction are evaluated in float64 and rounded once, so every output is within half a float32 ulp
// (plus ~1e-16) of the exact value: the reference's own tests compare against exact float64 values at
// rtol = 1e-7 (test_math.py:119-126), which a float32 lse (its rounding alone moves exp(lw - lse) by up to
// |lse| * 6e-8 relative) only meets by luck.  Off the hot path (the step kernel normalises its own weights).
__global__ void lognormexp_kernel(const float *__restrict__ lw, int B, int K, float *__restrict__ out, int exponentiate)
{
    __shared__ float scratch[32];
    __shared__ double scratch_d[32];
    for (int row = blockIdx.x; row < B; row += gridDim.x) {
        const float *r = lw + (size_t)row * K;
        float *o = out + (size_t)row * K;
        int bad;
        const float m = row_max(r, K, scratch, &bad);
        double lse;
        if (bad) lse = NAN;
        else if (!(fabsf(m) < INFINITY)) lse = m;
        else {
            double s = 0.0;
            for_each_in_row(r, K, [&](float v) { s += exp((double)v - (double)m); });
            s = block_allreduce(s, 0.0, OpSumD(), scratch_d);
            lse = (double)m + log(s);
        }
        auto f = [&](float v) { const double d = (double)v - lse; return (float)(exponentiate ? exp(d) : d); };
        if ((K & 3) == 0 && ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(o)) & 15) == 0) {
            for (int c = threadIdx.x; c < (K >> 2); c += blockDim.x) {
                float4 v = reinterpret_cast<const float4 *>(r)[c];
                v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
                reinterpret_cast<float4 *>(o)[c] = v;
            }
        } else {
            for (int k = threadIdx.x; k < K; k += blockDim.x) o[k] = f(r[k]);
        }
    }
}

template <typename T>
__global__ void log_ess_kernel(const T *__restrict__ lw, int B, int K, T *__restrict__ out)
{
    __shared__ T scratch[32];
    for (int row = blockIdx.x; row < B; row += gridDim.x) {
        const T *r = lw + (size_t)row * K;
        int bad;
        const T m = row_max(r, K, scratch, &bad);
        T s1 = 0, s2 = 0;
        for_each_in_row(r, K, [&](T v) {
            const T e = Acc<T>::ex(v - m);
            s1 += e;
            s2 += e * e;
        });
        s1 = block_allreduce(s1, (T)0, typename Acc<T>::OpSum(), scratch);
        s2 = block_allreduce(s2, (T)0, typename Acc<T>::OpSum(), scratch);
        // 2*(m + log s1) - (2m + log s2)
        if (threadIdx.x == 0) out[row] = bad ? (T)NAN : (T)2 * Acc<T>::lg(s1) - Acc<T>::lg(s2);
    }
}

// g = g_log_w + g_lse[b] * exp(log_w - lse[b]);  g_pos = g, g_neg = -g
__global__ void step_bwd_kernel(const float *__restrict__ lw, const float *__restrict__ lse,
                                const float *__restrict__ glw, const float *__restrict__ glse, int64_t n4,
                                int K, float *__restrict__ gpos, float *__restrict__ gneg)
{
    // vector path: K % 4 == 0 so a float4 never straddles two rows
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t row = (i * 4) / K;
        float4 g = glw ? __ldcs(reinterpret_cast<const float4 *>(glw) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (glse) {
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(lw) + i);
            const float l = __ldg(lse + row), s = __ldg(glse + row);
            g.x = fmaf(s, expf(v.x - l), g.x);
            g.y = fmaf(s, expf(v.y - l), g.y);
            g.z = fmaf(s, expf(v.z - l), g.z);
            g.w = fmaf(s, expf(v.w - l), g.w);
        }
        __stcs(reinterpret_cast<float4 *>(gpos) + i, g);
        if (gneg) __stcs(reinterpret_cast<float4 *>(gneg) + i, make_float4(-g.x, -g.y, -g.z, -g.w));
    }
}
__global__ void step_bwd_scalar_kernel(const float *__restrict__ lw, const float *__restrict__ lse,
                                       const float *__restrict__ glw, const float *__restrict__ glse, int64_t n,
                                       int K, float *__restrict__ gpos, float *__restrict__ gneg)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t row = i / K;
        float g = glw ? glw[i] : 0.f;
        if (glse) g = fmaf(glse[row], expf(lw[i] - lse[row]), g);
        gpos[i] = g;
        if (gneg) gneg[i] = -g;
    }
}

__global__ void is_accumulate_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                     const float *__restrict__ c, float *__restrict__ acc,
                                     float *__restrict__ lw, int64_t n, int first)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = a[i];
        if (b) v = __fadd_rn(v, b[i]);
        if (c) v = __fsub_rn(v, c[i]);
        if (lw) lw[i] = v;
        acc[i] = first ? v : __fadd_rn(acc[i], v);
    }
}

// mean[b,d] = sum_k softmax(lw)[b,k] x[b,k,d]; second[b,d] likewise with x^2.  One CTA per row,
// d processed in chunks of 8 register accumulators.
__global__ void weighted_moments_kernel(const float *__restrict__ x, const float *__restrict__ lw, int B, int K,
                                        int D, float *__restrict__ mean, float *__restrict__ second)
{
    __shared__ float scratch[32];
    for (int row = blockIdx.x; row < B; row += gridDim.x) {
        const float *r = lw + (size_t)row * K;
        const float *xr = x + (size_t)row * K * D;
        int bad;
        const float m = row_max(r, K, scratch, &bad);
        float s = 0.f;
        for_each_in_row(r, K, [&](float v) { s += expf(v - m); });
        s = block_allreduce(s, 0.f, OpSumF(), scratch);
        const float lse = bad ? NAN : m + logf(s);
        if (D == 1 && (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(xr)) & 15) == 0) {
            float a1 = 0.f, a2 = 0.f; // scalar latents: both tables move as 16-byte vectors
            for (int c = threadIdx.x; c < (K >> 2); c += blockDim.x) {
                const float4 l = reinterpret_cast<const float4 *>(r)[c], v = reinterpret_cast<const float4 *>(xr)[c];
                const float w0 = expf(l.x - lse), w1 = expf(l.y - lse), w2 = expf(l.z - lse), w3 = expf(l.w - lse);
                a1 = fmaf(w0, v.x, a1); a2 = fmaf(w0 * v.x, v.x, a2);
                a1 = fmaf(w1, v.y, a1); a2 = fmaf(w1 * v.y, v.y, a2);
                a1 = fmaf(w2, v.z, a1); a2 = fmaf(w2 * v.z, v.z, a2);
                a1 = fmaf(w3, v.w, a1); a2 = fmaf(w3 * v.w, v.w, a2);
            }
            a1 = block_allreduce(a1, 0.f, OpSumF(), scratch);
            a2 = block_allreduce(a2, 0.f, OpSumF(), scratch);
            if (threadIdx.x == 0) {
                mean[row] = a1;
                if (second) second[row] = a2;
            }
            continue;
        }
        for (int d0 = 0; d0 < D; d0 += 8) {
            float a1[8], a2[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { a1[q] = 0.f; a2[q] = 0.f; }
            const int nd = min(8, D - d0);
            for (int k = threadIdx.x; k < K; k += blockDim.x) {
                const float w = expf(r[k] - lse);
                const float *xp = xr + (size_t)k * D + d0;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < nd) { const float v = xp[q]; a1[q] = fmaf(w, v, a1[q]); a2[q] = fmaf(w * v, v, a2[q]); }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float t1 = block_allreduce(a1[q], 0.f, OpSumF(), scratch);
                const float t2 = block_allreduce(a2[q], 0.f, OpSumF(), scratch);
                if (threadIdx.x == 0 && q < nd) {
                    mean[(size_t)row * D + d0 + q] = t1;
                    if (second) second[(size_t)row * D + d0 + q] = t2;
                }
            }
        }
    }
}

// ---- torch.distributions.Normal.log_prob over a [B, K] table in ONE pass ----------------------------------
// state.log_prob (state.py:114-155) evaluates Normal.log_prob through six torch elementwise kernels
// (sub, pow, neg, div, sub, sub); this kernel does the same float32 operations in the same order --
//   -((value - loc) ** 2) / (2 * var) - log_scale - log(sqrt(2 pi))
// so the result is bit-identical.  Operand kinds: 0 = one element per particle [B*K], 1 = one element per
// row [B] (broadcast over particles), 2 = a single element in device memory.  scale is a scalar: either a
// device element (true IEEE division by 2*scale^2, log(scale) with logf, as torch does for a 0-dim CUDA
// tensor) or host-supplied (inv_two_var, log_scale) -- torch multiplies by the float32 reciprocal when
// the divisor is a CPU scalar.
struct NormalArgs {
    const float *value, *loc, *scale_dev;
    int value_kind, loc_kind;
    float loc_host, inv_two_var_host, log_scale_host, half_log_2pi;
    int scale_on_host, loc_on_host;
};

// One row per blockIdx.y sweep (no division per element); VEC: K % 4 == 0 and 16-byte aligned tables, four
// particles per thread.  Per-row and scalar operands are fetched once per row.
struct NormalRow {
    float v_row, mu_row;
    bool v_table, mu_table;
};
__device__ __forceinline__ NormalRow normal_row(const NormalArgs &a, int row)
{
    NormalRow r;
    r.v_table = a.value_kind == 0;
    r.v_row = r.v_table ? 0.f : __ldg(a.value + (a.value_kind == 1 ? row : 0));
    r.mu_table = !a.loc_on_host && a.loc_kind == 0;
    r.mu_row = a.loc_on_host ? a.loc_host : (r.mu_table ? 0.f : __ldg(a.loc + (a.loc_kind == 1 ? row : 0)));
    return r;
}

template <bool VEC>
__global__ void __launch_bounds__(256) normal_log_prob_kernel(const NormalArgs a, int B, int K, float *__restrict__ out)
{
    float two_var = 0.f, log_scale = a.log_scale_host;
    if (!a.scale_on_host) {
        const float sc = __ldg(a.scale_dev);
        two_var = __fmul_rn(2.0f, __fmul_rn(sc, sc));
        log_scale = logf(sc);
    }
    auto lp = [&](float v, float mu) {
        const float d = __fsub_rn(v, mu);
        const float num = -__fmul_rn(d, d);
        const float q = a.scale_on_host ? __fmul_rn(num, a.inv_two_var_host) : __fdiv_rn(num, two_var);
        return __fsub_rn(__fsub_rn(q, log_scale), a.half_log_2pi);
    };
    const int step = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
    for (int row = blockIdx.y; row < B; row += gridDim.y) {
        const NormalRow r = normal_row(a, row);
        const size_t base = (size_t)row * K;
        if (VEC) {
            for (int c = first; c < (K >> 2); c += step) {
                float4 v = make_float4(r.v_row, r.v_row, r.v_row, r.v_row), mu = make_float4(r.mu_row, r.mu_row, r.mu_row, r.mu_row);
                if (r.v_table) v = reinterpret_cast<const float4 *>(a.value + base)[c];
                if (r.mu_table) mu = reinterpret_cast<const float4 *>(a.loc + base)[c];
                reinterpret_cast<float4 *>(out + base)[c] = make_float4(lp(v.x, mu.x), lp(v.y, mu.y), lp(v.z, mu.z), lp(v.w, mu.w));
            }
        } else {
            for (int k = first; k < K; k += step)
                out[base + k] = lp(r.v_table ? a.value[base + k] : r.v_row, r.mu_table ? a.loc[base + k] : r.mu_row);
        }
    }
}

// backward: g_value = -g * d / var, and the per-particle terms of the loc / scale gradients
//   g_loc_term = g * d / var          g_scale_term = g * (d^2 / scale^3 - 1 / scale)
template <bool VEC>
__global__ void __launch_bounds__(256) normal_log_prob_bwd_kernel(const NormalArgs a, const float *__restrict__ g, float scale,
                                                                  int B, int K, float *__restrict__ g_value,
                                                                  float *__restrict__ g_loc, float *__restrict__ g_scale)
{
    const float sc = a.scale_on_host ? scale : __ldg(a.scale_dev);
    const float inv_var = 1.0f / (sc * sc), inv_sc = 1.0f / sc;
    const int step = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
    for (int row = blockIdx.y; row < B; row += gridDim.y) {
        const NormalRow r = normal_row(a, row);
        const size_t base = (size_t)row * K;
        if (VEC) {
            for (int c = first; c < (K >> 2); c += step) {
                float4 v = make_float4(r.v_row, r.v_row, r.v_row, r.v_row), mu = make_float4(r.mu_row, r.mu_row, r.mu_row, r.mu_row);
                if (r.v_table) v = reinterpret_cast<const float4 *>(a.value + base)[c];
                if (r.mu_table) mu = reinterpret_cast<const float4 *>(a.loc + base)[c];
                const float4 gi = reinterpret_cast<const float4 *>(g + base)[c];
                const float d0 = v.x - mu.x, d1 = v.y - mu.y, d2 = v.z - mu.z, d3 = v.w - mu.w;
                const float t0 = gi.x * d0 * inv_var, t1 = gi.y * d1 * inv_var, t2 = gi.z * d2 * inv_var, t3 = gi.w * d3 * inv_var;
                if (g_value) reinterpret_cast<float4 *>(g_value + base)[c] = make_float4(-t0, -t1, -t2, -t3);
                if (g_loc) reinterpret_cast<float4 *>(g_loc + base)[c] = make_float4(t0, t1, t2, t3);
                if (g_scale)
                    reinterpret_cast<float4 *>(g_scale + base)[c] =
                        make_float4(gi.x * (d0 * d0 * inv_var * inv_sc - inv_sc), gi.y * (d1 * d1 * inv_var * inv_sc - inv_sc),
                                    gi.z * (d2 * d2 * inv_var * inv_sc - inv_sc), gi.w * (d3 * d3 * inv_var * inv_sc - inv_sc));
            }
        } else {
            for (int k = first; k < K; k += step) {
                const size_t i = base + k;
                const float d = (r.v_table ? a.value[i] : r.v_row) - (r.mu_table ? a.loc[i] : r.mu_row), gi = g[i];
                const float t = gi * d * inv_var;
                if (g_value) g_value[i] = -t;
                if (g_loc) g_loc[i] = t;
                if (g_scale) g_scale[i] = gi * (d * d * inv_var * inv_sc - inv_sc);
            }
        }
    }
}

static bool aligned16(std::initializer_list<const void *> ptrs)
{
    uintptr_t v = 0;
    for (const void *p : ptrs) v |= reinterpret_cast<uintptr_t>(p);
    return (v & 15) == 0;
}
static dim3 rows_grid2d(int64_t B, int64_t per_row_threads)
{
    int64_t gx = (per_row_threads + 255) / 256;
    const int64_t cap = B >= 1024 ? 4 : 64;
    if (gx > cap) gx = cap;
    return dim3((unsigned)(gx < 1 ? 1 : gx), (unsigned)(B < 65535 ? B : 65535));
}

// Exhaustive device-side check of np_expf_nonpos against np_expf over every float in [-104, -0] and
// -inf: out[0] = number of bit mismatches, out[1] = bit pattern of one mismatching input.
__global__ void selftest_expf_kernel(unsigned long long *out)
{
    const unsigned long long n = 0x42D00000ull + 1ull; // patterns 0x80000000 .. 0xC2D00000, then -inf
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0, where = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        const unsigned bits = (i == n) ? 0xFF800000u : (0x80000000u + (unsigned)i);
        const float x = __uint_as_float(bits);
        const float want = np_expf(x);
        // the packed pair version, with this input in either slot next to an unrelated value
        float p0, p1, q0, q1;
        const float other = __uint_as_float(0x80000000u + (unsigned)((i * 2654435761ull) % n));
        np_expf_nonpos_pair(x, other, p0, p1);
        np_expf_nonpos_pair(other, x, q0, q1);
        const float s = np_expf_nonpos(x);
        if (__float_as_uint(s) != __float_as_uint(want) || __float_as_uint(p0) != __float_as_uint(want) ||
            __float_as_uint(q1) != __float_as_uint(want)) {
            ++bad;
            where = bits;
            out[2] = ((unsigned long long)__float_as_uint(want) << 32) | __float_as_uint(s);
            out[3] = ((unsigned long long)__float_as_uint(p0) << 32) | __float_as_uint(q1);
            out[4] = __float_as_uint(other);
        }
    }
    if (bad) { atomicAdd(out, bad); atomicExch(out + 1, where); }
}

static unsigned flat_grid(int64_t n, int threads);

int normal_log_prob_f32(const float *value, int value_kind, const float *loc, int loc_kind, float loc_host,
                        const float *scale_dev, float inv_two_var_host, float log_scale_host, float half_log_2pi,
                        int64_t B, int64_t K, float *out, cudaStream_t st)
{
    NormalArgs a;
    a.value = value; a.value_kind = value_kind; a.loc = loc; a.loc_kind = loc_kind; a.loc_host = loc_host;
    a.loc_on_host = (loc == nullptr); a.scale_dev = scale_dev; a.scale_on_host = (scale_dev == nullptr);
    a.inv_two_var_host = inv_two_var_host; a.log_scale_host = log_scale_host; a.half_log_2pi = half_log_2pi;
    const bool vec = (K & 3) == 0 && aligned16({value_kind == 0 ? value : nullptr, (loc && loc_kind == 0) ? loc : nullptr, out});
    if (vec) normal_log_prob_kernel<true><<<rows_grid2d(B, K / 4), 256, 0, st>>>(a, (int)B, (int)K, out);
    else normal_log_prob_kernel<false><<<rows_grid2d(B, K), 256, 0, st>>>(a, (int)B, (int)K, out);
    count_launch();
    return check_launch("normal_log_prob_kernel");
}

int normal_log_prob_bwd_f32(const float *value, int value_kind, const float *loc, int loc_kind, float loc_host,
                            const float *scale_dev, float scale_host, const float *g, int64_t B, int64_t K,
                            float *g_value, float *g_loc, float *g_scale, cudaStream_t st)
{
    NormalArgs a;
    a.value = value; a.value_kind = value_kind; a.loc = loc; a.loc_kind = loc_kind; a.loc_host = loc_host;
    a.loc_on_host = (loc == nullptr); a.scale_dev = scale_dev; a.scale_on_host = (scale_dev == nullptr);
    a.inv_two_var_host = 0.f; a.log_scale_host = 0.f; a.half_log_2pi = 0.f;
    const bool vec = (K & 3) == 0 && aligned16({value_kind == 0 ? value : nullptr, (loc && loc_kind == 0) ? loc : nullptr, g, g_value, g_loc, g_scale});
    if (vec) normal_log_prob_bwd_kernel<true><<<rows_grid2d(B, K / 4), 256, 0, st>>>(a, g, scale_host, (int)B, (int)K, g_value, g_loc, g_scale);
    else normal_log_prob_bwd_kernel<false><<<rows_grid2d(B, K), 256, 0, st>>>(a, g, scale_host, (int)B, (int)K, g_value, g_loc, g_scale);
    count_launch();
    return check_launch("normal_log_prob_bwd_kernel");
}

int launch_selftest_expf(unsigned long long *out, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(out, 0, 40, st);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
    selftest_expf_kernel<<<148 * 16, 256, 0, st>>>(out);
    count_launch();
    return check_launch("selftest_expf_kernel");
}

static int row_threads(int64_t K) { return K >= 4096 ? 256 : (K >= 1024 ? 128 : (K >= 128 ? 64 : 32)); }
static unsigned row_grid(int64_t B, int threads)
{
    int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t cap = (int64_t)sms * (2048 / threads);
    return (unsigned)(B < cap ? B : cap);
}
static unsigned flat_grid(int64_t n, int threads)
{
    int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

// row_stats.cu: one warp per row, single pass (used when there are enough rows to fill the machine)
bool row_stats_warp_supported(const void *lw, const void *x, int64_t B, int64_t K);
int launch_logsumexp_warp_f32(const float *, int64_t, int64_t, float *, int32_t *, cudaStream_t);
int launch_log_ess_warp_f32(const float *, int64_t, int64_t, float *, cudaStream_t);
int launch_weighted_moments_warp_f32(const float *, const float *, int64_t, int64_t, float *, float *, cudaStream_t);

int launch_logsumexp_f32(const float *lw, int64_t B, int64_t K, float *lse, int32_t *flags, cudaStream_t st)
{
    if (row_stats_warp_supported(lw, nullptr, B, K)) return launch_logsumexp_warp_f32(lw, B, K, lse, flags, st);
    const int t = row_threads(K);
    logsumexp_rows_kernel<float><<<row_grid(B, t), t, 0, st>>>(lw, (int)B, (int)K, lse, flags);
    count_launch();
    return check_launch("logsumexp_rows_kernel<float>");
}
int launch_logsumexp_f64(const double *lw, int64_t B, int64_t K, double *lse, int32_t *flags, cudaStream_t st)
{
    const int t = row_threads(K);
    logsumexp_rows_kernel<double><<<row_grid(B, t), t, 0, st>>>(lw, (int)B, (int)K, lse, flags);
    count_launch();
    return check_launch("logsumexp_rows_kernel<double>");
}
int launch_lognormexp_f32(const float *lw, int64_t B, int64_t K, float *out, int exponentiate, cudaStream_t st)
{
    const int t = row_threads(K);
    lognormexp_kernel<<<row_grid(B, t), t, 0, st>>>(lw, (int)B, (int)K, out, exponentiate);
    count_launch();
    return check_launch("lognormexp_kernel");
}
int launch_log_ess_f32(const float *lw, int64_t B, int64_t K, float *out, cudaStream_t st)
{
    if (row_stats_warp_supported(lw, nullptr, B, K)) return launch_log_ess_warp_f32(lw, B, K, out, st);
    const int t = row_threads(K);
    log_ess_kernel<float><<<row_grid(B, t), t, 0, st>>>(lw, (int)B, (int)K, out);
    count_launch();
    return check_launch("log_ess_kernel<float>");
}
int launch_log_ess_f64(const double *lw, int64_t B, int64_t K, double *out, cudaStream_t st)
{
    const int t = row_threads(K);
    log_ess_kernel<double><<<row_grid(B, t), t, 0, st>>>(lw, (int)B, (int)K, out);
    count_launch();
    return check_launch("log_ess_kernel<double>");
}
int launch_step_bwd_f32(const float *lw, const float *lse, const float *glw, const float *glse, int64_t B,
                        int64_t K, float *gpos, float *gneg, cudaStream_t st)
{
    const int64_t n = B * K;
    auto aligned = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if ((K & 3) == 0 && aligned(lw) && aligned(glw) && aligned(gpos) && aligned(gneg)) {
        step_bwd_kernel<<<flat_grid(n / 4, 256), 256, 0, st>>>(lw, lse, glw, glse, n / 4, (int)K, gpos, gneg);
    } else {
        step_bwd_scalar_kernel<<<flat_grid(n, 256), 256, 0, st>>>(lw, lse, glw, glse, n, (int)K, gpos, gneg);
    }
    count_launch();
    return check_launch("step_bwd_kernel");
}
int launch_is_accumulate_f32(const float *a, const float *b, const float *c, float *acc, float *lw, int64_t n,
                             int first, cudaStream_t st)
{
    is_accumulate_kernel<<<flat_grid(n, 256), 256, 0, st>>>(a, b, c, acc, lw, n, first);
    count_launch();
    return check_launch("is_accumulate_kernel");
}
int launch_weighted_moments_f32(const float *x, const float *lw, int64_t B, int64_t K, int64_t D, float *mean,
                                float *second, cudaStream_t st)
{
    if (D == 1 && row_stats_warp_supported(lw, x, B, K)) return launch_weighted_moments_warp_f32(x, lw, B, K, mean, second, st);
    const int t = row_threads(K);
    weighted_moments_kernel<<<row_grid(B, t), t, 0, st>>>(x, lw, (int)B, (int)K, (int)D, mean, second);
    count_launch();
    return check_launch("weighted_moments_kernel");
}

} // namespace aesmc
