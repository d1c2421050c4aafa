// smc_step.cu -- the fused per-time-step SMC kernel (one CTA per row, row resident in shared memory).
//
// For every row b of K particles, one pass does what the reference spreads over
//   inference.py:97-98,125-126   log_w = (a + b) - c
//   inference.py:130             lse[b] = logsumexp_k log_w           (per-step log-evidence term)
//   inference.py:234-269         sample_ancestral_index (+ math.py:6-51 exponentiate_and_normalize)
//   state.py:158-183             resample of the newest latent
// Phases per row (all intermediate data stays in shared memory / registers):
//   P1  load a,b,c (coalesced, float4 when aligned) -> log_w to HBM + smem, row max, NaN flag
//   P2  exponentiate against the max, row sum -> lse
//   P3  normalised weights -> cumulative distribution (sequential float32 chain in EXACT mode,
//       shuffle block scan in FAST mode)
//   P4  per particle j: c_j = #{k : (u+k)/K < cdf_j/cdf_{K-1}} in closed form (float64), run starts
//       marked in smem, max-scan expands them to ancestor indices  (replaces the K binary searches
//       of np.digitize; the merged sequence of positions and CDF entries is never materialised)
//   P5  coalesced store of idx, fused ancestral gather of the latent
#include <cstdlib>
#include "common.cuh"
#include "pairwise.cuh"
#include "scan.cuh"

namespace aesmc {

struct StepParams {
    const float *a, *b, *c;
    const double *u;
    int B, K;
    float *log_w, *lse;
    int32_t *idx;
    const float *x_in;
    float *x_out;
    int D;
    int32_t *flags;
    int Kp;        // K rounded up to a multiple of 4
    int max_nodes; // capacity of the pairwise tree (EXACT)
    int vec;       // 1: K % 4 == 0 and all row pointers 16-byte aligned -> float4 path
    int stage;     // 0: `a` holds log-probs (full step); 1: `a` holds normalised weights (enter at
                   // the cumulative sum); 2: `a` holds a normalised CDF (enter at the search)
};

constexpr int kMaxLevels = kPairwiseMaxLevels;

template <bool EXACT>
__global__ void __launch_bounds__(256) smc_step_kernel(const StepParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *bufA = reinterpret_cast<float *>(smem_raw); // log_w, later c_j (int)
    float *bufB = bufA + p.Kp;                         // exp / weights / cdf, later run marks (int)
    PwNode *nodes = reinterpret_cast<PwNode *>(bufB + p.Kp);
    __shared__ int s_lvl[kMaxLevels + 1];
    __shared__ int s_nlevels;
    __shared__ float s_redf[32];
    __shared__ int s_redi[32];
    __shared__ float s_wtot_f[32];
    __shared__ int s_wtot_i[32];

    const int tid = threadIdx.x, NT = blockDim.x, nwarp = NT >> 5;
    const int K = p.K;
    const bool vec = p.vec != 0;
    const bool resample = (p.idx != nullptr);
    const int seg = ((K + nwarp * 32 - 1) / (nwarp * 32)) * 32;
    const double Kd = (double)K;
    const double band = Kd * 8.8817841970012523e-16; // K * 2^-50

    if (EXACT && resample && p.stage == 0) {
        if (tid == 0) build_pairwise_tree(nodes, s_lvl, &s_nlevels, K);
        __syncthreads();
    }

    for (int row = blockIdx.x; row < p.B; row += gridDim.x) {
        const size_t off = (size_t)row * K;
        const float *__restrict__ ga = p.a + off;
        const float *__restrict__ gb = p.b ? p.b + off : nullptr;
        const float *__restrict__ gc = p.c ? p.c + off : nullptr;
        float *__restrict__ glw = p.log_w ? p.log_w + off : nullptr;

        if (p.stage != 0) { // parity-staging entry points (tests): weights or CDF injected in `a`
            for (int k = tid; k < K; k += NT) bufB[k] = ga[k];
            __syncthreads();
            float total_s = 1.0f;
            if (p.stage == 1) {
                if (EXACT) {
                    if (tid == 0) {
                        float acc = bufB[0];
                        for (int k = 1; k < K; ++k) { acc = __fadd_rn(acc, bufB[k]); bufB[k] = acc; }
                    }
                    __syncthreads();
                    total_s = bufB[K - 1];
                } else {
                    segment_scan_inplace(bufB, K, seg, 0.f, OpSumF(), s_wtot_f);
                    __syncthreads();
                    float run = 0.f;
                    for (int w = 0; w < nwarp; ++w) { const float t = s_wtot_f[w]; if (w == (tid >> 5)) s_redf[w] = run; run += t; }
                    total_s = run;
                    __syncthreads();
                }
            }
            const double us = p.u[row];
            int *cjs = reinterpret_cast<int *>(bufA);
            for (int j = tid; j < K; j += NT) {
                float cdf = bufB[j];
                if (p.stage == 1 && !EXACT) cdf += s_redf[j / seg];
                const float cdfn = (p.stage == 1) ? __fdiv_rn(cdf, total_s) : cdf;
                int c = count_positions_below(cdfn, us, K, Kd, band);
                if (j == K - 1) c = K;
                cjs[j] = c;
            }
            __syncthreads();
            int *mk = reinterpret_cast<int *>(bufB);
            for (int k = tid; k < K; k += NT) mk[k] = 0;
            __syncthreads();
            for (int j = tid; j < K; j += NT) {
                const int c = cjs[j], cp = j ? cjs[j - 1] : 0;
                if (c > cp) atomicMax(&mk[cp], j);
            }
            __syncthreads();
            segment_scan_inplace(mk, K, seg, 0, OpMaxI(), s_wtot_i);
            __syncthreads();
            {
                int run = 0;
                for (int w = 0; w < nwarp; ++w) { const int t = s_wtot_i[w]; if (w == (tid >> 5)) s_redi[w] = run; run = max(run, t); }
            }
            __syncthreads();
            for (int k = tid; k < K; k += NT) p.idx[off + k] = max(mk[k], s_redi[k / seg]);
            __syncthreads();
            continue;
        }

        // ---- P1: log-weights, row max, NaN detection --------------------------------------
        float vmax = -INFINITY;
        int bad = 0;
        if (vec) {
            const float4 *a4 = reinterpret_cast<const float4 *>(ga);
            const float4 *b4 = reinterpret_cast<const float4 *>(gb);
            const float4 *c4 = reinterpret_cast<const float4 *>(gc);
            float4 *o4 = reinterpret_cast<float4 *>(glw);
            float4 *s4 = reinterpret_cast<float4 *>(bufA);
            const int n4 = K >> 2;
#pragma unroll 4
            for (int i = tid; i < n4; i += NT) {
                float4 v = __ldcs(a4 + i);
                if (gb) { const float4 t = __ldcs(b4 + i); v.x = __fadd_rn(v.x, t.x); v.y = __fadd_rn(v.y, t.y); v.z = __fadd_rn(v.z, t.z); v.w = __fadd_rn(v.w, t.w); }
                if (gc) { const float4 t = __ldcs(c4 + i); v.x = __fsub_rn(v.x, t.x); v.y = __fsub_rn(v.y, t.y); v.z = __fsub_rn(v.z, t.z); v.w = __fsub_rn(v.w, t.w); }
                __stcs(o4 + i, v);
                s4[i] = v;
                bad |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
                vmax = fmaxf(fmaxf(vmax, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
            }
        } else {
            for (int k = tid; k < K; k += NT) {
                float v = ga[k];
                if (gb) v = __fadd_rn(v, gb[k]);
                if (gc) v = __fsub_rn(v, gc[k]);
                glw[k] = v;
                bufA[k] = v;
                bad |= (v != v);
                vmax = fmaxf(vmax, v);
            }
        }
        vmax = block_allreduce(vmax, -INFINITY, OpMaxF(), s_redf);
        bad = __syncthreads_or(bad);
        // row state: 0 ok, else flagged (NaN, or max not finite => normaliser not finite/positive)
        const bool degenerate = bad || !(fabsf(vmax) < INFINITY);
        if (tid == 0 && degenerate)
            atomicOr(p.flags, bad ? AESMC_FLAG_NAN : AESMC_FLAG_DEGENERATE);

        float lse;
        if (degenerate) {
            lse = bad ? __int_as_float(0x7fc00000) : vmax;
            if (tid == 0 && p.lse) p.lse[row] = lse;
            if (resample) { // identity ancestry keeps downstream gathers in range
                for (int k = tid; k < K; k += NT) p.idx[off + k] = k;
                if (p.x_in) {
                    const size_t xo = off * p.D;
                    for (int e = tid; e < K * p.D; e += NT) p.x_out[xo + e] = p.x_in[xo + e];
                }
            }
            __syncthreads();
            continue;
        }

        // ---- P2/P3: lse and cumulative distribution -----------------------------------------
        float total;
        if (EXACT && resample) {
            // scipy.special.logsumexp: the maxima are counted in m and excluded from the sum
            int cnt = 0;
            for (int k = tid; k < K; k += NT) {
                const float v = bufA[k];
                const bool is_max = (v == vmax);
                cnt += is_max;
                bufB[k] = is_max ? 0.0f : np_expf(__fsub_rn(v, vmax));
            }
            cnt = block_allreduce(cnt, 0, OpSumI(), s_redi); // contains __syncthreads
            float s = pairwise_tree_sum<false>(bufB, nodes, s_lvl, s_nlevels);
            const float m = (float)cnt;
            if (s != 0.0f) s = __fdiv_rn(s, m);
            lse = __fadd_rn(__fadd_rn(fd_log1pf(s), np_logf(m)), vmax);
            // normalised weights exp(lw - lse) (math.py:49)
            for (int k = tid; k < K; k += NT) bufB[k] = np_expf(__fsub_rn(bufA[k], lse));
            __syncthreads();
            // np.cumsum: strictly sequential float32 chain (inference.py:257)
            if (tid == 0) {
                float acc = bufB[0];
                int k = 1;
                for (; k < K && (k & 3); ++k) { acc = __fadd_rn(acc, bufB[k]); bufB[k] = acc; }
                float4 *w4 = reinterpret_cast<float4 *>(bufB);
#pragma unroll 4
                for (; k + 3 < K; k += 4) {
                    float4 v = w4[k >> 2];
                    v.x = acc = __fadd_rn(acc, v.x);
                    v.y = acc = __fadd_rn(acc, v.y);
                    v.z = acc = __fadd_rn(acc, v.z);
                    v.w = acc = __fadd_rn(acc, v.w);
                    w4[k >> 2] = v;
                }
                for (; k < K; ++k) { acc = __fadd_rn(acc, bufB[k]); bufB[k] = acc; }
            }
            __syncthreads();
            total = bufB[K - 1];
        } else {
            float part = 0.f;
            for (int k = tid; k < K; k += NT) {
                const float e = __expf(bufA[k] - vmax);
                bufB[k] = e;
                part += e;
            }
            const float ssum = block_allreduce(part, 0.f, OpSumF(), s_redf);
            lse = vmax + logf(ssum);
            total = 0.f;
            if (resample) {
                segment_scan_inplace(bufB, K, seg, 0.f, OpSumF(), s_wtot_f);
                __syncthreads();
                // fold preceding segment totals (same order in every thread)
                float run = 0.f;
                for (int w = 0; w < nwarp; ++w) { const float t = s_wtot_f[w]; if (w == (tid >> 5)) s_redf[w] = run; run += t; }
                total = run;
                __syncthreads();
            }
        }
        if (tid == 0 && p.lse) p.lse[row] = lse;
        if (!resample) { __syncthreads(); continue; }

        // ---- P4: closed-form offspring boundaries c_j, run marks, max-scan --------------------
        const double u = p.u[row];
        int *cj = reinterpret_cast<int *>(bufA);
        int *marks = reinterpret_cast<int *>(bufB);
        for (int j = tid; j < K; j += NT) {
            float cdf = bufB[j];
            if (!EXACT) cdf += s_redf[j / seg];
            const float cdfn = __fdiv_rn(cdf, total); // inference.py:260-261
            int c = count_positions_below(cdfn, u, K, Kd, band);
            if (j == K - 1) c = K; // positions that round to >= 1.0 stay in range
            cj[j] = c;
        }
        __syncthreads();
        for (int k = tid; k < K; k += NT) marks[k] = 0;
        __syncthreads();
        for (int j = tid; j < K; j += NT) {
            const int c = cj[j], cp = j ? cj[j - 1] : 0;
            if (c > cp) atomicMax(&marks[cp], j);
        }
        __syncthreads();
        segment_scan_inplace(marks, K, seg, 0, OpMaxI(), s_wtot_i);
        __syncthreads();
        {
            int run = 0;
            for (int w = 0; w < nwarp; ++w) { const int t = s_wtot_i[w]; if (w == (tid >> 5)) s_redi[w] = run; run = max(run, t); }
        }
        __syncthreads();

        // ---- P5: indices out, fused ancestral gather -----------------------------------------
        int32_t *__restrict__ gidx = p.idx + off;
        if (p.x_in == nullptr || p.D == 1) {
            const float *__restrict__ xin = p.x_in ? p.x_in + off : nullptr;
            float *__restrict__ xout = p.x_out ? p.x_out + off : nullptr;
            for (int k = tid; k < K; k += NT) {
                const int id = max(marks[k], s_redi[k / seg]);
                gidx[k] = id;
                if (xin) xout[k] = __ldg(xin + id);
            }
        } else {
            for (int k = tid; k < K; k += NT) {
                const int id = max(marks[k], s_redi[k / seg]);
                gidx[k] = id;
                cj[k] = id;
            }
            __syncthreads();
            const int D = p.D;
            const size_t xo = off * D;
            const float *__restrict__ xin = p.x_in + xo;
            float *__restrict__ xout = p.x_out + xo;
            const int n = K * D;
            for (int e = tid; e < n; e += NT) {
                const int k = e / D;
                xout[e] = __ldg(xin + (size_t)cj[k] * D + (e - k * D));
            }
        }
        __syncthreads();
    }
}

int64_t smc_step_large_workspace_bytes(int64_t B, int64_t K);
int launch_smc_step_large(const float *, const float *, const float *, const double *, int64_t, int64_t, float *, float *,
                          int32_t *, const float *, float *, int64_t, int32_t *, int, void *, int64_t, cudaStream_t);
bool smc_step_reg_supported(int64_t K, bool vec);
bool smc_step_x_supported(int64_t K, int mode, const void *idx, const void *x_in, int64_t D);
int launch_smc_step_x(const float *a, const float *b, const float *c, const double *u, int64_t B, int64_t K,
                      float *log_w, float *lse, int32_t *idx, const float *x_in, float *x_out, int32_t *flags,
                      cudaStream_t stream);
int launch_smc_step_reg(const float *, const float *, const float *, const double *, int64_t, int64_t, float *, float *,
                        int32_t *, const float *, float *, int64_t, int32_t *, int, cudaStream_t);

static int g_sm_count = 0;
static int sm_count()
{
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

constexpr int kSmemBudget = 227 * 1024 - 2048; // leave room for the static __shared__ arrays

static size_t step_smem_bytes(int K, bool exact)
{
    const int Kp = (K + 3) & ~3;
    size_t bytes = (size_t)Kp * 8;
    if (exact) bytes += (size_t)pairwise_max_nodes(K) * sizeof(PwNode);
    return bytes;
}

int64_t step_workspace_bytes(int64_t B, int64_t K)
{
    if (K <= 8192 || smc_step_reg_supported(K, (K & 3) == 0)) return 0;
    return smc_step_large_workspace_bytes(B, K);
}

int64_t max_particles_single_cta()
{
    int64_t K = 1024;
    while (step_smem_bytes((int)(K + 1024), true) <= (size_t)kSmemBudget) K += 1024;
    return K;
}

int launch_smc_step(const float *a, const float *b, const float *c, const double *u, int64_t B, int64_t K,
                    float *log_w, float *lse, int32_t *idx, const float *x_in, float *x_out, int64_t D,
                    int32_t *flags, int mode, int stage, void *workspace, int64_t workspace_bytes, cudaStream_t stream)
{
    const bool exact = (mode == AESMC_MODE_EXACT);
    StepParams p;
    p.a = a; p.b = b; p.c = c; p.u = u; p.B = (int)B; p.K = (int)K; p.log_w = log_w; p.lse = lse;
    p.idx = idx; p.x_in = x_in; p.x_out = x_out; p.D = (int)D; p.flags = flags;
    p.Kp = ((int)K + 3) & ~3;
    p.max_nodes = pairwise_max_nodes((int)K);
    p.stage = stage;
    {
        const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                               reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(log_w);
        p.vec = ((K & 3) == 0) && ((bits & 15) == 0);
    }
    const bool aligned16 = p.vec && ((reinterpret_cast<uintptr_t>(x_in) | reinterpret_cast<uintptr_t>(x_out) |
                                      reinterpret_cast<uintptr_t>(idx)) & 15) == 0;
    if (stage == 0 && aligned16 && smc_step_x_supported(K, mode, idx, x_in, D))
        return launch_smc_step_x(a, b, c, u, B, K, log_w, lse, idx, x_in, x_out, flags, stream);
    if (stage == 0 && smc_step_reg_supported(K, p.vec != 0) && aligned16)
        return launch_smc_step_reg(a, b, c, u, B, K, log_w, lse, idx, x_in, x_out, D, flags, mode, stream);
    const size_t smem = step_smem_bytes((int)K, exact);
    // long rows the register-blocked kernel did not take: the multi-CTA path is 3-4x faster than the
    // shared-memory kernel below (B = 1024, K = 20 000: 379 vs 1 610 us exact, 208 vs 482 us fast)
    if ((smem > (size_t)kSmemBudget || (workspace != nullptr && K > 8192)) && stage == 0)
        return launch_smc_step_large(a, b, c, u, B, K, log_w, lse, idx, x_in, x_out, D, flags, mode, workspace,
                                     workspace_bytes, stream);
    if (smem > (size_t)kSmemBudget) {
        set_error("aesmc_smc_step_f32: K=%lld exceeds the single-CTA shared-memory path (max %lld)",
                  (long long)K, (long long)max_particles_single_cta());
        return AESMC_ERR_UNSUPPORTED;
    }
    const int threads = K >= 2048 ? 256 : (K >= 512 ? 128 : 64);
    auto kern = exact ? smc_step_kernel<true> : smc_step_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > B) grid = B;
    kern<<<(unsigned)grid, threads, smem, stream>>>(p);
    count_launch();
    return check_launch("smc_step_kernel");
}

} // namespace aesmc
