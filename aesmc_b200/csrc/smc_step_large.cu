// smc_step_large.cu -- the SMC step when a row no longer fits one CTA (K > 16 384; BASELINE config 5,
// K up to 10^6 and beyond).  The row is cut into tiles of 4 096 particles; rows x tiles CTAs run the
// elementwise / scan / search phases, and the only cross-tile quantities (max, sum, running total,
// boundary of the expansion) are associative, so they travel through small per-tile arrays instead of
// a serial chain:
//
//   L1 prep      log_w = (a+b)-c -> HBM, per-tile max, NaN flag                      grid (tiles, B)
//   L2 weights   FAST : e = exp2(lw - max) -> W, per-tile sums                       grid (tiles, B)
//                EXACT: scipy/numpy-order lse, one CTA per row (pairwise tree with    grid (B)
//                       macro-leaves of <= 1024 particles evaluated per thread)
//   L3 cdf       FAST : tile scan + offset from the tile sums                        grid (tiles, B)
//                EXACT: the reference's sequential float32 cumulative sum, streamed   grid (B)
//                       16 384 particles at a time with the exact carry (exact_scan.cuh)
//   L4 search    closed-form offspring boundaries c_j, run starts scattered into      grid (tiles, B)
//                a zeroed [B, K] int32 mark table (atomicMax)
//   L5 expand    one float64 binary search per tile for the ancestor entering the     grid (tiles, B)
//                tile, max-scan of the tile's marks, idx out, ancestral gather
//
// Workspace (caller-allocated, aesmc_smc_step_workspace_bytes): W [B,K] f32, marks [B,K] i32, per-tile
// max / sum [B, tiles], per-row max / total / lse.
#include "common.cuh"
#include "pairwise.cuh"
#include "scan.cuh"
#include "exact_scan.cuh"

namespace aesmc {

constexpr int kTile = 4096;
constexpr int kTileThreads = 256;

struct LargeParams {
    const float *a, *b, *c;
    const double *u;
    int B, K, ntiles;
    float *log_w, *lse;
    int32_t *idx;
    const float *x_in;
    float *x_out;
    int D;
    int32_t *flags;
    float *W;        // [B, K]
    int *marks;      // [B, K]
    float *tmax;     // [B, ntiles]
    float *tsum;     // [B, ntiles]
    float *rowmax;   // [B]
    float *rowtotal; // [B]
    float *rowlse;   // [B]
    int *rowbad;     // [B] 1: NaN, 2: degenerate
    float tol32;
};

// ---- L1 ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads) large_prep_kernel(const LargeParams p)
{
    __shared__ float s_f[32];
    const int row = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const size_t off = (size_t)row * p.K;
    const int k0 = tile * kTile, k1 = min(k0 + kTile, p.K);
    float vmax = -INFINITY;
    int bad = 0;
    for (int k = k0 + tid; k < k1; k += kTileThreads) {
        float v = p.a[off + k];
        if (p.b) v = __fadd_rn(v, p.b[off + k]);
        if (p.c) v = __fsub_rn(v, p.c[off + k]);
        p.log_w[off + k] = v;
        bad |= (v != v);
        vmax = fmaxf(vmax, v);
    }
    vmax = block_allreduce(vmax, -INFINITY, OpMaxF(), s_f);
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        p.tmax[(size_t)row * p.ntiles + tile] = vmax;
        if (bad) { atomicOr(p.flags, AESMC_FLAG_NAN); atomicOr(p.rowbad + row, 1); }
    }
}

// row max from the per-tile maxima (every CTA of the row recomputes it: ntiles <= a few hundred)
__device__ __forceinline__ float row_max_from_tiles(const LargeParams &p, int row, float *s_f)
{
    float m = -INFINITY;
    for (int t = threadIdx.x; t < p.ntiles; t += blockDim.x) m = fmaxf(m, p.tmax[(size_t)row * p.ntiles + t]);
    return block_allreduce(m, -INFINITY, OpMaxF(), s_f);
}

// ---- L2 FAST ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads) large_expsum_kernel(const LargeParams p)
{
    __shared__ float s_f[32];
    const int row = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const size_t off = (size_t)row * p.K;
    const float vmax = row_max_from_tiles(p, row, s_f);
    if (tile == 0 && tid == 0) {
        p.rowmax[row] = vmax;
        if (!(fabsf(vmax) < INFINITY) && !p.rowbad[row]) { atomicOr(p.flags, AESMC_FLAG_DEGENERATE); atomicOr(p.rowbad + row, 2); }
    }
    const int k0 = tile * kTile, k1 = min(k0 + kTile, p.K);
    const float shift = vmax * 1.4426950408889634f;
    float part = 0.f;
    for (int k = k0 + tid; k < k1; k += kTileThreads) {
        const float e = exp2f(fmaf(p.log_w[off + k], 1.4426950408889634f, -shift));
        if (p.idx) p.W[off + k] = e;
        part += e;
    }
    part = block_allreduce(part, 0.f, OpSumF(), s_f);
    if (tid == 0) p.tsum[(size_t)row * p.ntiles + tile] = part;
}

// ---- L3 FAST ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads) large_scan_kernel(const LargeParams p)
{
    __shared__ float s_tile[kTile];
    __shared__ float s_wtot[32], s_pre[32];
    __shared__ double s_d[2];
    const int row = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, nwarp = kTileThreads >> 5;
    const size_t off = (size_t)row * p.K;
    if (tid == 0) { // tile sums are folded in a fixed order, in double: offsets are monotone across tiles
        double before = 0.0, all = 0.0;
        for (int t = 0; t < p.ntiles; ++t) {
            const double v = (double)p.tsum[(size_t)row * p.ntiles + t];
            if (t == tile) before = all;
            all += v;
        }
        s_d[0] = before;
        s_d[1] = all;
        if (tile == 0) {
            const float lse = p.rowbad[row] ? ((p.rowbad[row] & 1) ? __int_as_float(0x7fc00000) : p.rowmax[row])
                                             : p.rowmax[row] + (float)log(all);
            p.rowlse[row] = lse;
            p.rowtotal[row] = (float)all;
            if (p.lse) p.lse[row] = lse;
        }
    }
    if (!p.idx) return;
    const int k0 = tile * kTile, n = min(kTile, p.K - k0);
    for (int k = tid; k < n; k += kTileThreads) s_tile[k] = p.W[off + k0 + k];
    __syncthreads();
    const int seg = ((n + nwarp * 32 - 1) / (nwarp * 32)) * 32;
    segment_scan_inplace(s_tile, n, seg, 0.f, OpSumF(), s_wtot);
    __syncthreads();
    {
        float run = 0.f;
        for (int w = 0; w < nwarp; ++w) { const float t = s_wtot[w]; if (w == (tid >> 5)) s_pre[w] = run; run += t; }
    }
    __syncthreads();
    const float before = (float)s_d[0];
    for (int k = tid; k < n; k += kTileThreads) p.W[off + k0 + k] = before + (s_pre[k / seg] + s_tile[k]);
}

// ---- L2 EXACT: scipy.special.logsumexp in numpy's summation order, one CTA per row ---------------------
// numpy pairwise sum of e_i = (lw_i == vmax) ? 0 : np_exp(lw_i - vmax) over n consecutive particles
__device__ float pairwise_exp_rec(const float *lw, int n, float vmax)
{
    auto e = [&](int i) {
        const float d = __fsub_rn(lw[i], vmax);
        return (d == 0.0f) ? 0.0f : np_expf_nonpos(d);
    };
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, e(i));
        return r;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = e(j);
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], e(i + j));
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, e(i));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(pairwise_exp_rec(lw, n2, vmax), pairwise_exp_rec(lw + n2, n - n2, vmax));
}

constexpr int kMacroLeaf = 1024;

__global__ void __launch_bounds__(1024) large_exact_lse_kernel(const LargeParams p, int max_nodes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PwNode *nodes = reinterpret_cast<PwNode *>(smem_raw);
    __shared__ float s_f[32];
    __shared__ int s_i[32];
    __shared__ int s_lvl[kPairwiseMaxLevels + 1];
    __shared__ int s_nlevels;
    const int row = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const float *lw = p.log_w + (size_t)row * p.K;
    const float vmax = row_max_from_tiles(p, row, s_f);
    if (tid == 0) {
        p.rowmax[row] = vmax;
        if (!(fabsf(vmax) < INFINITY) && !p.rowbad[row]) { atomicOr(p.flags, AESMC_FLAG_DEGENERATE); atomicOr(p.rowbad + row, 2); }
        build_pairwise_tree(nodes, s_lvl, &s_nlevels, p.K, kMacroLeaf);
    }
    __syncthreads();
    if (p.rowbad[row]) {
        if (tid == 0) {
            const float lse = (p.rowbad[row] & 1) ? __int_as_float(0x7fc00000) : vmax;
            p.rowlse[row] = lse;
            if (p.lse) p.lse[row] = lse;
        }
        return;
    }
    int cnt = 0;
    for (int k = tid; k < p.K; k += NT) cnt += (lw[k] == vmax);
    cnt = block_allreduce(cnt, 0, OpSumI(), s_i);
    const int nlevels = s_nlevels, nnodes = s_lvl[nlevels];
    for (int n = tid; n < nnodes; n += NT)
        if (nodes[n].child < 0) nodes[n].val = pairwise_exp_rec(lw + nodes[n].start, nodes[n].len, vmax);
    __syncthreads();
    for (int L = nlevels - 2; L >= 0; --L) {
        for (int n = s_lvl[L] + tid; n < s_lvl[L + 1]; n += NT) {
            const int ch = nodes[n].child;
            if (ch >= 0) nodes[n].val = __fadd_rn(nodes[ch].val, nodes[ch + 1].val);
        }
        __syncthreads();
    }
    if (tid == 0) {
        float s = nodes[0].val;
        const float m = (float)cnt;
        if (s != 0.0f) s = __fdiv_rn(s, m);
        const float lse = __fadd_rn(__fadd_rn(fd_log1pf(s), np_logf(m)), vmax);
        p.rowlse[row] = lse;
        if (p.lse) p.lse[row] = lse;
    }
}

// ---- L3 EXACT: np.cumsum's sequential chain, streamed with the exact carry ----------------------------
__global__ void __launch_bounds__(1024) large_exact_scan_kernel(const LargeParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int row_chunks = NT * 4 + (NT * 4 >> 3);
    float4 *bufW4 = reinterpret_cast<float4 *>(smem_raw);
    int *scratch = reinterpret_cast<int *>(bufW4 + row_chunks);
    float *bufW = reinterpret_cast<float *>(bufW4);
    __shared__ ExactScanShared s_scan;
    __shared__ float s_carry;
    const int row = blockIdx.x;
    if (p.rowbad[row]) return;
    const size_t off = (size_t)row * p.K;
    const float lse = p.rowlse[row];
    const int span = NT * kScanItems;
    float carry = 0.f;
    for (int base = 0; base < p.K; base += span) {
        // normalised weights of this span into the padded buffer (striped, coalesced), zeros past K
        for (int e = tid; e < span; e += NT) {
            const int k = base + e;
            bufW[pad_elem(e)] = (k < p.K) ? np_expf_nonpos(__fsub_rn(p.log_w[off + k], lse)) : 0.0f;
        }
        __syncthreads();
        float w[kScanItems];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 v = bufW4[pad_chunk(4 * tid + i)];
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
        float total;
        if (!exact_cumsum_blocked(w, &total, bufW4, scratch, s_scan, carry)) {
            if (tid == 0) { // plain sequential chain over the span
                float acc = carry;
                for (int e = 0; e < span; ++e) { acc = __fadd_rn(acc, bufW[pad_elem(e)]); bufW[pad_elem(e)] = acc; }
                s_carry = acc;
            }
            __syncthreads();
            total = s_carry;
#pragma unroll
            for (int j = 0; j < kScanItems; ++j) w[j] = bufW[pad_elem(kScanItems * tid + j)];
        }
        carry = total;
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            const int k = base + kScanItems * tid + j;
            if (k < p.K) p.W[off + k] = w[j];
        }
        __syncthreads();
    }
    if (tid == 0) p.rowtotal[row] = carry;
}

// ---- L4: closed-form boundaries and run marks ---------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads) large_search_kernel(const LargeParams p, int exact)
{
    const int row = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    if (p.rowbad[row]) return;
    const size_t off = (size_t)row * p.K;
    const int K = p.K;
    const float total = p.rowtotal[row];
    const double u = p.u[row];
    const float u32 = (float)u, Kf = (float)K;
    const double Kd = (double)K, band = Kd * 8.8817841970012523e-16;
    const bool filtered = K < (1 << 20);
    const int k0 = tile * kTile, k1 = min(k0 + kTile, K);
    for (int j = k0 + tid; j < k1; j += kTileThreads) {
        // boundaries of particle j and of its predecessor (recomputed: avoids a cross-tile exchange)
        int c[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int jj = j - 1 + q;
            if (jj < 0) { c[q] = 0; continue; }
            const float cdf = p.W[off + jj];
            const float cdfn = exact ? __fdiv_rn(cdf, total) : cdf / total;
            c[q] = filtered ? count_positions_below_filtered(cdfn, u, u32, K, Kf, p.tol32)
                            : count_positions_below(cdfn, u, K, Kd, band);
            if (jj == K - 1) c[q] = K;
        }
        if (c[1] > c[0]) atomicMax(p.marks + off + c[0], j);
    }
}

// ---- L5: expansion of the run marks into ancestor indices, gather -------------------------------------
__global__ void __launch_bounds__(kTileThreads) large_expand_kernel(const LargeParams p, int exact)
{
    __shared__ int s_tile[kTile];
    __shared__ int s_wtot[32], s_pre[32];
    __shared__ int s_enter;
    const int row = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, nwarp = kTileThreads >> 5;
    const size_t off = (size_t)row * p.K;
    const int K = p.K, k0 = tile * kTile, n = min(kTile, K - k0);
    if (p.rowbad[row]) { // identity ancestry keeps downstream gathers in range
        for (int k = tid; k < n; k += kTileThreads) p.idx[off + k0 + k] = k0 + k;
        if (p.x_in) {
            const size_t xo = (off + k0) * p.D;
            for (int e = tid; e < n * p.D; e += kTileThreads) p.x_out[xo + e] = p.x_in[xo + e];
        }
        return;
    }
    if (tid == 0) {
        // ancestor of the position just before the tile: #{j : cdf_j / total <= pos}, float64 compare
        int enter = 0;
        if (k0 > 0) {
            const float total = p.rowtotal[row];
            const double pos = __ddiv_rn(__dadd_rn(p.u[row], (double)(k0 - 1)), (double)K);
            int lo = 0, hi = K;
            while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                const float cdf = p.W[off + mid];
                const float cdfn = exact ? __fdiv_rn(cdf, total) : cdf / total;
                if ((double)cdfn <= pos && mid != K - 1) lo = mid + 1; else hi = mid;
            }
            enter = min(lo, K - 1);
        }
        s_enter = enter;
    }
    for (int k = tid; k < n; k += kTileThreads) s_tile[k] = p.marks[off + k0 + k];
    __syncthreads();
    const int seg = ((n + nwarp * 32 - 1) / (nwarp * 32)) * 32;
    segment_scan_inplace(s_tile, n, seg, 0, OpMaxI(), s_wtot);
    __syncthreads();
    {
        int run = s_enter;
        for (int w = 0; w < nwarp; ++w) { const int t = s_wtot[w]; if (w == (tid >> 5)) s_pre[w] = run; run = max(run, t); }
    }
    __syncthreads();
    for (int k = tid; k < n; k += kTileThreads) {
        const int id = max(s_tile[k], s_pre[k / seg]);
        s_tile[k] = id;
        p.idx[off + k0 + k] = id;
    }
    if (p.x_in) {
        __syncthreads();
        const int D = p.D;
        const float *xin = p.x_in + off * D;
        float *xout = p.x_out + (off + k0) * D;
        for (int e = tid; e < n * D; e += kTileThreads) {
            const int k = e / D;
            xout[e] = __ldg(xin + (size_t)s_tile[k] * D + (e - k * D));
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------
static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

int64_t smc_step_large_workspace_bytes(int64_t B, int64_t K)
{
    const int64_t nt = (K + kTile - 1) / kTile;
    size_t bytes = 0;
    bytes += align_up((size_t)B * K * 4) * 2;      // W, marks
    bytes += align_up((size_t)B * nt * 4) * 2;     // tmax, tsum
    bytes += align_up((size_t)B * 4) * 4;          // rowmax, rowtotal, rowlse, rowbad
    return (int64_t)bytes;
}

int launch_smc_step_large(const float *a, const float *b, const float *c, const double *u, int64_t B, int64_t K,
                          float *log_w, float *lse, int32_t *idx, const float *x_in, float *x_out, int64_t D,
                          int32_t *flags, int mode, void *workspace, int64_t workspace_bytes, cudaStream_t stream)
{
    const bool exact = (mode == AESMC_MODE_EXACT);
    if (workspace == nullptr || workspace_bytes < smc_step_large_workspace_bytes(B, K)) {
        set_error("aesmc_smc_step_ws_f32: K=%lld needs a workspace of %lld bytes (aesmc_smc_step_workspace_bytes)",
                  (long long)K, (long long)smc_step_large_workspace_bytes(B, K));
        return AESMC_ERR_BAD_ARG;
    }
    if (B > 65535) { set_error("aesmc_smc_step_ws_f32: multi-CTA path supports B <= 65535 rows"); return AESMC_ERR_UNSUPPORTED; }
    LargeParams p;
    p.a = a; p.b = b; p.c = c; p.u = u; p.B = (int)B; p.K = (int)K; p.ntiles = (int)((K + kTile - 1) / kTile);
    p.log_w = log_w; p.lse = lse; p.idx = idx; p.x_in = x_in; p.x_out = x_out; p.D = (int)D; p.flags = flags;
    p.tol32 = (float)K * 1.1920928955078125e-07f + 5.9604644775390625e-08f;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    p.W = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * K * 4);
    p.marks = reinterpret_cast<int *>(ws); ws += align_up((size_t)B * K * 4);
    p.tmax = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * p.ntiles * 4);
    p.tsum = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * p.ntiles * 4);
    p.rowmax = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * 4);
    p.rowtotal = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * 4);
    p.rowlse = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * 4);
    p.rowbad = reinterpret_cast<int *>(ws);
    cudaError_t e = cudaMemsetAsync(p.rowbad, 0, (size_t)B * 4, stream);
    if (e == cudaSuccess && idx) e = cudaMemsetAsync(p.marks, 0, (size_t)B * K * 4, stream);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
    const dim3 grid((unsigned)p.ntiles, (unsigned)B);
    large_prep_kernel<<<grid, kTileThreads, 0, stream>>>(p);
    count_launch();
    if (exact && idx) {
        const int max_nodes = pairwise_max_nodes((int)K); // generous: macro-leaves are 8x larger than leaves
        const size_t smem_lse = (size_t)(2 * (K / 448 + 2)) * sizeof(PwNode); // macro-leaves hold >= 505 particles
        e = cudaFuncSetAttribute(large_exact_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lse);
        if (e != cudaSuccess) { set_error("K=%lld too large for the exact multi-CTA path: %s", (long long)K, cudaGetErrorString(e)); return AESMC_ERR_UNSUPPORTED; }
        large_exact_lse_kernel<<<(unsigned)B, 1024, smem_lse, stream>>>(p, max_nodes);
        count_launch();
        const int nt = 1024;
        const size_t row_chunks = (size_t)nt * 4 + ((size_t)nt * 4 >> 3);
        const size_t smem_scan = row_chunks * 16 + (size_t)(8 * nt + 8) * 4;
        e = cudaFuncSetAttribute(large_exact_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
        large_exact_scan_kernel<<<(unsigned)B, nt, smem_scan, stream>>>(p);
        count_launch();
    } else {
        large_expsum_kernel<<<grid, kTileThreads, 0, stream>>>(p);
        count_launch();
        large_scan_kernel<<<grid, kTileThreads, 0, stream>>>(p);
        count_launch();
    }
    if (idx) {
        large_search_kernel<<<grid, kTileThreads, 0, stream>>>(p, exact ? 1 : 0);
        count_launch();
        large_expand_kernel<<<grid, kTileThreads, 0, stream>>>(p, exact ? 1 : 0);
        count_launch();
    }
    return check_launch("smc_step_large");
}

} // namespace aesmc
