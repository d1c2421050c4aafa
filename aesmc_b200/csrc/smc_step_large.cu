// smc_step_large.cu -- the SMC step when a row no longer fits one CTA (K > 16 384; BASELINE config 5,
// K up to 10^6 and beyond).  The row is cut into tiles of 4 096 particles; rows x tiles CTAs run the
// elementwise / scan / search phases, and the only cross-tile quantities (max, sum, running total,
// boundary of the expansion) are associative, so they travel through small per-tile arrays instead of
// a serial chain:
//
//   L1 prep      log_w = (a+b)-c -> HBM, per-tile max, NaN flag                      grid (tiles, B)
//                FAST : also the tile-local running sums of exp2(lw - tile max) -> W
//   L2 weights   FAST : per-tile scale and offset, row total, lse                    grid (B)
//                EXACT: scipy/numpy-order lse.  numpy's pairwise-summation tree depends on K only and
//                       is never stored (heap indices, extents recomputed from the root): its leaves
//                       (64..128 particles, 8 lanes each) are evaluated by a (chunks, B) grid, one CTA
//                       per row folds the levels.
//   L3 cdf       FAST : none -- consumers evaluate before_t + W_j * scale_t on the fly
//                EXACT: the reference's sequential float32 cumulative sum.  The row is cut into spans
//                       of 16 384 particles owned by different CTAs (dynamic tickets, span-major): each
//                       classifies and composes its blocks against an estimate of its entry value while
//                       its predecessor is still running, then waits for the exact carry, walks its
//                       ~20 segments, publishes its exit value and only then replays (exact_scan.cuh).
//   L4 bounds    closed-form offspring boundary at the end of every input tile         grid (tiles/256, B)
//   L5 resample  per OUTPUT tile: the input tiles that can own its positions (binary   grid (tiles, B)
//                search in the L4 table), their boundaries recomputed from the CDF table,
//                run starts in a shared-memory tile, max-scan, ancestors, ancestral gather
//
// Workspace (caller-allocated, aesmc_smc_step_workspace_bytes): W [B,K] f32, per-tile max / sum / offset /
// scale / entry [B, tiles], per-row max / total / lse, the heap of summation-tree node values of each row,
// one 8-byte carry slot per span.
#include "common.cuh"
#include "pairwise.cuh"
#include "row_gather.cuh"
#include "scan.cuh"
#include "exact_scan.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace aesmc {

constexpr int kTile = 4096;
constexpr int kTileThreads = 256;
static_assert(kTile == 4096 && kTileThreads == 256, "tile-entry bookkeeping (c >> 12) and the 16-per-thread layouts assume 4096 / 256");

struct LargeParams {
    const float *a, *b, *c;
    const double *u;
    int B, K, ntiles;
    float *log_w, *lse;
    int32_t *idx;
    const float *x_in;
    float *x_out;
    int D;
    RowGather gather;
    int32_t *flags;
    float *W;        // [B, K]
    int *marks;      // [B, K]: the idx output doubles as the table of run marks
    float *tmax;     // [B, ntiles]
    float *tsum;     // [B, ntiles]
    float *rowmax;   // [B]
    float *rowtotal; // [B]
    float *rowlse;   // [B]
    int *rowbad;     // [B] 1: NaN, 2: degenerate
    float tol32;
    // fast mode: W holds tile-local sums, cdf_j = before[t] + W_j * scale[t]
    double *tbefore; // [B, ntiles + 1]
    double *tscale;  // [B, ntiles]
    int tiled;
    int *tenter;     // [B, ntiles] cend: boundary count of the last particle of each input tile (large_bounds_kernel)
    // exact mode
    float *vals;     // [B, heap_size] node values of numpy's pairwise tree, heap-indexed
    int *rowcnt;     // [B] number of particles equal to the row maximum
    unsigned long long *slots; // [B, nspans] (ready << 32 | carry bits) of each span of the chained scan
    unsigned long long *lsums; // [B, nspans] (ready << 32 | bits of the span's own sum of weights)
    int *ticket;     // [0] span tickets, [1] padding; then rowcnt
    int *stats;      // [4] redo after a missed estimate / after a failed replay check / sequential spans
    int nspans, heap_depth, heap_size; // heap_depth: depth of the deepest leaf; heap_size = 2 << heap_depth
};

// ---- L1 ---------------------------------------------------------------------------------------------
// FAST additionally leaves in W the tile-local inclusive sums L_j of e_j = exp2((lw_j - m_t) log2 e), m_t
// the TILE maximum: the row-level quantities then follow from the per-tile (m_t, L_last) pairs alone
// (large_rows_fast_kernel) and the CDF entry of particle j is before_t + L_j * scale_t, evaluated on the
// fly by the search and expansion kernels -- no second and third pass over the row.
template <bool FAST>
__global__ void __launch_bounds__(kTileThreads) large_prep_kernel(const LargeParams p)
{
    __shared__ float s_f[32];
    __shared__ __align__(16) float s_e[FAST ? kTile + kTile / 8 : 4];
    constexpr int kPer = kTile / kTileThreads;
    const int row = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t off = (size_t)row * p.K;
    const int K = p.K, k0 = tile * kTile;
    const bool vec = (K & 3) == 0;
    float lw[kPer]; // vec: float4 chunks tid + 256 i; else elements tid + 256 i
    float vmax = -INFINITY;
    int bad = 0;
    if (vec) {
#pragma unroll
        for (int i = 0; i < kPer / 4; ++i) {
            const int k = k0 + 4 * (tid + kTileThreads * i);
            float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (k < K) {
                v = __ldg(reinterpret_cast<const float4 *>(p.a + off + k));
                if (p.b) {
                    const float4 t = __ldg(reinterpret_cast<const float4 *>(p.b + off + k));
                    v.x = __fadd_rn(v.x, t.x); v.y = __fadd_rn(v.y, t.y); v.z = __fadd_rn(v.z, t.z); v.w = __fadd_rn(v.w, t.w);
                }
                if (p.c) {
                    const float4 t = __ldg(reinterpret_cast<const float4 *>(p.c + off + k));
                    v.x = __fsub_rn(v.x, t.x); v.y = __fsub_rn(v.y, t.y); v.z = __fsub_rn(v.z, t.z); v.w = __fsub_rn(v.w, t.w);
                }
                *reinterpret_cast<float4 *>(p.log_w + off + k) = v;
            }
            lw[4 * i] = v.x; lw[4 * i + 1] = v.y; lw[4 * i + 2] = v.z; lw[4 * i + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int k = k0 + tid + kTileThreads * i;
            float v = -INFINITY;
            if (k < K) {
                v = p.a[off + k];
                if (p.b) v = __fadd_rn(v, p.b[off + k]);
                if (p.c) v = __fsub_rn(v, p.c[off + k]);
                p.log_w[off + k] = v;
            }
            lw[i] = v;
        }
    }
#pragma unroll
    for (int i = 0; i < kPer; ++i) { bad |= (lw[i] != lw[i]); vmax = fmaxf(vmax, lw[i]); }
    vmax = block_allreduce(vmax, -INFINITY, OpMaxF(), s_f);
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        p.tmax[(size_t)row * p.ntiles + tile] = vmax;
        if (bad) { atomicOr(p.flags, AESMC_FLAG_NAN); atomicOr(p.rowbad + row, 1); }
    }
    if (!FAST) return;
    // e_j in the padded tile buffer, then 16 consecutive particles per thread
    const float shift = (fabsf(vmax) < INFINITY ? vmax : 0.f) * 1.4426950408889634f;
    if (vec) {
#pragma unroll
        for (int i = 0; i < kPer / 4; ++i) {
            float4 e;
            e.x = exp2f(fmaf(lw[4 * i], 1.4426950408889634f, -shift)); e.y = exp2f(fmaf(lw[4 * i + 1], 1.4426950408889634f, -shift));
            e.z = exp2f(fmaf(lw[4 * i + 2], 1.4426950408889634f, -shift)); e.w = exp2f(fmaf(lw[4 * i + 3], 1.4426950408889634f, -shift));
            reinterpret_cast<float4 *>(s_e)[pad_chunk(tid + kTileThreads * i)] = e;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kPer; ++i) s_e[pad_elem(tid + kTileThreads * i)] = exp2f(fmaf(lw[i], 1.4426950408889634f, -shift));
    }
    __syncthreads();
    float w[kPer];
#pragma unroll
    for (int i = 0; i < kPer / 4; ++i) {
        const float4 v = reinterpret_cast<const float4 *>(s_e)[pad_chunk(4 * tid + i)];
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 1; i < kPer; ++i) w[i] += w[i - 1];
    float incl = w[kPer - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_f[warp] = incl;
    __syncthreads();
    float pre = incl - w[kPer - 1];
    for (int v = 0; v < warp; ++v) pre += s_f[v];
#pragma unroll
    for (int i = 0; i < kPer; ++i) w[i] += pre;
    if (tid == kTileThreads - 1) p.tsum[(size_t)row * p.ntiles + tile] = w[kPer - 1]; // past K: e = 0, L stays flat
    if (!p.idx) return;
    const int kb = k0 + kPer * tid;
    if (vec) {
#pragma unroll
        for (int i = 0; i < kPer / 4; ++i)
            if (kb + 4 * i < K)
                reinterpret_cast<float4 *>(p.W + off + kb)[i] = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < kPer; ++i)
            if (kb + i < K) p.W[off + kb + i] = w[i];
    }
}

// row max from the per-tile maxima (every CTA of the row recomputes it: ntiles <= a few hundred)
__device__ __forceinline__ float row_max_from_tiles(const LargeParams &p, int row, float *s_f)
{
    float m = -INFINITY;
    for (int t = threadIdx.x; t < p.ntiles; t += blockDim.x) m = fmaxf(m, p.tmax[(size_t)row * p.ntiles + t]);
    return block_allreduce(m, -INFINITY, OpMaxF(), s_f);
}

// ---- L2 FAST: row maximum, per-tile scale exp(m_t - M) and offset, row total, lse -- one CTA per row ----
// before[t+1] = before[t] + L_last(t) * scale_t with explicitly rounded float64 operations, the same
// expression the consumers evaluate for every particle (cdf_from_tile), so the CDF is monotone across
// tile boundaries by construction.
__device__ __forceinline__ float cdf_from_tile(float L, double before, double scale)
{
    return __double2float_rn(__dadd_rn(before, __dmul_rn((double)L, scale)));
}

__global__ void __launch_bounds__(256) large_rows_fast_kernel(const LargeParams p)
{
    __shared__ float s_f[32];
    __shared__ double s_prod[1024];
    __shared__ double s_carry;
    const int row = blockIdx.x, tid = threadIdx.x, nt = p.ntiles;
    const float vmax = row_max_from_tiles(p, row, s_f);
    double *before = p.tbefore + (size_t)row * (nt + 1);
    double *scale = p.tscale + (size_t)row * nt;
    if (tid == 0) {
        p.rowmax[row] = vmax;
        if (!(fabsf(vmax) < INFINITY) && !p.rowbad[row]) { atomicOr(p.flags, AESMC_FLAG_DEGENERATE); p.rowbad[row] |= 2; }
        s_carry = 0.0;
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
    for (int base = 0; base < nt; base += 1024) {
        const int cnt = min(1024, nt - base);
        for (int t = tid; t < cnt; t += blockDim.x) {
            const float mt = p.tmax[(size_t)row * nt + base + t];
            const double sc = exp((double)mt - (double)vmax); // m_t = -inf: 0
            scale[base + t] = sc;
            s_prod[t] = __dmul_rn((double)p.tsum[(size_t)row * nt + base + t], sc);
        }
        __syncthreads();
        if (tid == 0) {
            double acc = s_carry;
            for (int t = 0; t < cnt; ++t) { before[base + t] = acc; acc = __dadd_rn(acc, s_prod[t]); }
            s_carry = acc;
        }
        __syncthreads();
    }
    if (tid == 0) {
        const double all = s_carry;
        before[nt] = all;
        const float lse = vmax + (float)log(all);
        p.rowlse[row] = lse;
        p.rowtotal[row] = __double2float_rn(all);
        if (p.lse) p.lse[row] = lse;
    }
}

// ---- L2 EXACT: scipy.special.logsumexp in numpy's summation order --------------------------------------
// numpy's pairwise tree over K particles depends on K only.  A node is named by its heap index (root 1,
// children 2h and 2h+1) and its extent is recomputed by walking the <= ~14 splits from the root, so no
// tree is ever stored: leaves hold 64..128 particles, hence each contains a multiple of 64 and the leaf
// kernel assigns one 8-lane group per such probe position (the first probe inside a leaf owns it).
__device__ __forceinline__ int pw_left(int len)
{
    const int n2 = len >> 1;
    return n2 - (n2 & 7);
}
__device__ __forceinline__ void pw_leaf_of(int K, int pos, int &start, int &len, int &heap)
{
    start = 0; len = K; heap = 1;
    while (len > 128) {
        const int n2 = pw_left(len);
        if (pos < start + n2) { len = n2; heap = 2 * heap; }
        else { start += n2; len -= n2; heap = 2 * heap + 1; }
    }
}
// length of heap node h, 0 if the tree has no such node
__device__ __forceinline__ int pw_node_len(int K, int h)
{
    int len = K;
    for (int b = 30 - __clz(h); b >= 0; --b) {
        if (len <= 128) return 0;
        const int n2 = pw_left(len);
        len = ((h >> b) & 1) ? len - n2 : n2;
    }
    return len;
}

// (a) leaves: e_i = (lw_i == max) ? 0 : np_exp(lw_i - max), 8 strided accumulators per leaf (8 lanes).
// A warp takes 32 consecutive probes: every lane walks to its probe's leaf, the owners are compacted
// with a ballot, and the four 8-lane groups evaluate them four at a time.
constexpr int kLeafThreads = 256;

__global__ void __launch_bounds__(kLeafThreads) large_leaf_kernel(const LargeParams p)
{
    __shared__ float s_f[32];
    const int row = blockIdx.y, tid = threadIdx.x, lane = tid & 31, K = p.K;
    const float *lw = p.log_w + (size_t)row * K;
    const float vmax = row_max_from_tiles(p, row, s_f);
    if (blockIdx.x == 0 && tid == 0) {
        p.rowmax[row] = vmax;
        if (!(fabsf(vmax) < INFINITY) && !p.rowbad[row]) { atomicOr(p.flags, AESMC_FLAG_DEGENERATE); atomicOr(p.rowbad + row, 2); }
    }
    if (!(fabsf(vmax) < INFINITY)) return; // degenerate row: the fold kernel writes its lse
    float *vals = p.vals + (size_t)row * p.heap_size;
    const int pos = (blockIdx.x * kLeafThreads + tid) * 64;
    int start = 0, len = 0, heap = 0;
    bool owner = false;
    if (pos < K) {
        pw_leaf_of(K, pos, start, len, heap);
        owner = (pos - 64 < start); // the first probe inside the leaf
    }
    const unsigned owners = __ballot_sync(kFull, owner);
    const int nown = __popc(owners), grp = lane >> 3, j = lane & 7;
    int cnt = 0;
    auto e1 = [&](float v) {
        const float d = __fsub_rn(v, vmax);
        cnt += (d == 0.0f);
        return (d == 0.0f) ? 0.0f : np_expf_nonpos(d);
    };
    auto e2 = [&](float v0, float v1, float &r0, float &r1) {
        const float d0 = __fsub_rn(v0, vmax), d1 = __fsub_rn(v1, vmax);
        np_expf_nonpos_pair(d0, d1, r0, r1);
        cnt += (d0 == 0.0f) + (d1 == 0.0f);
        if (d0 == 0.0f) r0 = 0.0f;
        if (d1 == 0.0f) r1 = 0.0f;
    };
    for (int r = 0; r * 4 < nown; ++r) { // warp-uniform
        const int which = r * 4 + grp;
        const bool valid = which < nown;
        const int src_lane = valid ? (int)__fns(owners, 0, which + 1) : 0;
        const int g_start = __shfl_sync(kFull, start, src_lane), g_len = __shfl_sync(kFull, len, src_lane);
        const int g_heap = __shfl_sync(kFull, heap, src_lane);
        float acc = 0.f;
        const int lim = g_len - (g_len & 7);
        if (valid) { // 64 <= g_len <= 128
            const float *src = lw + g_start + j;
            acc = e1(src[0]);
            int i = 8;
            for (; i + 24 < lim; i += 32) { // four loads in flight per lane, exps two at a time
                const float v0 = src[i], v1 = src[i + 8], v2 = src[i + 16], v3 = src[i + 24];
                float r0, r1, r2, r3;
                e2(v0, v1, r0, r1);
                e2(v2, v3, r2, r3);
                acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, r0), r1), r2), r3);
            }
            for (; i < lim; i += 8) acc = __fadd_rn(acc, e1(src[i]));
        }
        acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 1));
        acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 2));
        acc = __fadd_rn(acc, __shfl_xor_sync(kFull, acc, 4));
        if (valid && j == 0) {
            for (int i = lim; i < g_len; ++i) acc = __fadd_rn(acc, e1(lw[g_start + i]));
            vals[g_heap] = acc;
        }
    }
    cnt = warp_sum(cnt);
    if (lane == 0 && cnt) atomicAdd(p.rowcnt + row, cnt);
}

// (b) fold the levels bottom-up, finish the lse, one CTA per row
__global__ void __launch_bounds__(1024) large_fold_kernel(const LargeParams p)
{
    const int row = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const float vmax = p.rowmax[row];
    if (p.rowbad[row]) {
        if (tid == 0) {
            const float lse = (p.rowbad[row] & 1) ? __int_as_float(0x7fc00000) : vmax;
            p.rowlse[row] = lse;
            if (p.lse) p.lse[row] = lse;
        }
        return;
    }
    float *vals = p.vals + (size_t)row * p.heap_size;
    for (int d = p.heap_depth - 1; d >= 0; --d) {
        for (int h = (1 << d) + tid; h < (2 << d); h += NT)
            if (pw_node_len(p.K, h) > 128) vals[h] = __fadd_rn(vals[2 * h], vals[2 * h + 1]);
        __syncthreads();
    }
    if (tid == 0) {
        float s = vals[1];
        const float m = (float)p.rowcnt[row];
        if (s != 0.0f) s = __fdiv_rn(s, m);
        const float lse = __fadd_rn(__fadd_rn(fd_log1pf(s), np_logf(m)), vmax);
        p.rowlse[row] = lse;
        if (p.lse) p.lse[row] = lse;
    }
}

// ---- L3 EXACT: np.cumsum's sequential chain, spans chained across CTAs --------------------------------
// Span = NT * 16 particles.  A hand-off costs ~0.5 us, so 1024-thread spans (one CTA per SM) keep the
// carry chain of a row short and win while the launch is chain-bound (few rows); with many rows the
// chains overlap and 512-thread spans (two CTAs per SM, one hiding the other's barriers) have the better
// throughput.  Measured on B200, K = 1e6: B = 8: 71 vs 106 us; B = 64: 331 vs 256 us.

// Decoupled look-back (round 2).  A slot holds, in its top two bits, 0: nothing yet, 1: a MAP, 2: the span's exit
// VALUE (float bits).  A span whose blocks all lie in one binade e -- most spans of a long row: half of them sit in
// [0.5, 1) -- is, as a whole, the parity map bits -> bits + c[bits & 1]; it publishes (e, c0, c1 - c0) as soon as the
// map is composed, BEFORE its own carry has arrived, and its exit value later.  A waiting span reads the 32 slots in
// front of it in one round trip (one per lane), takes the nearest value and pushes it through the maps in between;
// every map is validated on the way (entry in binade e, no carry out of it beyond an exact hit of 2^(e+1)) -- a map
// composed against a missed estimate simply does not apply, and the waiter falls back to its predecessor's value.
// The serial chain of a row is thereby cut from one segment walk per span to one per span that crosses a binade.
struct ChainedCarry {
    static constexpr bool kChained = true;
    float est, tol;
    int span;                 // index of this span in its row (0: the chain starts at 0.0)
    unsigned long long *mine; // slots of a row are contiguous: mine[-1 - d] is the span d + 1 in front
    __device__ __forceinline__ float anchor() const { return est; }
    __device__ __forceinline__ float slack() const { return tol; }
    static __device__ __forceinline__ unsigned long long ld_slot(const unsigned long long *p)
    {
        unsigned long long v;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    }
    // called by every lane of one warp; the result is returned to all of them
    __device__ __forceinline__ float wait(int lane) const
    {
        if (span == 0) return 0.f;
        for (;;) {
            // lane d looks at the span d + 1 in front; in front of span 0 the chain value is 0.0
            const unsigned long long v = (lane < span) ? ld_slot(mine - 1 - lane) : (2ull << 62);
            const unsigned st = (unsigned)(v >> 62);
            const unsigned valmask = __ballot_sync(kFull, st == 2u), mapmask = __ballot_sync(kFull, st == 1u);
            if (!valmask) continue; // 32 maps in a row (or nothing yet): poll again
            const int dv = __ffs(valmask) - 1;
            const unsigned nearer = (1u << dv) - 1u;
            if ((mapmask & nearer) != nearer) continue; // a nearer span has published nothing yet
            int sb = __shfl_sync(kFull, (int)(unsigned)v, dv);
            bool ok = true;
            for (int k = dv - 1; k >= 0; --k) { // forward through the maps (warp-uniform)
                const unsigned long long m = __shfl_sync(kFull, v, k);
                const int e = (int)((m >> 27) & 0xffu), c0 = (int)(m & 0x1ffffffu), c1 = c0 + (int)((m >> 25) & 3u) - 1;
                const int mm = (sb & 0x7fffff) | 0x800000;
                const int c = (mm & 1) ? c1 : c0;
                if ((sb >> 23) != e || mm + c > 0x1000000) { ok = false; break; }
                sb += c; // (m + c = 2^24 carries into the exponent: exactly 2^(e+1))
            }
            if (ok) return __int_as_float(sb);
            unsigned long long pv; // a map did not apply: wait for the predecessor's own value
            do { pv = ld_slot(mine - 1); } while ((pv >> 62) != 2ull);
            return __int_as_float((int)(unsigned)pv);
        }
    }
    __device__ __forceinline__ void publish_map(int e, int c0, int c1) const
    {
        const unsigned long long v = (1ull << 62) | ((unsigned long long)(unsigned)(e & 0xff) << 27) |
                                     ((unsigned long long)(unsigned)(c1 - c0 + 1) << 25) | (unsigned long long)(unsigned)(c0 & 0x1ffffff);
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"(v) : "memory");
    }
    __device__ __forceinline__ void publish(float s) const
    {
        const unsigned long long v = (2ull << 62) | (unsigned long long)(unsigned)__float_as_int(s);
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"(v) : "memory");
    }
};

template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) large_exact_scan_kernel(const LargeParams p)
{
    constexpr int kSpan = NT * kScanItems;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int row_chunks = NT * 4 + (NT * 4 >> 3);
    float4 *bufW4 = reinterpret_cast<float4 *>(smem_raw);
    int *scratch = reinterpret_cast<int *>(bufW4 + row_chunks);
    float *bufW = reinterpret_cast<float *>(bufW4);
    __shared__ ExactScanShared s_scan;
    __shared__ double s_d[32];
    __shared__ float s_f[32];
    __shared__ float s_carry;
    __shared__ int s_ticket;
    // span-major tickets: whoever holds ticket t waits only for ticket t - B, which is already running
    if (tid == 0) s_ticket = atomicAdd(p.ticket, 1);
    __syncthreads();
    const int span = s_ticket / p.B, row = s_ticket - span * p.B;
    if (p.rowbad[row]) return;
    const size_t off = (size_t)row * p.K;
    const float lse = p.rowlse[row];
    const int base = span * kSpan;

    unsigned long long *slots = p.slots + (size_t)row * p.nspans;
    unsigned long long *lsums = p.lsums + (size_t)row * p.nspans;
    float mysum = 0.f;

    // normalised weights of this span into the padded buffer (striped, coalesced), zeros past K
    if ((p.K & 3) == 0) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = base + 4 * (tid + NT * i);
            v[i] = (k < p.K) ? __ldg(reinterpret_cast<const float4 *>(p.log_w + off + k))
                             : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 r;
            np_expf_nonpos_pair(__fsub_rn(v[i].x, lse), __fsub_rn(v[i].y, lse), r.x, r.y);
            np_expf_nonpos_pair(__fsub_rn(v[i].z, lse), __fsub_rn(v[i].w, lse), r.z, r.w);
            if (base + 4 * (tid + NT * i) >= p.K) r = make_float4(0.f, 0.f, 0.f, 0.f);
            bufW4[pad_chunk(tid + NT * i)] = r;
            mysum += (r.x + r.y) + (r.z + r.w);
        }
    } else {
        for (int e = tid; e < kSpan; e += NT) {
            const int k = base + e;
            const float r = (k < p.K) ? np_expf_nonpos(__fsub_rn(p.log_w[off + k], lse)) : 0.0f;
            bufW[pad_elem(e)] = r;
            mysum += r;
        }
    }
    // Estimate of the chain at the span start: every span publishes the plain sum of its weights as soon
    // as it has them (no dependency), and sums those of its predecessors (all started earlier: tickets).
    // The chain drifts from the real sum by ~sqrt(k) half-ulps on generic data; the tolerance is a
    // multiple of that, not the bound k * 2^-24: a miss is detected by the walker and costs one redo of
    // this span with the exact carry, never a wrong result.
    mysum = block_allreduce(mysum, 0.f, OpSumF(), s_f); // also orders the buffer writes before the reads below
    if (tid == 0) {
        const unsigned long long v = (1ull << 32) | (unsigned long long)(unsigned)__float_as_int(mysum);
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(lsums + span), "l"(v) : "memory");
    }
    double part = 0.0;
    for (int t = tid; t < span; t += NT) {
        unsigned long long v;
        do {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(lsums + t) : "memory");
        } while ((v >> 32) == 0ull);
        part += (double)__int_as_float((int)(unsigned)v);
    }
    part = block_allreduce(part, 0.0, OpSumD(), s_d);
    ChainedCarry cc;
    cc.est = (float)part;
    cc.tol = cc.est * 5.9604644775390625e-08f * (8.0f * sqrtf((float)base) + 64.0f);
    cc.span = span;
    cc.mine = slots + span;
    float w[kScanItems];
    auto load_block = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 v = bufW4[pad_chunk(4 * tid + i)];
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    };
    load_block();
    float total;
    if (!exact_cumsum_blocked(w, &total, bufW4, scratch, s_scan, cc)) {
        const bool published = (s_scan.fail == 0); // the walker succeeded, a replay check did not
        const float carry = s_scan.carry_in;
        if (tid == 0) atomicAdd(p.stats + (s_scan.fail == 2 ? 0 : 1), 1);
        __syncthreads();
        load_block();
        if (!exact_cumsum_blocked(w, &total, bufW4, scratch, s_scan, LocalCarry{carry})) {
            if (tid == 0) { // plain sequential chain over the span
                atomicAdd(p.stats + 2, 1);
                float acc = carry;
                for (int e = 0; e < kSpan; ++e) { acc = __fadd_rn(acc, bufW[pad_elem(e)]); bufW[pad_elem(e)] = acc; }
                s_carry = acc;
            }
            __syncthreads();
            total = s_carry;
#pragma unroll
            for (int j = 0; j < kScanItems; ++j) w[j] = bufW[pad_elem(kScanItems * tid + j)];
        }
        if (!published && tid == 0) cc.publish(total);
    }
    const int kb = base + kScanItems * tid;
    if ((p.K & 3) == 0 && kb + kScanItems <= p.K) {
        float4 *dst = reinterpret_cast<float4 *>(p.W + off + kb);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < kScanItems; ++j)
            if (kb + j < p.K) p.W[off + kb + j] = w[j];
    }
    if (span == p.nspans - 1 && tid == 0) p.rowtotal[row] = total;
}

// ---- L4 + L5: offspring boundaries, ancestors and gather, fused per OUTPUT tile ---------------------------
// Round 1 ran an input-centric search (scatter of run marks into the zeroed idx output: a memset, a pass of
// sector read-modify-writes and a read, 16 bytes of HBM traffic per particle for a table of marks) followed by an
// expansion per output tile.  Here nothing of the sort touches HBM: a tiny kernel tabulates the boundary count at
// the END of every input tile, cend[t] = c_{4096 (t + 1) - 1} (monotone in t); the CTA of output tile T finds, by
// binary search in that table, the input tiles whose particles can own positions of [4096 T, 4096 T + 4096),
// recomputes their boundaries from the CDF table (tiles without a single offspring are skipped from the table
// alone, 16-particle blocks that miss the output tile after two evaluations), drops the run starts into a
// shared-memory tile, max-scans it and writes ancestors and gathered latents.  A particle's boundaries are
// recomputed by every output tile that looks at its input tile (~2 on average): arithmetic for traffic.
template <bool EXACT, bool TILED> struct Boundary {
    const LargeParams &p;
    float total, rcp, u32, Kf;
    bool safe_total, filtered;
    double u, Kd, band;
    int K;
    __device__ __forceinline__ Boundary(const LargeParams &p_, int row) : p(p_)
    {
        K = p.K;
        total = p.rowtotal[row];
        rcp = refined_rcp(total);
        safe_total = total > 9.3132257e-10f && total < 2.0f;
        u = p.u[row];
        u32 = (float)u; Kf = (float)K;
        Kd = (double)K; band = Kd * 8.8817841970012523e-16;
        // the float32 pre-filter sends 2 * tol32 of the particles to float64 anyway: past ~1 % nearly every
        // warp runs both forms, so large rows use the float64 form alone
        filtered = K <= 65536;
    }
    // c: a CDF entry (exact mode: the reference's float64 comparison, common.cuh)
    __device__ __forceinline__ int operator()(float c, double, double) const
    {
        const float cdfn = div_hoisted(c, total, rcp, safe_total);
        return filtered ? count_positions_below_filtered(cdfn, u, u32, K, Kf, p.tol32)
                        : count_positions_below(cdfn, u, K, Kd, band);
    }
    // FAST mode owes nobody the reference's float64 rounding of (u + k) / K: #{k >= 0 : u + k < cdfn K} = ceil(cdfn K - u)
    // evaluated EXACTLY in 32.32 fixed point -- cdfn = m 2^e is a float32, so m K is an exact 45-bit integer; u is
    // taken to 2^-32 -- a dozen integer instructions instead of the float64 sequence (which rows beyond 65 536
    // particles would run for every particle), monotone in cdfn by construction.
    unsigned long long u_fix;
    __device__ __forceinline__ void init_fast() { u_fix = (unsigned long long)(u * 4294967296.0); }
    __device__ __forceinline__ int from_cdf_fast(float cdf) const
    {
        const float cdfn = fminf(__fmul_rn(cdf, rcp), 1.0f);
        const int b = __float_as_int(cdfn);
        int ex = b >> 23; // cdfn >= 0
        const unsigned m = ex ? ((b & 0x7fffff) | 0x800000) : (b & 0x7fffff);
        ex = max(ex, 1);
        const unsigned long long P = (unsigned long long)m * (unsigned)K; // cdfn K = P 2^(ex - 150)
        const int sh = ex - 150 + 32;                                     // <= 9 because cdfn <= 1
        const unsigned long long T = sh >= 0 ? (P << sh) : (sh > -64 ? (P >> (-sh)) : 0ull);
        const long long n = (long long)(T - u_fix);
        const int c = n <= 0 ? 0 : (int)((unsigned long long)(n + 0xffffffffll) >> 32);
        return min(c, K);
    }
};
// FAST: the CDF entry of a particle from its tile-local running sum L, in float32 with one rounding, clamped to the
// next tile's offset -- which IS, by definition, the entry of the tile's last particle: monotone inside a tile (fma
// is monotone in L) and across tiles (the clamp), consistent between the table of tile ends and the particles
struct TileCdf {
    float b0, b1, sc;
    __device__ __forceinline__ TileCdf(const LargeParams &p, int row, int t, bool on = true)
    {
        b0 = b1 = sc = 0.f;
        if (!on) return;
        const double *before = p.tbefore + (size_t)row * (p.ntiles + 1) + t;
        b0 = __double2float_rn(before[0]); b1 = __double2float_rn(before[1]);
        sc = __double2float_rn(p.tscale[(size_t)row * p.ntiles + t]);
    }
    __device__ __forceinline__ float operator()(float L) const { return fminf(__fmaf_rn(L, sc, b0), b1); }
};

// cend[row][t]: boundary count of the last particle of input tile t (K for the last tile: particle K-1 owns every
// remaining position).  One thread per tile.
template <bool EXACT, bool TILED>
__global__ void __launch_bounds__(256) large_bounds_kernel(const LargeParams p)
{
    const int row = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.ntiles || p.rowbad[row]) return;
    int c = p.K;
    if (t + 1 < p.ntiles) {
        Boundary<EXACT, TILED> boundary(p, row);
        if (TILED) {
            boundary.init_fast();
            c = boundary.from_cdf_fast(TileCdf(p, row, t).b1);
        } else {
            c = boundary(p.W[(size_t)row * p.K + (size_t)(t + 1) * kTile - 1], 0.0, 0.0);
        }
    }
    p.tenter[(size_t)row * p.ntiles + t] = c;
}

// Global traffic is chunk-striped (thread t owns 16-byte chunks t + 256 i: fully coalesced, and the
// gather of a warp stays within a few sectors because ancestors are sorted); the max-scan wants 16
// consecutive positions per thread; the padded tile buffer converts between the two.
#ifndef AESMC_LARGE_RS_CTAS
#define AESMC_LARGE_RS_CTAS 4
#endif
template <bool EXACT, bool TILED>
__global__ void __launch_bounds__(kTileThreads, AESMC_LARGE_RS_CTAS) large_resample_kernel(const LargeParams p)
{
    extern __shared__ __align__(16) int s_dyn[];
    int *s_tile = s_dyn;                          // [kTile + kTile / 8] run marks, then ancestors, of the output tile
    int *s_cend = s_dyn + kTile + kTile / 8;      // [ntiles]
    __shared__ int s_w[32];
    constexpr int kPer = kTile / kTileThreads;
    const int row = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t off = (size_t)row * p.K;
    const int K = p.K, p0 = tile * kTile, n = min(kTile, K - p0), p1 = p0 + n, nt = p.ntiles;
    if (p.rowbad[row]) { // identity ancestry keeps downstream gathers in range
        for (int k = tid; k < n; k += kTileThreads) p.idx[off + p0 + k] = p0 + k;
        if (p.x_in) {
            const size_t xo = (off + p0) * p.D;
            for (int e = tid; e < n * p.D; e += kTileThreads) p.x_out[xo + e] = p.x_in[xo + e];
        }
        return;
    }
    const bool vec = (K & 3) == 0;
    int4 *s_tile4 = reinterpret_cast<int4 *>(s_tile);
    for (int t = tid; t < nt; t += kTileThreads) s_cend[t] = p.tenter[(size_t)row * nt + t];
    for (int c = tid; c < (kTile + kTile / 8) / 4; c += kTileThreads) s_tile4[c] = make_int4(0, 0, 0, 0);
    __syncthreads();
    // input tiles that can own positions of [p0, p1): first t with cend[t] > p0 ... first t with cend[t] >= p1
    int t_first, t_last;
    {
        int lo = 0, hi = nt - 1; // cend[nt - 1] = K > p0
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_cend[mid] > p0) hi = mid; else lo = mid + 1; }
        t_first = lo;
        hi = nt - 1;             // cend[nt - 1] = K >= p1
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_cend[mid] >= p1) hi = mid; else lo = mid + 1; }
        t_last = lo;
    }
    Boundary<EXACT, TILED> bnd(p, row);
    if (TILED) bnd.init_fast();
    const float *Wrow = p.W + off;
    for (int t = t_first; t <= t_last; ++t) {
        const int c_tile_in = t ? s_cend[t - 1] : 0;       // boundary of the particle in front of the tile
        if (s_cend[t] == c_tile_in) continue;               // not one offspring in the whole tile
        const TileCdf tcdf(p, row, t, TILED);
        auto boundary = [&](float v, double, double) { return TILED ? bnd.from_cdf_fast(tcdf(v)) : bnd(v, 0.0, 0.0); };
        const double t_before = 0.0, t_scale = 0.0;
        const int j0 = t * kTile + kPer * tid;
        // the block's 16 CDF entries, loaded before anybody knows whether the block matters: one global round trip
        // per input tile instead of two (the kernel is latency-bound: ~3 input tiles per output tile, 4 CTAs per SM;
        // measured: 562 -> 529 us per step at B = 64, K = 10^6.  Letting every thread evaluate the boundary in front
        // of its block itself -- no exchange, no barrier in this loop -- was slower again: 562 us)
        float cdf[kPer];
        if (vec) {
#pragma unroll
            for (int i = 0; i < kPer / 4; ++i) {
                const float4 v = (j0 + 4 * i < K) ? __ldg(reinterpret_cast<const float4 *>(Wrow + j0) + i)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                cdf[4 * i] = v.x; cdf[4 * i + 1] = v.y; cdf[4 * i + 2] = v.z; cdf[4 * i + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < kPer; ++i) cdf[i] = (j0 + i < K) ? Wrow[j0 + i] : 0.f;
        }
        // the boundary at the end of this thread's block, and (from its neighbour) of the particle in front of it
        int c_end = s_cend[t];
        if (j0 + kPer < (t + 1) * kTile && j0 + kPer < K) c_end = boundary(cdf[kPer - 1], t_before, t_scale);
        else if (j0 >= K) c_end = K;
        int c_prev = __shfl_up_sync(kFull, c_end, 1);
        if (lane == 0) c_prev = tid ? 0 : c_tile_in;
        __syncthreads(); // warp boundaries travel through shared memory (the previous pass's readers are done)
        if (lane == 31) s_w[warp] = c_end;
        __syncthreads();
        if (lane == 0 && tid) c_prev = s_w[warp - 1];
        if (j0 < K && c_end > c_prev && c_prev < p1 && c_end > p0) { // the block owns positions of this output tile
            const int last_i = K - 1 - j0; // particle K-1 owns every remaining position; later slots are past the row
            int cp = c_prev;
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                int c = (i == kPer - 1) ? c_end : boundary(cdf[i], t_before, t_scale);
                if (i >= last_i) c = K;
                if (c > cp && cp < p1 && c > p0) s_tile[pad_elem(max(cp, p0) - p0)] = j0 + i; // one writer per position
                cp = c;
            }
        }
    }
    __syncthreads();
    int m[kPer];
#pragma unroll
    for (int i = 0; i < kPer / 4; ++i) {
        const int4 v = s_tile4[pad_chunk(4 * tid + i)];
        m[4 * i] = v.x; m[4 * i + 1] = v.y; m[4 * i + 2] = v.z; m[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 1; i < kPer; ++i) m[i] = max(m[i], m[i - 1]);
    int incl = m[kPer - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl = max(incl, v);
    }
    int pre = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) pre = 0;
    __syncthreads(); // (s_w carried the boundary exchange above)
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    for (int v = 0; v < warp; ++v) pre = max(pre, s_w[v]);
#pragma unroll
    for (int i = 0; i < kPer / 4; ++i)
        s_tile4[pad_chunk(4 * tid + i)] = make_int4(max(m[4 * i], pre), max(m[4 * i + 1], pre), max(m[4 * i + 2], pre), max(m[4 * i + 3], pre));
    __syncthreads();
    if (vec) {
        const bool gather1 = p.x_in && p.D == 1;
        const float *xin = p.x_in + off;
#pragma unroll
        for (int i = 0; i < kPer / 4; ++i) {
            const int c = tid + kTileThreads * i;
            if (4 * c < n) {
                const int4 v = s_tile4[pad_chunk(c)];
                reinterpret_cast<int4 *>(p.idx + off + p0)[c] = v;
                if (gather1)
                    reinterpret_cast<float4 *>(p.x_out + off + p0)[c] = make_float4(__ldg(xin + v.x), __ldg(xin + v.y), __ldg(xin + v.z), __ldg(xin + v.w));
            }
        }
        if (!p.x_in || gather1) return;
    } else {
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int k = tid + kTileThreads * i;
            if (k < n) p.idx[off + p0 + k] = s_tile[pad_elem(k)];
        }
        if (!p.x_in) return;
    }
    gather_rows(p.x_in + off * p.D, p.x_out + (off + p0) * p.D, s_tile, n, p.gather);
}

// ---- host side ----------------------------------------------------------------------------------------
static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// depth of the deepest leaf of numpy's pairwise tree over n elements (host; a few dozen distinct sizes)
static int pairwise_depth(int n)
{
    if (n <= 128) return 0;
    int n2 = n / 2;
    n2 -= n2 % 8;
    const int l = pairwise_depth(n2);
    return 1 + (n - n2 == n2 ? l : std::max(l, pairwise_depth(n - n2)));
}

int64_t smc_step_large_workspace_bytes(int64_t B, int64_t K)
{
    const int64_t nt = (K + kTile - 1) / kTile;
    size_t bytes = 0;
    bytes += align_up((size_t)B * K * 4);          // W
    bytes += align_up((size_t)B * nt * 4) * 3;     // tmax, tsum, tenter
    bytes += align_up((size_t)B * (nt + 1) * 8) * 2; // tbefore, tscale
    bytes += align_up((size_t)B * 4) * 4;          // rowmax, rowtotal, rowlse, rowbad
    const size_t nspans = (size_t)((K + 8191) / 8192); // room for the smallest span size
    bytes += align_up((size_t)B * ((size_t)2 << pairwise_depth((int)K)) * 4); // heap of node values
    bytes += align_up((size_t)B * 4 + (size_t)B * nspans * 16 + 8 + 16); // rowcnt, ticket, carry and sum slots, stats (one memset)
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
    p.gather = rows_gather_params(K, D);
    p.tol32 = (float)K * 1.1920928955078125e-07f + 5.9604644775390625e-08f;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    p.W = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * K * 4);
    p.marks = nullptr; // (round 1 kept a table of run marks in the idx output; the fused resampling kernel needs none)
    p.tmax = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * p.ntiles * 4);
    p.tsum = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * p.ntiles * 4);
    p.tenter = reinterpret_cast<int *>(ws); ws += align_up((size_t)B * p.ntiles * 4);
    p.tbefore = reinterpret_cast<double *>(ws); ws += align_up((size_t)B * (p.ntiles + 1) * 8);
    p.tscale = reinterpret_cast<double *>(ws); ws += align_up((size_t)B * (p.ntiles + 1) * 8);
    p.tiled = exact ? 0 : 1;
    p.rowmax = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * 4);
    p.rowtotal = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * 4);
    p.rowlse = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * 4);
    p.rowbad = reinterpret_cast<int *>(ws); ws += align_up((size_t)B * 4);
    static const int span_env = getenv("AESMC_SPAN_THREADS") ? atoi(getenv("AESMC_SPAN_THREADS")) : 0;
    int span_threads = (B >= 24) ? 512 : 1024;
    if (span_env == 512 || span_env == 1024) span_threads = span_env;
    p.nspans = (int)((K + span_threads * kScanItems - 1) / (span_threads * kScanItems));
    p.heap_depth = pairwise_depth((int)K);
    p.heap_size = 2 << p.heap_depth;
    p.vals = reinterpret_cast<float *>(ws); ws += align_up((size_t)B * p.heap_size * 4);
    const size_t zero_bytes = (size_t)B * 4 + (size_t)B * p.nspans * 16 + 8 + 16;
    p.slots = reinterpret_cast<unsigned long long *>(ws);
    p.lsums = p.slots + (size_t)B * p.nspans;
    p.ticket = reinterpret_cast<int *>(ws + (size_t)B * p.nspans * 16);
    p.stats = p.ticket + 2;
    p.rowcnt = p.stats + 4;
    cudaError_t e = cudaMemsetAsync(p.rowbad, 0, (size_t)B * 4, stream);
    if (e == cudaSuccess && exact) e = cudaMemsetAsync(p.slots, 0, zero_bytes, stream);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
    const dim3 grid((unsigned)p.ntiles, (unsigned)B);
    if (exact) large_prep_kernel<false><<<grid, kTileThreads, 0, stream>>>(p);
    else large_prep_kernel<true><<<grid, kTileThreads, 0, stream>>>(p);
    count_launch();
    if (exact) {
        const int nprobes = (int)((K + 63) / 64);
        const dim3 lgrid((unsigned)((nprobes + kLeafThreads - 1) / kLeafThreads), (unsigned)B);
        large_leaf_kernel<<<lgrid, kLeafThreads, 0, stream>>>(p);
        count_launch();
        large_fold_kernel<<<(unsigned)B, 1024, 0, stream>>>(p);
        count_launch();
        if (idx) {
            const size_t row_chunks = (size_t)span_threads * 4 + ((size_t)span_threads * 4 >> 3);
            const size_t smem_scan = row_chunks * 16 + (size_t)(8 * span_threads + 8) * 4;
            auto launch = [&](auto kernel) {
                cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan);
                if (err != cudaSuccess) return err;
                kernel<<<(unsigned)(B * p.nspans), span_threads, smem_scan, stream>>>(p);
                return cudaSuccess;
            };
            e = span_threads == 1024 ? launch(large_exact_scan_kernel<1024>) : launch(large_exact_scan_kernel<512>);
            if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
            count_launch();
        }
    } else {
        large_rows_fast_kernel<<<(unsigned)B, 256, 0, stream>>>(p);
        count_launch();
    }
    if (idx) {
        const dim3 bgrid((unsigned)((p.ntiles + 255) / 256), (unsigned)B);
        const size_t smem_rs = (size_t)(kTile + kTile / 8 + p.ntiles) * 4;
        if (smem_rs > 200 * 1024) { set_error("aesmc_smc_step_ws_f32: K=%lld exceeds the multi-CTA path (tile table)", (long long)K); return AESMC_ERR_UNSUPPORTED; }
        auto launch_rs = [&](auto bounds, auto resample) {
            cudaError_t err = cudaFuncSetAttribute(resample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rs);
            if (err != cudaSuccess) return err;
            bounds<<<bgrid, 256, 0, stream>>>(p);
            count_launch();
            resample<<<grid, kTileThreads, smem_rs, stream>>>(p);
            count_launch();
            return cudaSuccess;
        };
        e = exact ? launch_rs(large_bounds_kernel<true, false>, large_resample_kernel<true, false>)
                  : launch_rs(large_bounds_kernel<false, true>, large_resample_kernel<false, true>);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
    }
    if (exact && idx && getenv("AESMC_DEBUG_STATS")) { // debugging aid: synchronises
        int h[4] = {0, 0, 0, 0};
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, p.stats, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[aesmc] chained scan: %d spans, redo after missed estimate %d, after replay check %d, sequential %d\n",
                (int)(B * p.nspans), h[0], h[1], h[2]);
    }
    return check_launch("smc_step_large");
}

} // namespace aesmc
