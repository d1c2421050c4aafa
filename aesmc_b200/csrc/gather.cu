// gather.cu -- ancestral gather (state.py:158-183), its backward, and genealogy index utilities
// (inference.py:196-231).  Pure data movement: vectorised to the widest unit the particle size and
// pointer alignment allow (16 B -> LDG.128/STG.128), coalesced along the particle axis.
#include "common.cuh"

namespace aesmc {

template <typename IdxT>
__device__ __forceinline__ int load_index(const IdxT *idx_row, int k, int K, int32_t *flags)
{
    long long id = (long long)idx_row[k];
    if (id < 0 || id >= K) {
        if (flags) atomicOr(flags, AESMC_FLAG_INDEX_RANGE);
        id = id < 0 ? 0 : K - 1;
    }
    return (int)id;
}

// k = e / d by multiply-high with ceil(2^32 / d): exact for e < 2^32 / d; mul == ~0u: plain division
struct FlatDiv {
    unsigned d, mul;
};
static FlatDiv flat_div(int64_t n_rows, int64_t d)
{
    FlatDiv f;
    f.d = (unsigned)d;
    f.mul = d > 1 ? (unsigned)(((1ull << 32) + (unsigned long long)d - 1) / (unsigned long long)d) : 0u;
    if ((unsigned long long)n_rows * (unsigned long long)d * (unsigned long long)d >= (1ull << 32)) f.mul = 0xffffffffu;
    return f;
}
__device__ __forceinline__ int flat_row(int e, const FlatDiv f)
{
    return f.d == 1 ? e : (f.mul == 0xffffffffu ? (int)((unsigned)e / f.d) : (int)__umulhi((unsigned)e, f.mul));
}

// dst[b,k,q] = src[b, idx[b,k], q] for q in [0, upr) units of type V per particle: the (k, q) plane is swept
// flat, so stores are fully coalesced; U loads in flight per thread.
template <typename V, typename IdxT>
__global__ void __launch_bounds__(256) gather_kernel(const V *__restrict__ src, const IdxT *__restrict__ idx, int B, int K,
                                                     const FlatDiv upr, V *__restrict__ dst, int32_t *flags)
{
    constexpr int U = sizeof(V) >= 16 ? 2 : 4;
    const int step = gridDim.x * blockDim.x, n = K * (int)upr.d;
    for (int row = blockIdx.y; row < B; row += gridDim.y) {
        const size_t roff = (size_t)row * K;
        const IdxT *irow = idx + roff;
        const V *srow = src + roff * upr.d;
        V *drow = dst + roff * upr.d;
        for (int e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += U * step) {
            V v[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int e = e0 + q * step;
                if (e < n) {
                    const int k = flat_row(e, upr);
                    v[q] = __ldg(srow + (size_t)load_index(irow, k, K, flags) * upr.d + (e - k * (int)upr.d));
                }
            }
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int e = e0 + q * step;
                if (e < n) drow[e] = v[q];
            }
        }
    }
}

// 4-byte particles (scalar float latents, the common case), K % 4 == 0, 16-byte aligned rows: four
// outputs per thread, indices and results move as 16-byte vectors
template <typename IdxT>
__global__ void __launch_bounds__(256) gather4_kernel(const uint32_t *__restrict__ src, const IdxT *__restrict__ idx, int B,
                                                      int K, uint32_t *__restrict__ dst, int32_t *flags)
{
    const int step = gridDim.x * blockDim.x, n4 = K >> 2;
    for (int row = blockIdx.y; row < B; row += gridDim.y) {
        const size_t roff = (size_t)row * K;
        const IdxT *irow = idx + roff;
        const uint32_t *srow = src + roff;
        uint4 *drow = reinterpret_cast<uint4 *>(dst + roff);
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n4; c += step) {
            long long i0, i1, i2, i3;
            if (sizeof(IdxT) == 4) {
                const int4 v = __ldg(reinterpret_cast<const int4 *>(irow) + c);
                i0 = v.x; i1 = v.y; i2 = v.z; i3 = v.w;
            } else {
                const longlong2 v0 = __ldg(reinterpret_cast<const longlong2 *>(irow) + 2 * c);
                const longlong2 v1 = __ldg(reinterpret_cast<const longlong2 *>(irow) + 2 * c + 1);
                i0 = v0.x; i1 = v0.y; i2 = v1.x; i3 = v1.y;
            }
            if ((i0 | i1 | i2 | i3) < 0 || i0 >= K || i1 >= K || i2 >= K || i3 >= K) {
                if (flags) atomicOr(flags, AESMC_FLAG_INDEX_RANGE);
                i0 = i0 < 0 ? 0 : (i0 >= K ? K - 1 : i0); i1 = i1 < 0 ? 0 : (i1 >= K ? K - 1 : i1);
                i2 = i2 < 0 ? 0 : (i2 >= K ? K - 1 : i2); i3 = i3 < 0 ? 0 : (i3 >= K ? K - 1 : i3);
            }
            drow[c] = make_uint4(__ldg(srow + i0), __ldg(srow + i1), __ldg(srow + i2), __ldg(srow + i3));
        }
    }
}

static unsigned gather_grid_x(int64_t n_per_row, int per_thread, int64_t B)
{
    // enough CTAs per row to cover it in one or two sweeps, fewer when there are plenty of rows
    int64_t gx = (n_per_row + 256 * (int64_t)per_thread - 1) / (256 * (int64_t)per_thread);
    const int64_t cap = B >= 1024 ? 4 : 64;
    if (gx > cap) gx = cap;
    return (unsigned)(gx < 1 ? 1 : gx);
}

template <typename V, typename IdxT>
static void launch_gather_t(const void *src, const void *idx, int64_t B, int64_t K, int64_t upr, void *dst,
                            int32_t *flags, cudaStream_t st)
{
    const unsigned gy = (unsigned)(B < 65535 ? B : 65535);
    const bool idx_aligned = (reinterpret_cast<uintptr_t>(idx) & 15) == 0;
    if (sizeof(V) == 4 && upr == 1 && (K & 3) == 0 && idx_aligned && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        gather4_kernel<IdxT><<<dim3(gather_grid_x(K / 4, 1, B), gy), 256, 0, st>>>(
            static_cast<const uint32_t *>(src), static_cast<const IdxT *>(idx), (int)B, (int)K, static_cast<uint32_t *>(dst), flags);
        return;
    }
    gather_kernel<V, IdxT><<<dim3(gather_grid_x(K * upr, 4, B), gy), 256, 0, st>>>(
        static_cast<const V *>(src), static_cast<const IdxT *>(idx), (int)B, (int)K, flat_div(K, upr), static_cast<V *>(dst), flags);
}

int launch_gather_bytes(const void *src, const void *idx, int idx_is_i64, int64_t B, int64_t K, int64_t row_bytes,
                        void *dst, int32_t *flags, cudaStream_t st)
{
    const uintptr_t align = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)row_bytes;
#define AESMC_GATHER(V)                                                                                         \
    do {                                                                                                        \
        if (idx_is_i64) launch_gather_t<V, int64_t>(src, idx, B, K, row_bytes / (int64_t)sizeof(V), dst, flags, st); \
        else launch_gather_t<V, int32_t>(src, idx, B, K, row_bytes / (int64_t)sizeof(V), dst, flags, st);       \
    } while (0)
    if ((align & 15) == 0) AESMC_GATHER(uint4);
    else if ((align & 7) == 0) AESMC_GATHER(uint2);
    else if ((align & 3) == 0) AESMC_GATHER(uint32_t);
    else if ((align & 1) == 0) AESMC_GATHER(uint16_t);
    else AESMC_GATHER(uint8_t);
#undef AESMC_GATHER
    count_launch();
    return check_launch("gather_kernel");
}

// Backward for non-decreasing indices (what the step kernel emits): the children of parent j form one
// contiguous run of k.  The (k, component) plane is swept flat; the thread that sees a run start sums the
// run in k order -- the same order as the reference's CPU scatter_add, so the result is deterministic and
// needs no atomics -- and also zero-fills the childless parents between the previous run's parent and its
// own (the thread of the last particle fills those after it): every entry of gsrc is written exactly once,
// no memset.  VT: vector of VW components of T (the widest that divides D).
template <typename T> struct VecOf;
template <> struct VecOf<float> { using v2 = float2; using v4 = float4; };
template <> struct VecOf<double> { using v2 = double2; using v4 = double4; };
__device__ __forceinline__ void vacc(float &a, float b) { a += b; }
__device__ __forceinline__ void vacc(double &a, double b) { a += b; }
__device__ __forceinline__ void vacc(float2 &a, float2 b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ void vacc(double2 &a, double2 b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ void vacc(float4 &a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

__device__ __forceinline__ float vshfl(float v, int m) { return __shfl_xor_sync(kFull, v, m); }
__device__ __forceinline__ double vshfl(double v, int m) { return __shfl_xor_sync(kFull, v, m); }
__device__ __forceinline__ float2 vshfl(float2 v, int m) { return make_float2(vshfl(v.x, m), vshfl(v.y, m)); }
__device__ __forceinline__ double2 vshfl(double2 v, int m) { return make_double2(vshfl(v.x, m), vshfl(v.y, m)); }
__device__ __forceinline__ float4 vshfl(float4 v, int m) { return make_float4(vshfl(v.x, m), vshfl(v.y, m), vshfl(v.z, m), vshfl(v.w, m)); }

// A thread owns P consecutive particles (P = 4 when a particle is one vector, else 1) and walks their
// components Q at a time, so P * Q independent loads are in flight and the indices are read once per particle.
// A run that is still open at the end of the thread's particles is followed by that thread for up to kTail
// more children (k order, the reference's scatter_add order); the particles of a run that began before the
// thread's first one belong to the thread that saw its start.  Longer runs -- collapsed weights put thousands
// of children under one parent, which a single thread would chase for hundreds of microseconds -- and long
// stretches of childless parents are finished by the whole warp: coalesced 32-particle windows, per-lane
// partial sums, one fixed shuffle tree (deterministic; the summation order of those runs is no longer k).
constexpr int kTail = 16;

template <typename VT, typename IdxT, int P, int Q>
__global__ void __launch_bounds__(256, sizeof(VT) * Q <= 8 ? 6 : 4) gather_bwd_sorted_kernel(const VT *__restrict__ g, const IdxT *__restrict__ idx, int B,
                                                                int K, int D, VT *__restrict__ gsrc)
{
    constexpr int kWin = Q == 1 ? 4 : 1; // windows in flight in the cooperative section (register budget)
    const int step = gridDim.x * blockDim.x * P, lane = threadIdx.x & 31;
    VT zero;
    memset(&zero, 0, sizeof(VT));
    auto clampi = [K](long long v) { return (int)(v < 0 ? 0 : (v >= K ? K - 1 : v)); };
    for (int row = blockIdx.y; row < B; row += gridDim.y) {
        const size_t roff = (size_t)row * K;
        const IdxT *irow = idx + roff;
        const VT *grow = g + roff * D;
        VT *orow = gsrc + roff * D;
        // warp-uniform trip count: the cooperative sections below need every lane
        for (int kw = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * P; kw < K; kw += step) {
            const int k0 = kw + lane * P;
            const int np = max(0, min(P, K - k0));
            int id[P + 2]; // ancestors of k0 - 1, k0 .. k0 + P - 1, k0 + P (clamped: memory safety only; -1 / K: none)
            id[0] = -1;
            if (k0 && np) id[0] = clampi((long long)irow[k0 - 1]);
#pragma unroll
            for (int p = 0; p <= P; ++p) {
                id[p + 1] = K;
                if (k0 + p < K) id[p + 1] = clampi((long long)irow[k0 + p]);
            }
            for (int q0 = 0; q0 < D; q0 += Q) {
                VT gv[P][Q];
#pragma unroll
                for (int p = 0; p < P; ++p)
#pragma unroll
                    for (int q = 0; q < Q; ++q)
                        if (p < np && q0 + q < D) gv[p][q] = grow[(size_t)(k0 + p) * D + q0 + q];
                VT cur[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) cur[q] = zero;
                bool open = false;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (p < np) {
                        if (id[p + 1] != id[p]) { // run start
                            open = true;
#pragma unroll
                            for (int q = 0; q < Q; ++q) cur[q] = gv[p][q];
                        } else if (open) {
#pragma unroll
                            for (int q = 0; q < Q; ++q) vacc(cur[q], gv[p][q]);
                        }
                        if (open && p + 1 < np && id[p + 2] != id[p + 1]) { // the run ends inside the thread's particles
#pragma unroll
                            for (int q = 0; q < Q; ++q)
                                if (q0 + q < D) orow[(size_t)id[p + 1] * D + q0 + q] = cur[q];
                            open = false;
                        }
                    }
                }
                int j = 0; // ancestor of the thread's last particle (no dynamic indexing: id[] stays in registers)
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if (p + 1 == np) j = id[p + 1];
                int next = k0 + np; // first particle not yet accounted for
                bool long_run = false;
                if (open) { // follow the last run
                    int cnt = 0;
                    for (; next < K && cnt < kTail && (long long)irow[next] == (long long)j; ++next, ++cnt)
#pragma unroll
                        for (int q = 0; q < Q; ++q)
                            if (q0 + q < D) vacc(cur[q], grow[(size_t)next * D + q0 + q]);
                    long_run = cnt == kTail && next < K && (long long)irow[next] == (long long)j;
                    if (!long_run) {
#pragma unroll
                        for (int q = 0; q < Q; ++q)
                            if (q0 + q < D) orow[(size_t)j * D + q0 + q] = cur[q];
                    }
                }
                for (unsigned pend = __ballot_sync(kFull, long_run); pend; pend &= pend - 1) {
                    const int src = __ffs(pend) - 1;
                    const int jj = __shfl_sync(kFull, j, src);
                    int start = __shfl_sync(kFull, next, src);
                    VT part[Q];
#pragma unroll
                    for (int q = 0; q < Q; ++q) part[q] = zero;
                    for (;;) { // idx is sorted: the lanes still inside the run form a prefix of each 32-particle window
                        bool in[kWin];
#pragma unroll
                        for (int w = 0; w < kWin; ++w) { // kWin windows in flight: the exit test is a full round trip
                            const int i = start + 32 * w + lane;
                            in[w] = i < K && (long long)irow[i] == (long long)jj;
                        }
#pragma unroll
                        for (int w = 0; w < kWin; ++w) {
                            const int i = start + 32 * w + lane;
                            if (in[w]) {
#pragma unroll
                                for (int q = 0; q < Q; ++q)
                                    if (q0 + q < D) vacc(part[q], grow[(size_t)i * D + q0 + q]);
                            }
                        }
                        if (__ballot_sync(kFull, in[kWin - 1]) != kFull) break;
                        start += 32 * kWin;
                    }
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
#pragma unroll
                        for (int m = 16; m > 0; m >>= 1) vacc(part[q], vshfl(part[q], m));
                        if (lane == src && q0 + q < D) {
                            vacc(cur[q], part[q]);
                            orow[(size_t)jj * D + q0 + q] = cur[q];
                        }
                    }
                }
            }
            // childless parents: between the previous particle's ancestor and each run start of this thread, and
            // after the last particle of the row; long stretches are left to the whole warp
            int gap_lo = 0, gap_hi = 0; // at most one deferred stretch per thread and pass is kept; others are filled here
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (p < np && id[p + 1] != id[p]) {
                    const int lo = (id[p] + 1) * D, hi = id[p + 1] * D;
                    if (hi - lo > 64 && gap_hi == gap_lo) { gap_lo = lo; gap_hi = hi; }
                    else for (int e = lo; e < hi; ++e) orow[e] = zero;
                }
            }
            if (np && k0 + np == K) {
                int jl = 0;
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if (p + 1 == np) jl = id[p + 1];
                const int lo = (jl + 1) * D, hi = K * D;
                if (hi - lo > 64 && gap_hi == gap_lo) { gap_lo = lo; gap_hi = hi; }
                else for (int e = lo; e < hi; ++e) orow[e] = zero;
            }
            for (unsigned pend = __ballot_sync(kFull, gap_hi > gap_lo); pend; pend &= pend - 1) {
                const int src = __ffs(pend) - 1;
                const int lo = __shfl_sync(kFull, gap_lo, src), hi = __shfl_sync(kFull, gap_hi, src);
                for (int e = lo + lane; e < hi; e += 32) orow[e] = zero;
            }
        }
    }
}

// Backward for arbitrary indices: atomic scatter-add into zero-initialised gsrc.
template <typename T, typename IdxT>
__global__ void gather_bwd_atomic_kernel(const T *__restrict__ g, const IdxT *__restrict__ idx, int B, int K, int D,
                                         T *__restrict__ gsrc)
{
    for (int row = blockIdx.y; row < B; row += gridDim.y) {
        const size_t roff = (size_t)row * K;
        const IdxT *irow = idx + roff;
        const T *grow = g + roff * D;
        T *orow = gsrc + roff * D;
        const int n = K * D;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
            const int k = e / D, d = e - k * D;
            const long long j = (long long)irow[k];
            if (j < 0 || j >= K) continue;
            atomicAdd(orow + (size_t)j * D + d, grow[e]);
        }
    }
}

bool gather_bwd_rows_supported(const void *g, const void *idx, const void *gsrc, int64_t K, int64_t D);
int launch_gather_bwd_rows_f32(const float *g, const void *idx, int idx_is_i64, int64_t B, int64_t K, float *gsrc, cudaStream_t st);

template <typename T>
static int launch_gather_bwd_t(const T *g, const void *idx, int idx_is_i64, int64_t B, int64_t K, int64_t D, T *gsrc,
                               int sorted, cudaStream_t st, const char *name)
{
    const int threads = 256;
    unsigned gy = (unsigned)(B < 65535 ? B : 65535);
    if (sorted && sizeof(T) == 4 && gather_bwd_rows_supported(g, idx, gsrc, K, D)) // scalar latents: the row kernel
        return launch_gather_bwd_rows_f32(reinterpret_cast<const float *>(g), idx, idx_is_i64, B, K,
                                          reinterpret_cast<float *>(gsrc), st);
    if (sorted) {
        const uintptr_t align = reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(gsrc);
        const int vw = (D % 4 == 0 && sizeof(T) == 4 && (align & 15) == 0) ? 4 : ((D % 2 == 0 && (align & (2 * sizeof(T) - 1)) == 0) ? 2 : 1);
        const int dv = (int)(D / vw);
        const dim3 grid(gather_grid_x(K, dv == 1 ? 4 : 1, B), gy);
#define AESMC_BWD(VT, IT)                                                                                        \
        do {                                                                                                     \
            if (dv == 1) gather_bwd_sorted_kernel<VT, IT, 4, 1><<<grid, threads, 0, st>>>(reinterpret_cast<const VT *>(g), static_cast<const IT *>(idx), (int)B, (int)K, dv, reinterpret_cast<VT *>(gsrc)); \
            else gather_bwd_sorted_kernel<VT, IT, 1, 4><<<grid, threads, 0, st>>>(reinterpret_cast<const VT *>(g), static_cast<const IT *>(idx), (int)B, (int)K, dv, reinterpret_cast<VT *>(gsrc)); \
        } while (0)
        if (vw == 4) { if (idx_is_i64) AESMC_BWD(float4, int64_t); else AESMC_BWD(float4, int32_t); }
        else if (vw == 2) { if (idx_is_i64) AESMC_BWD(typename VecOf<T>::v2, int64_t); else AESMC_BWD(typename VecOf<T>::v2, int32_t); }
        else { if (idx_is_i64) AESMC_BWD(T, int64_t); else AESMC_BWD(T, int32_t); }
#undef AESMC_BWD
    } else {
        cudaError_t e = cudaMemsetAsync(gsrc, 0, sizeof(T) * (size_t)(B * K * D), st);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
        unsigned gx = (unsigned)((K * D + threads - 1) / threads);
        if (gx > 64) gx = 64;
        dim3 grid(gx ? gx : 1, gy);
        if (idx_is_i64) gather_bwd_atomic_kernel<T, int64_t><<<grid, threads, 0, st>>>(g, static_cast<const int64_t *>(idx), (int)B, (int)K, (int)D, gsrc);
        else gather_bwd_atomic_kernel<T, int32_t><<<grid, threads, 0, st>>>(g, static_cast<const int32_t *>(idx), (int)B, (int)K, (int)D, gsrc);
    }
    count_launch();
    return check_launch(name);
}

int launch_gather_bwd_f32(const float *g, const void *idx, int idx_is_i64, int64_t B, int64_t K, int64_t D,
                          float *gsrc, int sorted, cudaStream_t st)
{
    return launch_gather_bwd_t<float>(g, idx, idx_is_i64, B, K, D, gsrc, sorted, st, "gather_bwd<float>");
}
int launch_gather_bwd_f64(const double *g, const void *idx, int idx_is_i64, int64_t B, int64_t K, int64_t D,
                          double *gsrc, int sorted, cudaStream_t st)
{
    return launch_gather_bwd_t<double>(g, idx, idx_is_i64, B, K, D, gsrc, sorted, st, "gather_bwd<double>");
}

__global__ void iota_index_kernel(int64_t n, int K, int32_t *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int32_t)(i % K);
}
__global__ void widen_kernel(const int32_t *__restrict__ in, int64_t *__restrict__ out, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int64_t)in[i];
}
__global__ void narrow_kernel(const int64_t *__restrict__ in, int32_t *__restrict__ out, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int32_t)in[i];
}

static unsigned flat_blocks(int64_t n)
{
    int64_t b = (n + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    return (unsigned)(b < 1 ? 1 : b);
}

int launch_compose_index(const int32_t *prev, const int32_t *cur, int64_t B, int64_t K, int32_t *out, cudaStream_t st)
{
    // out[b,k] = prev[b, cur[b,k]]: a gather of 4-byte particles (out-of-range entries of cur are clamped)
    launch_gather_t<uint32_t, int32_t>(prev, cur, B, K, 1, out, nullptr, st);
    count_launch();
    return check_launch("compose_index");
}
int launch_iota_index(int64_t B, int64_t K, int32_t *out, cudaStream_t st)
{
    iota_index_kernel<<<flat_blocks(B * K), 256, 0, st>>>(B * K, (int)K, out);
    count_launch();
    return check_launch("iota_index_kernel");
}
int launch_index_widen(const int32_t *in, int64_t *out, int64_t n, cudaStream_t st)
{
    widen_kernel<<<flat_blocks(n), 256, 0, st>>>(in, out, n);
    count_launch();
    return check_launch("widen_kernel");
}
int launch_index_narrow(const int64_t *in, int32_t *out, int64_t n, cudaStream_t st)
{
    narrow_kernel<<<flat_blocks(n), 256, 0, st>>>(in, out, n);
    count_launch();
    return check_launch("narrow_kernel");
}

} // namespace aesmc
