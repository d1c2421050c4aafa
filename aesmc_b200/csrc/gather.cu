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

// dst[b,k,q] = src[b, idx[b,k], q] for q in [0, upr) units of type V per particle.
template <typename V, typename IdxT>
__global__ void gather_kernel(const V *__restrict__ src, const IdxT *__restrict__ idx, int B, int K, int upr,
                              V *__restrict__ dst, int32_t *flags)
{
    for (int row = blockIdx.y; row < B; row += gridDim.y) {
        const size_t roff = (size_t)row * K;
        const IdxT *irow = idx + roff;
        const V *srow = src + roff * upr;
        V *drow = dst + roff * upr;
        const int n = K * upr;
        if (upr == 1) {
            for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
                drow[k] = __ldg(srow + load_index(irow, k, K, flags));
        } else {
            for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
                const int k = e / upr, q = e - k * upr;
                drow[e] = __ldg(srow + (size_t)load_index(irow, k, K, flags) * upr + q);
            }
        }
    }
}

template <typename V, typename IdxT>
static void launch_gather_t(const void *src, const void *idx, int64_t B, int64_t K, int64_t upr, void *dst,
                            int32_t *flags, cudaStream_t st)
{
    const int threads = 256;
    const int64_t n = K * upr;
    unsigned gx = (unsigned)((n + threads - 1) / threads);
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    unsigned gy = (unsigned)(B < 65535 ? B : 65535);
    dim3 grid(gx, gy);
    gather_kernel<V, IdxT><<<grid, threads, 0, st>>>(static_cast<const V *>(src), static_cast<const IdxT *>(idx),
                                                      (int)B, (int)K, (int)upr, static_cast<V *>(dst), flags);
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

// Backward for non-decreasing indices (what the step kernel emits): the children of parent j form
// one contiguous run of k.  The thread that sees a run start walks the run and sums it in k order --
// the same order as the reference's CPU scatter_add -- so the result is deterministic and needs no
// atomics.  gsrc must be zero-initialised (parents without children).
template <typename T, typename IdxT>
__global__ void gather_bwd_sorted_kernel(const T *__restrict__ g, const IdxT *__restrict__ idx, int B, int K, int D,
                                         T *__restrict__ gsrc)
{
    for (int row = blockIdx.y; row < B; row += gridDim.y) {
        const size_t roff = (size_t)row * K;
        const IdxT *irow = idx + roff;
        const T *grow = g + roff * D;
        T *orow = gsrc + roff * D;
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
            const IdxT j = irow[k];
            if (k > 0 && irow[k - 1] == j) continue; // not a run start
            if (j < 0 || j >= K) continue;
            int end = k + 1;
            while (end < K && irow[end] == j) ++end;
            for (int d = 0; d < D; ++d) {
                T acc = grow[(size_t)k * D + d];
                for (int i = k + 1; i < end; ++i) acc += grow[(size_t)i * D + d];
                orow[(size_t)j * D + d] = acc;
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

template <typename T>
static int launch_gather_bwd_t(const T *g, const void *idx, int idx_is_i64, int64_t B, int64_t K, int64_t D, T *gsrc,
                               int sorted, cudaStream_t st, const char *name)
{
    cudaError_t e = cudaMemsetAsync(gsrc, 0, sizeof(T) * (size_t)(B * K * D), st);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
    const int threads = 256;
    unsigned gy = (unsigned)(B < 65535 ? B : 65535);
    if (sorted) {
        unsigned gx = (unsigned)((K + threads - 1) / threads);
        if (gx > 64) gx = 64;
        dim3 grid(gx ? gx : 1, gy);
        if (idx_is_i64) gather_bwd_sorted_kernel<T, int64_t><<<grid, threads, 0, st>>>(g, static_cast<const int64_t *>(idx), (int)B, (int)K, (int)D, gsrc);
        else gather_bwd_sorted_kernel<T, int32_t><<<grid, threads, 0, st>>>(g, static_cast<const int32_t *>(idx), (int)B, (int)K, (int)D, gsrc);
    } else {
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

__global__ void compose_index_kernel(const int32_t *__restrict__ prev, const int32_t *__restrict__ cur, int B, int K,
                                     int32_t *__restrict__ out)
{
    for (int row = blockIdx.y; row < B; row += gridDim.y) {
        const size_t roff = (size_t)row * K;
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
            int c = cur[roff + k];
            c = c < 0 ? 0 : (c >= K ? K - 1 : c);
            out[roff + k] = __ldg(prev + roff + c);
        }
    }
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
    unsigned gx = (unsigned)((K + 255) / 256);
    if (gx > 64) gx = 64;
    dim3 grid(gx ? gx : 1, (unsigned)(B < 65535 ? B : 65535));
    compose_index_kernel<<<grid, 256, 0, st>>>(prev, cur, (int)B, (int)K, out);
    count_launch();
    return check_launch("compose_index_kernel");
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
