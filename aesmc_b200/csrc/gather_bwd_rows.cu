// gather_bwd_rows.cu -- backward of the ancestral gather for SCALAR latents and the sorted ancestors systematic
// resampling produces (state.py:158-183 backward: torch.gather's scatter_add), one CTA per row, parent-centric:
//
//     g_src[b, j] = sum over the children k of parent j of g[b, k],     children of j = the run idx[b, k] == j
//
// The thread-per-particle kernel in gather.cu writes each run's sum with a scattered 4-byte store and zero-fills the
// childless parents in between (0.24 of the HBM roofline at B = K = 4096).  Here the row's gradients and ancestors are
// read with 16-byte loads, the gradients staged in shared memory, the first and the last child of every parent marked
// in a 16-bit pair, and then every LANE OWNS A PARENT: it sums its children in particle order -- exactly the order of
// the reference's CPU scatter_add -- for runs of up to 32 children; longer runs (collapsed weights) are queued and
// summed by a whole warp.  Children of consecutive parents are consecutive particles, so consecutive lanes read
// consecutive shared-memory words, and the results, zeros for childless parents included, leave as coalesced stores.  Deterministic, no atomics on the
// data, no memset.  12 bytes per particle: read g, read idx, write g_src.
#include "common.cuh"

namespace aesmc {

namespace {

constexpr int kLongRun = 32;   // children summed by the owning thread; longer runs go to a warp
constexpr int kMaxLong = 512;  // queue entries per row (a row of K particles has at most K / 33 long runs)

template <typename IdxT> struct Idx4 { IdxT v[4]; };
__device__ __forceinline__ void load_idx4(const int32_t *p, int c, int (&o)[4])
{
    const int4 v = __ldg(reinterpret_cast<const int4 *>(p) + c);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void load_idx4(const int64_t *p, int c, int (&o)[4])
{
    const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(p) + 2 * c), b = __ldg(reinterpret_cast<const longlong2 *>(p) + 2 * c + 1);
    o[0] = (int)a.x; o[1] = (int)a.y; o[2] = (int)b.x; o[3] = (int)b.y;
}

template <typename IdxT>
__global__ void __launch_bounds__(256) gather_bwd_rows_kernel(const float *__restrict__ g, const IdxT *__restrict__ idx, int B,
                                                              int K, float *__restrict__ gsrc)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sg = reinterpret_cast<float *>(smem_raw);                 // [K] the row's gradients
    unsigned *marks = reinterpret_cast<unsigned *>(sg + K);           // [K] (first child + 1) | (last child + 1) << 16
    __shared__ int qn;
    __shared__ int qj[kMaxLong], qs[kMaxLong], ql[kMaxLong];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nch = K >> 2;
    const unsigned marks_s = (unsigned)__cvta_generic_to_shared(marks);
    auto clampk = [K](int v) { return v < 0 ? 0 : (v >= K ? K - 1 : v); }; // memory safety only
    for (int row = blockIdx.x; row < B; row += gridDim.x) {
        const size_t roff = (size_t)row * K;
        const IdxT *irow = idx + roff;
        if (tid == 0) qn = 0;
        for (int c = tid; c < nch; c += blockDim.x) reinterpret_cast<uint4 *>(marks)[c] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        // stage the gradients and mark the first and the last child of every parent; four chunks of four particles per
        // thread and pass, ALL loads of a pass issued before the first mark store (the volatile shared-memory stores
        // would otherwise fence each iteration's loads behind the previous iteration)
        for (int c0 = tid; c0 < nch; c0 += 4 * blockDim.x) {
            float4 gv[4];
            int id[4][6];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c0 + q * blockDim.x;
                if (c < nch) {
                    gv[q] = __ldg(reinterpret_cast<const float4 *>(g + roff) + c);
                    int t[4];
                    load_idx4(irow, c, t);
                    id[q][1] = t[0]; id[q][2] = t[1]; id[q][3] = t[2]; id[q][4] = t[3];
                    id[q][0] = c ? (int)irow[4 * c - 1] : -1;
                    id[q][5] = c + 1 < nch ? (int)irow[4 * c + 4] : -1;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c0 + q * blockDim.x;
                if (c < nch) {
                    reinterpret_cast<float4 *>(sg)[c] = gv[q];
#pragma unroll
                    for (int p = 0; p < 6; ++p) id[q][p] = (p == 0 && c == 0) || (p == 5 && c + 1 == nch) ? -1 : clampk(id[q][p]);
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int k1 = 4 * c + p + 1; // particle + 1
                        if (id[q][p + 1] != id[q][p]) asm volatile("st.shared.u16 [%0], %1;" ::"r"(marks_s + 4u * id[q][p + 1]), "h"((unsigned short)k1) : "memory");
                        if (id[q][p + 1] != id[q][p + 2]) asm volatile("st.shared.u16 [%0], %1;" ::"r"(marks_s + 4u * id[q][p + 1] + 2u), "h"((unsigned short)k1) : "memory");
                    }
                }
            }
        }
        __syncthreads();
        // lane = parent: sum the children in particle order (the reference's scatter_add order).  Children of
        // consecutive parents are consecutive particles, so consecutive lanes read consecutive shared-memory words.
        float *orow = gsrc + roff;
        for (int j = tid; j < K; j += blockDim.x) {
            const unsigned m = marks[j];
            const int first = (int)(m & 0xffffu) - 1, last = (int)(m >> 16) - 1;
            float s = 0.f;
            if (first >= 0 && last >= first) {
                if (last - first < kLongRun) {
                    s = sg[first];
                    for (int k = first + 1; k <= last; ++k) s += sg[k];
                } else {
                    const int slot = atomicAdd(&qn, 1);
                    if (slot < kMaxLong) { qj[slot] = j; qs[slot] = first; ql[slot] = last; }
                }
            }
            orow[j] = s;
        }
        __syncthreads();
        const int nq = min(qn, kMaxLong);
        for (int e = warp; e < nq; e += blockDim.x >> 5) { // long runs: one warp each, fixed shuffle tree
            const int first = qs[e], last = ql[e];
            float s = 0.f;
            for (int k = first + lane; k <= last; k += 32) s += sg[k];
            s = warp_sum(s);
            if (lane == 0) orow[qj[e]] = s; // (ordered after this CTA's own zero store to the same address by the barrier)
        }
        __syncthreads(); // the next row rewrites sg / marks / the queue
    }
}

int sm_count_gb()
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

} // namespace

bool gather_bwd_rows_supported(const void *g, const void *idx, const void *gsrc, int64_t K, int64_t D)
{
    if (D != 1 || (K & 3) != 0 || K < 128 || K > 16384) return false; // 16-bit marks, row + marks in shared memory
    return ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(gsrc)) & 15) == 0;
}

int launch_gather_bwd_rows_f32(const float *g, const void *idx, int idx_is_i64, int64_t B, int64_t K, float *gsrc, cudaStream_t st)
{
    const size_t smem = (size_t)K * 8; // gradients | (first, last) child marks
    static int per_sm[2] = {0, 0};
    static size_t attr_for[2] = {0, 0};
    const int which = idx_is_i64 ? 1 : 0;
    const void *kern = idx_is_i64 ? (const void *)gather_bwd_rows_kernel<int64_t> : (const void *)gather_bwd_rows_kernel<int32_t>;
    static size_t occ_for[2] = {0, 0};
    if (smem > attr_for[which]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
        attr_for[which] = smem;
    }
    if (occ_for[which] != smem) { // (the occupancy query costs tens of microseconds of host time: once per shape)
        int n = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 256, smem);
        per_sm[which] = n < 1 ? 1 : n;
        occ_for[which] = smem;
    }
    long long grid = (long long)sm_count_gb() * per_sm[which];
    if (grid > B) grid = B;
    if (idx_is_i64) gather_bwd_rows_kernel<int64_t><<<(unsigned)grid, 256, smem, st>>>(g, static_cast<const int64_t *>(idx), (int)B, (int)K, gsrc);
    else gather_bwd_rows_kernel<int32_t><<<(unsigned)grid, 256, smem, st>>>(g, static_cast<const int32_t *>(idx), (int)B, (int)K, gsrc);
    count_launch();
    return check_launch("gather_bwd_rows_kernel");
}

} // namespace aesmc
