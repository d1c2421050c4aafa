// row_gather.cuh -- ancestral gather of vector latents (state.py:158-183 for [B, K, D] values).
#pragma once
#include "common.cuh"
#include "pairwise.cuh"

namespace aesmc {

// ---- ancestral gather of vector latents: x_out[k, :] = x_in[idx[k], :], rows of D floats ----------------
// The (k, d) plane is swept flat in units of V floats (V = 4, 2 or 1, the widest that divides D), so stores
// are fully coalesced; ancestors are sorted, so the loads of a warp stay inside a few rows.  k = e / Dv
// uses a multiply-high by ceil(2^32 / Dv), exact for e < 2^32 / Dv (rows_gather_params checks it).
struct RowGather {
    int vec;       // V
    int dv;        // D / V
    unsigned mul;  // ceil(2^32 / dv); unused when dv == 1
};
inline RowGather rows_gather_params(long long K, long long D)
{
    RowGather g;
    g.vec = (D % 4 == 0) ? 4 : (D % 2 == 0) ? 2 : 1;
    g.dv = (int)(D / g.vec);
    g.mul = g.dv > 1 ? (unsigned)(((1ull << 32) + (unsigned long long)g.dv - 1) / (unsigned long long)g.dv) : 0u;
    if ((unsigned long long)K * (unsigned long long)g.dv * (unsigned long long)g.dv >= (1ull << 32)) g.mul = 0xffffffffu; // plain division
    return g;
}
template <typename T>
__device__ __forceinline__ void gather_rows_t(const float *__restrict__ xin, float *__restrict__ xout,
                                              const int *idx_padded, int n_rows, int dv, unsigned mul)
{
    const T *__restrict__ src = reinterpret_cast<const T *>(xin);
    T *__restrict__ dst = reinterpret_cast<T *>(xout);
    // loads in flight per thread: the callers run at their register limit, 16 bytes is what they can spare
    constexpr int kGatherUnroll = sizeof(T) == 16 ? 1 : 2;
    const int nv = n_rows * dv, tid = threadIdx.x, NT = blockDim.x;
#pragma unroll 1
    for (int e0 = tid; e0 < nv; e0 += kGatherUnroll * NT) {
        T v[kGatherUnroll];
#pragma unroll
        for (int q = 0; q < kGatherUnroll; ++q) {
            const int e = e0 + q * NT;
            if (e < nv) {
                const int k = dv == 1 ? e : (mul == 0xffffffffu ? e / dv : (int)__umulhi((unsigned)e, mul));
                v[q] = __ldg(src + (size_t)idx_padded[pad_elem(k)] * dv + (e - k * dv));
            }
        }
#pragma unroll
        for (int q = 0; q < kGatherUnroll; ++q) {
            const int e = e0 + q * NT;
            if (e < nv) __stcs(dst + e, v[q]);
        }
    }
}
// xin: the row block the indices refer to; xout: the first output row of this sweep; idx_padded: shared
// memory, ancestor of output row k at pad_elem(k) (k counted from the start of the sweep).
__device__ __forceinline__ void gather_rows(const float *xin, float *xout, const int *idx_padded, int n_rows,
                                                const RowGather g)
{
    if (g.vec == 4) gather_rows_t<float4>(xin, xout, idx_padded, n_rows, g.dv, g.mul);
    else if (g.vec == 2) gather_rows_t<float2>(xin, xout, idx_padded, n_rows, g.dv, g.mul);
    else gather_rows_t<float>(xin, xout, idx_padded, n_rows, g.dv, g.mul);
}

} // namespace aesmc
