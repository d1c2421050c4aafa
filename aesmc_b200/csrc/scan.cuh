// scan.cuh -- shared-memory block scan used by the generic and the multi-CTA step kernels.
#pragma once
#include "common.cuh"

namespace aesmc {

// In-place inclusive scan of buf[0..K): warp w owns the contiguous segment [w*seg, (w+1)*seg) and
// sweeps it 32 elements at a time (conflict-free smem access, shuffle scan + carry).  The scan is
// local to each warp's segment; warp_tot[w] receives the segment total and the caller folds the
// totals of preceding warps in when it consumes the values.
template <typename T, typename Op>
__device__ __forceinline__ void segment_scan_inplace(T *buf, int K, int seg, T identity, Op op, T *warp_tot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int begin = min(warp * seg, K), end = min(begin + seg, K);
    T carry = identity;
    for (int base = begin; base < end; base += 32) {
        const int k = base + lane;
        T v = (k < end) ? buf[k] : identity;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T n = __shfl_up_sync(kFull, v, o);
            if (lane >= o) v = op(n, v);
        }
        v = op(carry, v);
        if (k < end) buf[k] = v;
        carry = __shfl_sync(kFull, v, 31);
    }
    if (lane == 0) warp_tot[warp] = carry;
}

} // namespace aesmc
