// pairwise.cuh -- numpy's pairwise float32 summation order, reproduced on the device.
//
// scipy.special.logsumexp (math.py:22 of the reference) reduces exp(a - max) with np.sum, i.e. numpy's
// pairwise_sum (numpy/_core/src/umath/loops_utils.h.src): n <= 128 is a leaf summed with 8 strided
// accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail; larger n
// splits at n2 = n/2 - (n/2 % 8).  The recursion tree depends on K only: it is built once per CTA
// (BFS order, children adjacent), leaves are evaluated 8 lanes per leaf, levels are folded bottom-up.
#pragma once
#include "common.cuh"

namespace aesmc {

struct PwNode {
    int start, len, child; // child < 0: leaf; else children are nodes child, child + 1
    float val;
};

constexpr int kPairwiseMaxLevels = 40;

__host__ __device__ inline int pairwise_max_nodes(int K) { return 2 * (K / 56 + 2); }

// Padded row layout of the register-blocked kernel: one spare 16-byte chunk after every 8 chunks
// (4 floats per 32).  Striped float4 accesses (chunk = t + NT*i) and blocked ones (chunk = 4t + i)
// are both bank-conflict free, and addresses stay affine for unrolled loops.
__device__ __forceinline__ int pad_chunk(int c) { return c + (c >> 3); }
__device__ __forceinline__ int pad_elem(int k) { return k + ((k >> 5) << 2); }

static __device__ void build_pairwise_tree(PwNode *nodes, int *lvl_start, int *nlevels, int K)
{
    nodes[0].start = 0; nodes[0].len = K; nodes[0].child = -1; nodes[0].val = 0.f;
    int begin = 0, end = 1, L = 0;
    lvl_start[0] = 0;
    while (begin < end) {
        int cnt = end;
        for (int i = begin; i < end; ++i) {
            const int len = nodes[i].len, start = nodes[i].start;
            if (len > 128) {
                int n2 = len / 2;
                n2 -= n2 % 8;
                nodes[i].child = cnt;
                nodes[cnt].start = start; nodes[cnt].len = n2; nodes[cnt].child = -1; nodes[cnt].val = 0.f;
                ++cnt;
                nodes[cnt].start = start + n2; nodes[cnt].len = len - n2; nodes[cnt].child = -1; nodes[cnt].val = 0.f;
                ++cnt;
            }
        }
        begin = end;
        end = cnt;
        lvl_start[++L] = begin;
    }
    *nlevels = L;
}

// Sum of the K floats of a row in numpy's pairwise order.  PAD: the row is stored in the padded
// layout above.  All threads of the CTA call; the result is returned to all.
template <bool PAD>
static __device__ float pairwise_tree_sum(const float *buf, PwNode *nodes, const int *lvl_start, int nlevels)
{
    const int tid = threadIdx.x, NT = blockDim.x;
    const int nnodes = lvl_start[nlevels];
    const int grp = tid >> 3, j = tid & 7, ngrp = NT >> 3;
    auto at = [&](int k) { return buf[PAD ? pad_elem(k) : k]; };
    for (int base = 0; base < nnodes; base += ngrp) { // warp-uniform trip count
        const int n = base + grp;
        const bool valid = (n < nnodes) && (nodes[n].child < 0);
        int start = 0, len = 0, lim = 0;
        float r = 0.f;
        if (valid) {
            start = nodes[n].start;
            len = nodes[n].len;
            if (len >= 8) {
                lim = len - (len % 8);
                r = at(start + j);
                for (int i = 8; i < lim; i += 8) r = __fadd_rn(r, at(start + i + j));
            }
        }
        r = __fadd_rn(r, __shfl_xor_sync(kFull, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(kFull, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(kFull, r, 4));
        if (valid && j == 0) {
            float res;
            if (len < 8) {
                res = 0.f;
                for (int i = 0; i < len; ++i) res = __fadd_rn(res, at(start + i));
            } else {
                res = r;
                for (int i = lim; i < len; ++i) res = __fadd_rn(res, at(start + i));
            }
            nodes[n].val = res;
        }
    }
    __syncthreads();
    for (int L = nlevels - 2; L >= 0; --L) {
        for (int n = lvl_start[L] + tid; n < lvl_start[L + 1]; n += NT) {
            const int ch = nodes[n].child;
            if (ch >= 0) nodes[n].val = __fadd_rn(nodes[ch].val, nodes[ch + 1].val);
        }
        __syncthreads();
    }
    return nodes[0].val;
}

} // namespace aesmc
