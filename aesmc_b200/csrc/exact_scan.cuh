// exact_scan.cuh -- the reference's SEQUENTIAL float32 cumulative sum, computed in parallel, bit-exactly.
//
// np.cumsum over float32 weights (inference.py:257) is the chain s_k = RN(s_{k-1} + w_k).  Rounding
// makes it non-associative, so a parallel scan in float arithmetic flips ~1e-3 of the ancestor indices
// at K = 4096 (hard part 1).  The chain is nevertheless parallelisable exactly:
//
//   * while s stays inside one binade [2^e, 2^(e+1)) it is an integer multiple m*u of u = 2^(e-23),
//     and RN(s + w) = RN_even(m + w/u) * u: the rounding depends on s only through the PARITY of m
//     (ties), so a block of additions is the map m -> m + c[m & 1] for two integers (c0, c1), and such
//     maps compose associatively.  (c0, c1) are obtained by running the block's additions, scaled by
//     1/u, from the two representative starts 2^23 and 2^23 + 1: one FADD per particle per parity.
//   * weights are non-negative, so s is monotone and visits each binade once; an approximate prefix
//     (plain float scan) with a rigorous error bound tells, for each thread's block of 16 particles,
//     whether the whole block provably lies inside one binade ("pure") or may straddle a boundary
//     ("mixed", ~10 blocks per row).
//
//   * a block whose weights are all below half an ulp of the running sum ("absorbed": zero padding,
//     -inf log-weights, the plateaus of a collapsed weight vector) changes nothing: it is the identity
//     in every binade and joins the pure run it follows.
//
// Runs of pure blocks are composed with a segmented shuffle scan, one thread walks the ~20 run/mixed
// segments (O(1) per run, 16 real float additions per mixed block), and every thread then replays its
// own block from its exact entry state.  Every block re-verifies that its partial sums stayed inside
// the assumed binade; if any check fails the caller falls back to the plain sequential chain, so the
// result is always the reference's bits.
#pragma once
#include "common.cuh"
#include "pairwise.cuh"

namespace aesmc {

constexpr int kScanItems = 16;

struct ExactScanShared {
    int tail0[32], tail1[32], tailf[32], cnt[32];
    float wsum[32];
    int nseg, fail;
    float total;
    float carry_in; // chained rows: the exact value the chain entered this span with
};

// Where the chain's entry value comes from.
//   LocalCarry   : known up front (a whole row in one CTA, or spans streamed by one CTA).
//   chained spans: a row cut into spans owned by different CTAs (smc_step_large.cu).  The blocks are
//                  classified against an ESTIMATE of the entry value (anchor +- slack) while the
//                  previous span is still running; the segment walker then waits for the exact carry,
//                  checks that it lies inside the assumed band (otherwise the call fails with
//                  sh.fail == 2 and sh.carry_in set, and the caller redoes the span with a LocalCarry),
//                  and publishes the span's exit value as soon as it is known -- before the replay.
struct LocalCarry {
    static constexpr bool kChained = false;
    float s_in;
    __device__ __forceinline__ float anchor() const { return s_in; }
    __device__ __forceinline__ float slack() const { return 0.f; }
    __device__ __forceinline__ float wait(int) const { return s_in; }
    __device__ __forceinline__ void publish(float) const {}
    __device__ __forceinline__ void publish_map(int, int, int) const {}
};

// (prev then next): H[p] = P[p] + N[(p + P[p]) & 1]
__device__ __forceinline__ void compose_maps(int p0, int p1, int n0, int n1, int &h0, int &h1)
{
    h0 = p0 + ((p0 & 1) ? n1 : n0);
    h1 = p1 + (((1 + p1) & 1) ? n1 : n0);
}

// w: in, this thread's 16 weights (blocked layout, zeros beyond K); out, the reference's cumulative
// sums for the same particles.  bufW4: the weights in the padded row buffer (read by the segment
// walker for mixed blocks; left untouched).  scratch: >= 8*NT + 4 ints of shared memory, 16-byte
// aligned.  Returns
// true on success with *total = cumulative sum of the whole row; false if a verification failed (w
// is then unspecified and the caller recomputes the row with the sequential chain).
template <class Carry>
__device__ __forceinline__ bool exact_cumsum_blocked(float (&w)[kScanItems], float *total, const float4 *bufW4,
                                                     int *scratch, ExactScanShared &sh, const Carry carry)
{
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    int *recE = scratch;                 // biased exponent of a pure block's binade, 0 = mixed, -1 = all zero
    int *recG0 = scratch + NT;           // inclusive composed map of the run up to this block
    int *recG1 = scratch + 2 * NT;
    float *seg_state = reinterpret_cast<float *>(scratch + 3 * NT); // chain value at each segment start (NT+1)
    int4 *seg_rec = reinterpret_cast<int4 *>(scratch + 4 * NT + 4);  // per segment: (last block, binade, c0, c1)

    // ---- approximate prefix with an error bound ---------------------------------------------------
    float ls = 0.f;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) ls += w[j];
    float incl = ls;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) sh.wsum[warp] = incl;
    __syncthreads();
    float woff = 0.f;
    for (int v = 0; v < warp; ++v) woff += sh.wsum[v];
    const float s_anchor = carry.anchor();
    const float p_in = s_anchor + (woff + (incl - ls)), p_out = s_anchor + (woff + incl);
    // |chain - real prefix| <= k * 2^-24 relative (each RN adds <= 2^-24 of the running sum); the
    // float scan above adds < 64 further roundings.
    const float eps = (float)(kScanItems * (tid + 1) + 64) * 5.9604644775390625e-08f;
    float lo, hi;
    if constexpr (Carry::kChained) {
        lo = __fmul_rd(__fsub_rd(p_in, carry.slack()), 1.0f - eps);
        hi = __fmul_ru(__fadd_ru(p_out, carry.slack()), 1.0f + eps);
    } else {
        lo = __fmul_rd(p_in, 1.0f - eps);
        hi = __fmul_ru(p_out, 1.0f + eps);
    }
    int eb = 0;
    if (lo >= 7.8886090522101181e-31f) { // 2^-100: keeps 2^(23-e) representable
        const int el = __float_as_int(lo) >> 23, eh = __float_as_int(hi) >> 23;
        if (el == eh) eb = el;
    }
    // A block whose weights are all absorbed -- each below half an ulp of the running sum, so that every
    // RN(s + w) returns s -- is the identity map in every binade (eb = -1); zero weights (padding past K, -inf
    // log-weights) are the trivial case.  This is what keeps the plateaus of a collapsed weight vector out of
    // the walker: after the last heavy particle the sum sits within a few ulps of the binade boundary 1.0, and
    // every block there would otherwise be "mixed".  ls < lo * 2^-26 bounds every weight of the block by
    // 2^-25 of a lower bound of the entry value, i.e. strictly below half an ulp of anything the sum can be.
    if (ls == 0.f || ls < __fmul_rd(lo, 1.4901161193847656e-08f)) eb = -1;
    // ... and it joins the pure run it follows (within the warp) as a pure block of that binade -- its scaled
    // chain below yields the map (0, 0) -- instead of cutting the run in two: a collapsed row alternates
    // absorbed and live blocks and would otherwise hand the walker one segment per block
    {
        const unsigned live = __ballot_sync(kFull, eb != -1), below = live & ((1u << lane) - 1u);
        const int e_prev = __shfl_sync(kFull, eb, below ? 31 - __clz(below) : 0);
        if (eb == -1 && below && e_prev > 0) eb = e_prev;
    }

    // ---- pure block -> (c0, c1): the scaled chain from an even and an odd start -------------------
    const float scale = __int_as_float((277 - (eb > 0 ? eb : 127)) << 23); // 2^(23 - e), exact multiplier
    int c0 = 0, c1 = 0;
    if (eb > 0) {
        // (m0, m1) = (2^23, 2^23 + 1): ulp 1, so RN == round-half-even to integer; both parities advance
        // in one packed FADD2 per particle
        f32x2 m01 = pack2(8388608.0f, 8388609.0f);
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) m01 = add2(m01, splat2(__fmul_rn(w[j], scale)));
        float m0, m1;
        unpack2(m01, m0, m1);
        if (m1 < 16777216.0f) { // still in the ulp-1 regime (always true for a genuinely pure block)
            c0 = __float_as_int(m0) & 0x7fffff;
            c1 = (__float_as_int(m1) & 0x7fffff) - 1;
        } else {
            eb = 0;
        }
    }
    recE[tid] = eb;
    __syncthreads();

    // ---- segmented composition of the maps over runs of pure blocks -------------------------------
    const int eb_prev = tid ? recE[tid - 1] : 0;
    const int eb_next = (tid + 1 < NT) ? recE[tid + 1] : 0;
    const bool head = (eb == 0) || (eb_prev != eb);                         // first block of its segment
    const bool tail = (tid + 1 == NT) || (eb_next == 0) || (eb_next != eb); // last block of its segment
    int g0 = c0, g1 = c1, hf = head;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int p0 = __shfl_up_sync(kFull, g0, o), p1 = __shfl_up_sync(kFull, g1, o);
        const int pf = __shfl_up_sync(kFull, hf, o);
        if (lane >= o && !hf) {
            compose_maps(p0, p1, g0, g1, g0, g1);
            hf = pf;
        }
    }
    const unsigned endmask = __ballot_sync(kFull, tail);
    if (lane == 31) { sh.tail0[warp] = g0; sh.tail1[warp] = g1; sh.tailf[warp] = hf; sh.cnt[warp] = __popc(endmask); }
    __syncthreads();
    int k0 = 0, k1 = 0, segbase = 0; // carry map of the open segment entering this warp
    for (int v = 0; v < warp; ++v) {
        const int t0 = sh.tail0[v], t1 = sh.tail1[v];
        if (sh.tailf[v]) { k0 = t0; k1 = t1; }
        else compose_maps(k0, k1, t0, t1, k0, k1);
        segbase += sh.cnt[v];
    }
    if (!hf) compose_maps(k0, k1, g0, g1, g0, g1);
    recG0[tid] = g0;
    recG1[tid] = g1;
    const int segidx = segbase + __popc(endmask & ((1u << lane) - 1u));
    if (tail) seg_rec[segidx] = make_int4(tid, eb, g0, g1); // everything the walker needs, one load
    if (tid == NT - 1) sh.nseg = segidx + 1;
    __syncthreads();

    // ---- one thread walks the segments: O(1) per pure run, 16 float additions per mixed block -----
    // (lane 0 of the LAST warp: the warp scheduler favours the highest warp id among eligible warps, and
    // every other warp is about to wait at the barrier for this chain)
    if (tid >= NT - 32) {
      // chained spans: a span that is ONE pure (or absorbed) segment publishes its map before waiting for its carry,
      // and the whole warp looks back over the spans in front of it (decoupled look-back, smc_step_large.cu)
      if constexpr (Carry::kChained) {
          if (lane == 0 && sh.nseg == 1) {
              const int4 only = seg_rec[0];
              // (an all-absorbed span is the identity only above a threshold its consumer could not check: it waits)
              if (only.y > 0) carry.publish_map(only.y, only.z, only.w);
          }
      }
      float s = carry.wait(lane);
      if (lane == 0) {
        int fail = 0;
        int nseg = sh.nseg;
        if constexpr (Carry::kChained) {
            sh.carry_in = s;
            if (!(fabsf(__fsub_rn(s, s_anchor)) <= carry.slack())) { fail = 2; nseg = 0; }
        }
        seg_state[0] = s;
        int4 rec = seg_rec[0];
        for (int i = 0; i < nseg; ++i) {
            const int t = rec.x, e = rec.y, r0 = rec.z, r1 = rec.w;
            if (i + 1 < nseg) rec = seg_rec[i + 1]; // independent of the chain: overlaps with it
            if (e > 0) {
                const int sb = __float_as_int(s);
                if ((sb >> 23) != e) { fail = 1; break; }
                int m = (sb & 0x7fffff) | 0x800000;
                m += (m & 1) ? r1 : r0;
                if (m > 0x1000000) { fail = 1; break; }
                s = (m == 0x1000000) ? __int_as_float((e + 1) << 23) : __int_as_float((e << 23) | (m & 0x7fffff));
            } else if (e == 0) {
#pragma unroll
                for (int c = 0; c < kScanItems / 4; ++c) {
                    const float4 v = bufW4[pad_chunk(4 * t + c)];
                    s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w);
                }
            }
            seg_state[i + 1] = s;
        }
        if (!fail) carry.publish(s);
        sh.total = s;
        sh.fail = fail;
      }
    }
    __syncthreads();

    // ---- every thread replays its own block from its exact entry state ----------------------------
    int bad = sh.fail;
    const float s0 = seg_state[segidx];
    if (eb > 0) {
        const int sb = __float_as_int(s0);
        int m = (sb & 0x7fffff) | 0x800000;
        if ((sb >> 23) != eb) bad = 1;
        if (!head) m += (m & 1) ? recG1[tid - 1] : recG0[tid - 1];
        if (m >= 0x1000000) { bad = 1; m = 0x800000; }
        float mf = __int_as_float(0x4B000000 | (m & 0x7fffff)); // m as a float in [2^23, 2^24)
        const int unscale = (150 - eb) << 23;                   // multiply by u = 2^(e-23) on the bit pattern
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            mf = __fadd_rn(mf, __fmul_rn(w[j], scale));
            w[j] = __int_as_float(__float_as_int(mf) - unscale);
        }
        // The scaled chain IS the reference chain (power-of-two scaling commutes with rounding), so
        // these 16 values are exact whatever happens; but the (c0, c1) maps handed to the walker and
        // to later blocks of this run assumed the ulp-1 regime: landing exactly on 2^(e+1) is still
        // consistent (a following block of the same run rejects m_in = 2^24 above), going past is not.
        if (mf > 16777216.0f) bad = 1;
    } else {
        float s = s0;
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) { s = __fadd_rn(s, w[j]); w[j] = s; }
    }
    *total = sh.total;
    return __syncthreads_or(bad) == 0;
}

__device__ __forceinline__ bool exact_cumsum_blocked(float (&w)[kScanItems], float *total, const float4 *bufW4,
                                                     int *scratch, ExactScanShared &sh, float s_in = 0.f)
{
    return exact_cumsum_blocked(w, total, bufW4, scratch, sh, LocalCarry{s_in});
}

} // namespace aesmc
