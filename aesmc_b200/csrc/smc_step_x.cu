// smc_step_x.cu -- second-generation EXACT-mode fused SMC step for the common row shapes
// (K = 16 * NT particles, NT = 64 ... 1024 threads, i.e. K = 1024, 2048, 4096, 8192, 16384; scalar latent).
//
// Same contract and the same bits as smc_step_reg.cu's exact instance -- log_w = (a + b) - c, scipy's
// logsumexp in numpy's pairwise order, np.cumsum's sequential float32 chain evaluated in parallel
// (see exact_scan.cuh for the theory), IEEE cdf / total, the float64 position comparison, the ancestral
// gather (inference.py:97-104,125-126,130,234-269; state.py:158-183) -- reorganised around what the
// round-1 profile showed (issue-bound at 49 % utilisation, 15 CTA barriers per row, 16.6 % of the stall
// samples behind the one-thread segment walker):
//
//   * the CTA size is a template parameter: K, the chunk strides and every padded shared-memory
//     address are compile-time constants (the generic kernel spends ~5 % of its instructions on them);
//   * HBM I/O is striped PER WARP (warp w owns particles [512 w, 512 w + 512), lane l moves chunks
//     l + 32 i of that span: still 512 contiguous bytes per instruction), so both striped <-> blocked
//     transposes stay inside the warp's own slice of the padded row buffer: __syncwarp instead of
//     __syncthreads, twice per row;
//   * warp boundaries are segment boundaries of the exact cumulative sum: inside a warp the parity maps of the
//     non-mixed blocks compose by shuffles (level 1); across warps the exact chain value is handed from warp w to
//     warp w + 1 through one tagged 8-byte shared-memory word (level 2: no walker warp, no barrier; a warp without
//     mixed blocks forwards the value through the map of its 32 blocks, only mixed blocks -- ~10 per row -- are
//     walked with real additions, on the registers of the lanes that own them);
//   * cross-warp prefixes (row max, pairwise partials, approximate prefix, max-scan carry) are
//     log-depth shuffles over one shared array instead of dependent loops over it;
//   * the verification flag of the exact scan rides on the next barrier instead of its own;
//   * 7 barriers per row instead of 15 (warp 0 only arrives at the fourth: it heads level 2's serial chain); no
//     barrier at the end of a row; the latent row is staged by one bulk asynchronous copy (cp.async.bulk +
//     mbarrier) into the weight buffer once the weights are in registers, and the next row's inputs are pulled
//     into L2 by cp.async.bulk.prefetch while this one is computed;
//   * the maxima that scipy excludes from the sum are found by one compare per thread (its own maximum
//     against the row's) instead of three instructions per particle; the IEEE-division range test and the
//     clamp of the boundary count are hoisted out of the per-particle loop (the CDF is monotone: the entry in
//     front of a block bounds the rest, and cdf / total <= 1 bounds the count by K); exp's denominator is
//     evaluated negated instead of being negated afterwards;
//   * the per-row loop state (row index, mbarrier phase) lives in shared memory, and the per-particle loops
//     contain neither calls nor exits: at 48 registers per thread (five CTAs per SM) anything live across a whole
//     row or across a call site is spilled, and a spill reload in this kernel is an L2 round trip.
//
// Everything it cannot take (other K, vector latents, fast mode, the fused-model step, no resampling)
// stays with smc_step_reg.cu / smc_step.cu / smc_step_large.cu.
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "pairwise.cuh"
#include "lg_model.cuh"
#include "step_x.cuh"

namespace aesmc {


namespace xk {

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// ---- one-dimensional bulk copy global -> shared, completion on an mbarrier (TMA engine, no tensor map) ----
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // earlier generic accesses to the buffer before the async write
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase)
{
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}"
                 ::"r"(b), "r"(phase) : "memory");
}
__device__ __forceinline__ float4 ld_stream(const float4 *p)
{
    float4 v;
#ifndef AESMC_X_LD
#define AESMC_X_LD 0
#endif
#if AESMC_X_LD == 0
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#elif AESMC_X_LD == 1
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#elif AESMC_X_LD == 2
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#elif AESMC_X_LD == 3
    asm volatile("ld.global.L1::no_allocate.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#else
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#endif
    return v;
}

// np_expf_nonpos_pair (common.cuh) with the denominator evaluated NEGATED: fma(-c1, r, -c0) and
// fma(., r, -1) are the exact negations of the reference's Horner steps (rounding is symmetric), which
// saves the two sign flips per pair; rcp.approx of the negated value is the negated reciprocal.
// NORMAL: the caller guarantees x >= -86.5, i.e. exponents >= -125 and normal results: no subnormal handling, no branch
template <bool NORMAL>
__device__ __forceinline__ void np_expf_nonpos_pair_x(float x0, float x1, float &r0, float &r1)
{
    const f32x2 x = pack2(x0, x1);
    const f32x2 magic = splat2(12582912.0f);
    const f32x2 tq = add2(pack2(__fmul_rn(x0, 1.44269504088896340736f), __fmul_rn(x1, 1.44269504088896340736f)), magic);
    const f32x2 q = sub2(tq, magic);
    f32x2 r = fma2(q, splat2(-6.93145752e-1f), x);
    r = fma2(q, splat2(-1.42860677e-6f), r);
    f32x2 num = fma2(splat2(5.082762527590693718096e-04f), r, splat2(6.757896990527504603057e-03f));
    num = fma2(num, r, splat2(5.114512081637298353406e-02f));
    num = fma2(num, r, splat2(2.473615434895520810817e-01f));
    num = fma2(num, r, splat2(7.257664613233124478488e-01f));
    num = fma2(num, r, splat2(9.999999999980870924916e-01f));
    f32x2 nden = fma2(splat2(-2.159509375685829852307e-02f), r, splat2(2.742335390411667452936e-01f));
    nden = fma2(nden, r, splat2(-1.0f));
    float n0, n1;
    unpack2(nden, n0, n1);
    f32x2 y = pack2(rcp_approx(-n0), rcp_approx(-n1)); // -nden IS den, bit for bit; the sign flip folds into the MUFU operand
    y = fma2(fma2(nden, y, splat2(1.0f)), y, y);
    const f32x2 q0 = mul2(num, y);
    const f32x2 poly = fma2(fma2(nden, q0, num), y, q0);
    float p0, p1, t0, t1;
    unpack2(poly, p0, p1);
    unpack2(tq, t0, t1);
    const int k0 = __float_as_int(t0) - 0x4B400000, k1 = __float_as_int(t1) - 0x4B400000;
    if (NORMAL || min(k0, k1) >= -125) {
        r0 = __int_as_float(__float_as_int(p0) + (k0 << 23));
        r1 = __int_as_float(__float_as_int(p1) + (k1 << 23));
        return;
    }
    r0 = (k0 >= -125) ? __int_as_float(__float_as_int(p0) + (k0 << 23))
       : (x0 <= -103.97208404541015625f) ? 0.0f
       : __fmul_rn(__int_as_float(__float_as_int(p0) + ((k0 + 64) << 23)), 5.42101086242752217e-20f);
    r1 = (k1 >= -125) ? __int_as_float(__float_as_int(p1) + (k1 << 23))
       : (x1 <= -103.97208404541015625f) ? 0.0f
       : __fmul_rn(__int_as_float(__float_as_int(p1) + ((k1 + 64) << 23)), 5.42101086242752217e-20f);
}

// explicit shared-window accesses for the serial walker: generic pointers make the compiler rebuild the shared base
// (S2R SR_CgaCtaId, LEA ...) inside the loop, and every instruction of that loop is on the row's critical path
__device__ __forceinline__ int4 lds_v4(unsigned addr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_f4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_b32(unsigned addr, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// run mark of a particle that owns the positions [cp, c): marks[pad_elem(cp)] = id iff c > cp, as ONE predicated store
// (an if around a store costs BSSY + BRA + BSYNC per particle; the alu pipe, which takes one warp-instruction every two
// cycles per scheduler, is what bounds this kernel); byte offset of pad_elem(cp): 4 cp + 16 (cp >> 5)
__device__ __forceinline__ void mark_run(unsigned marks_s, int cp, int c, int id)
{
    int hi; // cp >> 5, opaque to the compiler: (cp >> 5) << 4 would become (cp >> 1) & ~15 and cost a fourth instruction
    asm("shr.s32 %0, %1, 5;" : "=r"(hi) : "r"(cp));
    const unsigned addr = (marks_s + 4 * cp) + 16 * hi; // two shift-and-add instructions
    asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %1, %2;\n\t@p st.shared.b32 [%0], %3;\n\t}"
                 ::"r"(addr), "r"(c), "r"(cp), "r"(id) : "memory");
}


// the boundary count of common.cuh with the float64 uniform read from shared memory inside the rare fix-up
static __device__ __noinline__ int count_positions_below_slow_x(float cdf_entry, const double *u_sh, int K)
{
    const double Kd = (double)K;
    return count_positions_below(cdf_entry, *u_sh, K, Kd, Kd * 8.8817841970012523e-16);
}
// The boundary count for a CDF entry whose float32 closed form landed within tol32 of an integer (~0.1 % of the
// particles), without float64 arithmetic.  K is a power of two in this kernel, so T = cdfn K and the reference's
// pos_k = fl64(u + k) / K < cdfn  <=>  fl64(u + k) < T are exact scalings.  With T = Ti + Tf (integer and fraction,
// both exact in float32): every k < Ti counts, no k > Ti does, and k = Ti counts iff u < Tf -- unless float64's
// rounding of u + k interferes, which needs |u - Tf| or 1 - u below K 2^-52.  u = u_hi + u_lo (u_hi = fl32(u),
// u_lo = fl32(u - u_hi): 2^-50 absolute error), u_hi - Tf is exact whenever the two are close (Sterbenz), so the
// sign of d = (u_hi - Tf) + u_lo is the sign of u - Tf unless |d| is below a band of K 2^-46; only then (and for
// u_hi = 1) the reference's own float64 expression is evaluated (count_positions_below_slow_x).
__device__ __forceinline__ bool count_positions_near_x(float cdfn, float u_hi, const float *u_lo_sh, float Kf, int &count)
{
    const float T = __fmul_rn(cdfn, Kf);
    const float Ti = floorf(T);
    const float d = __fadd_rn(__fsub_rn(u_hi, __fsub_rn(T, Ti)), *u_lo_sh);
    count = (int)Ti + (d < 0.0f);
    return fabsf(d) > Kf * 1.4210854715202004e-14f && u_hi < 1.0f; // false: undecided in float32
}
__device__ __forceinline__ int count_positions_below_filtered_x(float cdf_entry, const double *u_sh, const float *u_lo_sh,
                                                                float u32, int K, float Kf, float tol32)
{
    const float tf = __fmaf_rn(cdf_entry, Kf, -u32);
    const float tm = __fadd_rn(tf, 12582912.0f);
    const float d = __fsub_rn(tf, __fsub_rn(tm, 12582912.0f)); // tf - rint(tf), exact
    if (fabsf(d) > tol32) return __float_as_int(__fadd_ru(tf, 12582912.0f)) - 0x4B400000; // ceil(tf) <= K
    int c;
    if (count_positions_near_x(cdf_entry, u32, u_lo_sh, Kf, c)) return c;
    return count_positions_below_slow_x(cdf_entry, u_sh, K);
}

#ifndef AESMC_X_CHAIN_SLEEP
#define AESMC_X_CHAIN_SLEEP 0 // > 0: nanoseconds of back-off between two polls of the level-2 hand-off
#endif
// the warp chain of level 2: one 8-byte shared-memory word per warp, (value, tag), written and read as a unit
#ifndef AESMC_X_CHAIN_SYNC
#define AESMC_X_CHAIN_SYNC 1 // 0: volatile accesses, 1: st.relaxed / ld.relaxed at CTA scope (the record carries its own payload: no
                             // other data is ordered by it, so no fence -- release semantics cost a MEMBAR.CTA per hand-off),
                             // 2: shared-memory atomics (the racecheck control), 3: st.release / ld.acquire
#endif
__device__ __forceinline__ void chain_publish(unsigned addr, int value, int tag)
{
    const unsigned long long rec = ((unsigned long long)(unsigned)tag << 32) | (unsigned)value;
#if AESMC_X_CHAIN_SYNC == 0
    asm volatile("st.volatile.shared.b64 [%0], %1;" ::"r"(addr), "l"(rec) : "memory");
#elif AESMC_X_CHAIN_SYNC == 1
    asm volatile("st.relaxed.cta.shared.b64 [%0], %1;" ::"r"(addr), "l"(rec) : "memory");
#elif AESMC_X_CHAIN_SYNC == 3
    asm volatile("st.release.cta.shared.b64 [%0], %1;" ::"r"(addr), "l"(rec) : "memory");
#else
    unsigned long long old;
    asm volatile("atom.shared.exch.b64 %0, [%1], %2;" : "=l"(old) : "r"(addr), "l"(rec) : "memory");
#endif
}
__device__ __forceinline__ int chain_wait(unsigned addr, int tag)
{
    unsigned long long rec;
    for (;;) {
#if AESMC_X_CHAIN_SYNC == 0
        asm volatile("ld.volatile.shared.b64 %0, [%1];" : "=l"(rec) : "r"(addr) : "memory");
#elif AESMC_X_CHAIN_SYNC == 1
        asm volatile("ld.relaxed.cta.shared.b64 %0, [%1];" : "=l"(rec) : "r"(addr) : "memory");
#elif AESMC_X_CHAIN_SYNC == 3
        asm volatile("ld.acquire.cta.shared.b64 %0, [%1];" : "=l"(rec) : "r"(addr) : "memory");
#else
        asm volatile("atom.shared.or.b64 %0, [%1], 0;" : "=l"(rec) : "r"(addr) : "memory");
#endif
        if ((int)(rec >> 32) == tag) break;
#if AESMC_X_CHAIN_SLEEP
        __nanosleep(AESMC_X_CHAIN_SLEEP);
#endif
    }
    return (int)(unsigned)rec;
}

// (prev then next): H[p] = P[p] + N[(p + P[p]) & 1]
__device__ __forceinline__ void compose(int p0, int p1, int n0, int n1, int &h0, int &h1)
{
    h0 = p0 + ((p0 & 1) ? n1 : n0);
    h1 = p1 + (((1 + p1) & 1) ? n1 : n0);
}

template <int NW> struct Shared {
    int cur_row, next_row;    // the row being processed / the next one (kept here, not in a register that is live --
                              // and spilled -- across the whole row); written after barrier (1), read after barrier (2)
    unsigned xphase;          // phase parity of the latent row's bulk-copy barrier
    double u64;               // this row's uniform (read by the rarest fix-up of the boundary count)
    float ulo;                // u64 - fl32(u64) (read by the rare float32 fix-up)
    float wmax[NW], part[NW], wsum[NW];
    int cnt[NW], i2[NW];
    float lse, total;
    int bad, fail, slow;
    int2 chain[NW + 1];       // (bits of the exact chain value entering warp w's span, row tag); [NW]: the row's total
};

template <int NW> __device__ __forceinline__ float across_max(const float *arr, int lane)
{
    float m = arr[lane & (NW - 1)];
#pragma unroll
    for (int o = NW >> 1; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    return m;
}

} // namespace xk

#ifndef AESMC_X_PREFETCH
#define AESMC_X_PREFETCH 1 // 1: bulk-prefetch the inputs of the next row this CTA will process into L2 (P1's wait for its loads
                           // drops from ~9.5 to ~6.7 thousand cycles of a ~30 thousand cycle row; 117.6 -> 116.2 us per launch)
#endif
#ifndef AESMC_X_REDUNDANT_TAIL
#define AESMC_X_REDUNDANT_TAIL 0 // 1: every warp evaluates the scalar lse tail itself, no barrier (3) (measured: -1 %)
#endif
#ifndef AESMC_X_BULK
#define AESMC_X_BULK 1 // 1: the latent row is staged by ONE bulk asynchronous copy (cp.async.bulk + mbarrier: the TMA
#endif                 // engine moves the 4 K bytes, no thread issues a load) instead of four 16-byte cp.async per thread
#ifndef AESMC_X_ARRIVE4
#define AESMC_X_ARRIVE4 1 // 1: warp 0 arrives at barrier (4) without waiting (see there)
#endif
#ifndef AESMC_X_TIMELINE
#define AESMC_X_TIMELINE 0 // 1 (profiling builds): per-warp clock stamps of the phases of CTA 0's rows, read by aesmc_debug_timeline
#endif
#if AESMC_X_TIMELINE
__device__ long long g_timeline[32 * 16]; // [warp][stage]: cycles since the row's start, summed over CTA 0's rows; [.][15] = rows
#define TL(stage) do { if (blockIdx.x == 0 && lane == 0) atomicAdd((unsigned long long *)&g_timeline[16 * warp + (stage)], (unsigned long long)(clock64() - tl0)); } while (0) /* (a reduction without a return value: the warp does not wait for it) */
#else
#define TL(stage) do { } while (0)
#endif
#ifndef AESMC_X_FORCE_GENERAL
#define AESMC_X_FORCE_GENERAL 0 // 1 (test builds): every row raises sh.slow and has its marks redone by the general loop
#endif
#ifndef AESMC_X_FORCE_FAIL
#define AESMC_X_FORCE_FAIL 0 // 1 (test builds): every row fails the scan's verification and takes the sequential redo path
#endif
#ifndef AESMC_X_ABLATE
#define AESMC_X_ABLATE 0 // timing experiments only (WRONG results): bit 0 no mixed-block walk, 1 no level-2 walk at all,
#endif                   // 2 no lse tail, 3 no float64 fix-up, 4 verification ignored

// ALIAS (plain step with a latent row to gather): the latent row is staged into the weight buffer once the exact
// scan is done with it -- 36 KB of shared memory per 256-thread CTA instead of 52, so five CTAs fit an SM, and the
// register budget is sized for them (48 registers; measured 151 us against 156 us for four CTAs at 64 registers).
// The fused-model step produces the latents itself in P1 and keeps the separate staging row (four CTAs per SM).
template <bool HAS_X, bool FUSED> struct XConfig {
    static constexpr bool kAlias = HAS_X && !FUSED;
#ifdef AESMC_X_THREADS_PER_SM
    static constexpr int kThreadsPerSM = AESMC_X_THREADS_PER_SM;
#else
    static constexpr int kThreadsPerSM = FUSED ? 1024 : 1280;
#endif
};

template <int NT, bool HAS_X, bool FUSED>
__global__ void __launch_bounds__(NT, (XConfig<HAS_X, FUSED>::kThreadsPerSM / NT) > 0 ? (XConfig<HAS_X, FUSED>::kThreadsPerSM / NT) : 1)
    smc_step_x_kernel(const XStepParams p)
{
    using namespace xk;
    constexpr bool AESMC_X_ALIAS_X = XConfig<HAS_X, FUSED>::kAlias;
    constexpr int NW = NT / 32, K = 16 * NT, NCH = 4 * NT;
    constexpr int ROWCH = NCH + NCH / 8; // padded row, in 16-byte chunks (one spare chunk per 8)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *bufW4 = reinterpret_cast<float4 *>(smem_raw);          // exp / weights (padded)
    int4 *bufM4 = reinterpret_cast<int4 *>(bufW4 + ROWCH);                                  // run marks; exact-scan segment lists
    float4 *bufX4 = AESMC_X_ALIAS_X ? bufW4 : reinterpret_cast<float4 *>(bufM4 + ROWCH); // staged latent row
    int *bufM = reinterpret_cast<int *>(bufM4);
    const float *bufX = reinterpret_cast<const float *>(bufX4);
    // [NW][32] records of the exact scan, written by every warp, read by the walker warp: (c0, mixed block's chunk + 1
    // or 0 | (c1 - c0 + 1) << 16) -- the two entries of a parity map differ by -1, 0 or 1 (two chains that start one
    // unit apart stay 0, 1 or 2 units apart: round-to-nearest shifts both alike except at ties)
    __shared__ Shared<NW> sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wbase = 144 * warp;                 // the warp's 128 chunks (+16 spare) of a padded row
    const int sl = wbase + lane + (lane >> 3);    // striped chunk lane + 32 i  ->  sl + 36 i
    const int bl = wbase + 4 * lane + (lane >> 1); // blocked chunk 4 lane + i  ->  bl + i
    const int gc = 128 * warp + lane;             // the same striped chunk in global memory (+ 32 i)
    const unsigned lt_mask = (1u << lane) - 1u;
    const float Kf = (float)K;

    constexpr bool kBulkX = AESMC_X_BULK && HAS_X && !FUSED && AESMC_X_ALIAS_X;
    constexpr bool kArrive4 = AESMC_X_ARRIVE4 && NT > 32 && (kBulkX || !(HAS_X && !FUSED && AESMC_X_ALIAS_X));
    __shared__ __align__(8) unsigned long long xbar; // completion of the latent row's bulk copy
    if (tid == 0) {
        sh.bad = 0;
        sh.slow = 0;
        sh.xphase = 0;
        for (int i = 0; i <= NW; ++i) sh.chain[i] = make_int2(0, 0); // (tags are row + 1: never 0)
        if (kBulkX) mbar_init(&xbar, 1);
    }
    __syncthreads();

    auto cur_off = [&]() { return (size_t)(*(volatile int *)&sh.cur_row) * K; }; // (from barrier (2) on)
    int row = blockIdx.x;
    while (row < p.B) {
        const size_t off = (size_t)row * K; // P1 only: later phases use cur_off()
#if AESMC_X_TIMELINE
        const long long tl0 = clock64();
        if (blockIdx.x == 0 && lane == 0) atomicAdd((unsigned long long *)&g_timeline[16 * warp + 15], 1ull);
#endif
        const float u32 = (float)p.u[row];
        if (tid == 0) { // (visible after barrier (1); the last readers are in front of the previous row's barrier (8))
            const double ud = p.u[row];
            sh.u64 = ud;
            sh.ulo = (float)(ud - (double)u32);
        }

        float4 lw[4];
        float tmax = -INFINITY, tmin = INFINITY; // (tmin: do all of this thread's weights stay normal numbers?)
        int bad = 0;
        if (FUSED) {
            // propose x ~ q(. | x_prev, y), then log_w = (log p(x | x_prev) + log p(y | x)) - log q(x | x_prev, y), each
            // term with torch.distributions.Normal's float32 arithmetic (lg_model.cuh)
            const float yv = p.y[row];
            const LgAffine mt = p.params_dev ? lg_load_affine(p.params_dev) : p.t;
            const LgAffine me = p.params_dev ? lg_load_affine(p.params_dev + 5) : p.e;
            const LgAffine mq = p.params_dev ? lg_load_affine(p.params_dev + 10) : p.q;
            const bool q_same_t = p.params_dev ? (p.q_off == nullptr && lg_same(mq, mt)) : (p.q_same_t != 0);
            const float qoff = p.q_off ? p.q_off[row] : mq.off;
            const unsigned long long seed = p.seed_dev ? *p.seed_dev : p.seed;
            const float rcp_t = refined_rcp(mt.two_var), rcp_e = refined_rcp(me.two_var), rcp_q = refined_rcp(mq.two_var);
            const float4 *__restrict__ xp4 = p.x_prev ? reinterpret_cast<const float4 *>(p.x_prev + off) + gc : nullptr;
            const float4 *__restrict__ nz4 = p.noise ? reinterpret_cast<const float4 *>(p.noise + off) + gc : nullptr;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 xp = xp4 ? ld_stream(xp4 + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 ep = nz4 ? ld_stream(nz4 + 32 * i)
                                      : philox_normal4(seed, p.stream_offset, (unsigned long long)off / 4 + gc + 32 * i);
                const f32x2 xs2[2] = {pack2(xp.x, xp.y), pack2(xp.z, xp.w)}, es2[2] = {pack2(ep.x, ep.y), pack2(ep.z, ep.w)};
                float xo[4], lo[4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const f32x2 loc_q = add2(mul2_sep(xs2[h], mq.mult), splat2(qoff));
                    const f32x2 x2 = add2(loc_q, mul2_sep(es2[h], mq.scale)); // Normal.rsample
                    const f32x2 lq = normal_log_prob2(x2, loc_q, mq.two_var, rcp_q, mq.log_scale, p.half_log_2pi);
                    const f32x2 lt = q_same_t ? lq
                                   : normal_log_prob2(x2, add2(mul2_sep(xs2[h], mt.mult), splat2(mt.off)),
                                                      mt.two_var, rcp_t, mt.log_scale, p.half_log_2pi);
                    const f32x2 le = normal_log_prob2(splat2(yv), add2(mul2_sep(x2, me.mult), splat2(me.off)),
                                                      me.two_var, rcp_e, me.log_scale, p.half_log_2pi);
                    unpack2(x2, xo[2 * h], xo[2 * h + 1]);
                    unpack2(sub2(add2(lt, le), lq), lo[2 * h], lo[2 * h + 1]);
                }
                const float4 xv = make_float4(xo[0], xo[1], xo[2], xo[3]);
                const float4 v = make_float4(lo[0], lo[1], lo[2], lo[3]);
                if (p.x_new) __stcs(reinterpret_cast<float4 *>(p.x_new + off) + gc + 32 * i, xv);
                if (p.log_w) __stcs(reinterpret_cast<float4 *>(p.log_w + off) + gc + 32 * i, v);
                bufX4[gc + 32 * i] = xv; // the gather source of P5 (the previous row's readers passed the barrier below)
                bad |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
                tmax = fmaxf(fmaxf(tmax, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
                tmin = fminf(fminf(tmin, fminf(v.x, v.y)), fminf(v.z, v.w));
                lw[i] = v;
            }
        } else {
            const float4 *__restrict__ a4 = reinterpret_cast<const float4 *>(p.a + off) + gc;
            const float4 *__restrict__ b4 = p.b ? reinterpret_cast<const float4 *>(p.b + off) + gc : nullptr;
            const float4 *__restrict__ c4 = p.c ? reinterpret_cast<const float4 *>(p.c + off) + gc : nullptr;
            float4 *__restrict__ o4 = reinterpret_cast<float4 *>(p.log_w + off) + gc;
            auto combine = [&](auto both) { // (a + b) - c; the three-operand case without per-load predicates
                constexpr bool kBoth = decltype(both)::value;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 v = ld_stream(a4 + 32 * i);
                    f32x2 lo = pack2(v.x, v.y), hi = pack2(v.z, v.w);
                    if (kBoth || b4) { const float4 t = ld_stream(b4 + 32 * i); lo = add2(lo, pack2(t.x, t.y)); hi = add2(hi, pack2(t.z, t.w)); }
                    if (kBoth || c4) { const float4 t = ld_stream(c4 + 32 * i); lo = sub2(lo, pack2(t.x, t.y)); hi = sub2(hi, pack2(t.z, t.w)); }
                    unpack2(lo, v.x, v.y);
                    unpack2(hi, v.z, v.w);
                    __stcs(o4 + 32 * i, v);
                    bad |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
                    tmax = fmaxf(fmaxf(tmax, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
                    tmin = fminf(fminf(tmin, fminf(v.x, v.y)), fminf(v.z, v.w));
                    lw[i] = v;
                }
            };
            if (b4 && c4) combine(std::true_type{}); else combine(std::false_type{});
        }
        {
            const float wm = warp_max(tmax);
            if (lane == 0) sh.wmax[warp] = wm;
            if (bad) sh.bad = 1;
        }
        TL(0);
        __syncthreads(); // (1) warp maxima; every warp is done with the previous row's staged latents
        TL(1);
        if (tid == 0) {
            sh.cur_row = row; sh.next_row = row + (int)gridDim.x; sh.fail = 0;
            sh.chain[0] = make_int2(0, row + 1); // the chain value entering the row, 0.0f (level 2; read after barrier (4))
        }
        if (HAS_X && !FUSED && !AESMC_X_ALIAS_X) { // stage this row's latents for the gather in P5
            const float4 *__restrict__ x4 = reinterpret_cast<const float4 *>(p.x_in + off) + gc;
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async_16(bufX4 + gc + 32 * i, x4 + 32 * i);
        }
        if (AESMC_X_PREFETCH) { // pull the next row this CTA will process into L2 while this one is being computed
            const int next = row + gridDim.x;
            const int pt = tid - 32 * (NW / 2); // (issued by a warp in the middle of the row: warp 0 heads level 2's chain)
            if (next < p.B && pt >= 0 && pt < 4) {
                const float *src = pt == 0 ? p.a : (pt == 1 ? p.b : (pt == 2 ? p.c : (HAS_X ? p.x_in : nullptr)));
                if (src) prefetch_l2_bulk(src + (size_t)next * K, (unsigned)K * 4u);
            }
        }
        const float vmax = across_max<NW>(sh.wmax, lane);
        if (sh.bad || !(fabsf(vmax) < INFINITY)) { // NaN / all -inf / +inf: flag, identity ancestors (CTA-uniform branch)
            const int isnan_row = sh.bad;
            if (tid == 0) {
                atomicOr(p.flags, isnan_row ? AESMC_FLAG_NAN : AESMC_FLAG_DEGENERATE);
                if (p.lse) p.lse[row] = isnan_row ? __int_as_float(0x7fc00000) : vmax;
            }
            for (int k = tid; k < K; k += NT) {
                if (p.idx) p.idx[off + k] = k;
                if (HAS_X) p.x_out[off + k] = FUSED ? bufX[k] : p.x_in[off + k];
            }
            if (!FUSED) cp_async_wait_all();
            __syncthreads();
            if (tid == 0) sh.bad = 0;
            __syncthreads();
            row += gridDim.x;
            continue;
        }

        // ---- P2a: e = np.exp(lw - max); numpy's pairwise sum of the non-maxima (scipy logsumexp) --
        {
            const bool has_max = (tmax == vmax); // this thread holds (one of) the row maxima: excluded and counted
            int cnt = 0;
            auto pass = [&](auto normal) { // one branch per thread and pass instead of one per pair (see np_expf_nonpos_pair_x)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = lw[i];
                    const float dx = __fsub_rn(v.x, vmax), dy = __fsub_rn(v.y, vmax), dz = __fsub_rn(v.z, vmax), dw = __fsub_rn(v.w, vmax);
                    float4 e;
                    np_expf_nonpos_pair_x<decltype(normal)::value>(dx, dy, e.x, e.y);
                    np_expf_nonpos_pair_x<decltype(normal)::value>(dz, dw, e.z, e.w);
                    if (has_max) {
                        if (dx == 0.0f) { e.x = 0.0f; ++cnt; }
                        if (dy == 0.0f) { e.y = 0.0f; ++cnt; }
                        if (dz == 0.0f) { e.z = 0.0f; ++cnt; }
                        if (dw == 0.0f) { e.w = 0.0f; ++cnt; }
                    }
                    bufW4[sl + 36 * i] = e;
                }
            };
            // (warp-uniform choice: a divergent warp would run both forms)
            if (__all_sync(kFull, __fsub_rn(tmin, vmax) >= -86.5f)) pass(std::true_type{}); else pass(std::false_type{});
            __syncwarp();
            // leaf L = particles [128 L, 128 L + 128) of the warp's span, summed by lanes 8L .. 8L+7 with numpy's
            // 8 strided accumulators; xor-shuffles 1, 2, 4 combine them as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
            // 8 and 16 fold the warp's four leaves as the balanced tree numpy's recursion is for K = 128 * 2^n
            const float *base = reinterpret_cast<const float *>(bufW4 + wbase) + 144 * (lane >> 3) + (lane & 7);
            float r = base[0];
#pragma unroll
            for (int i = 1; i < 16; ++i) r = __fadd_rn(r, base[8 * i + 4 * (i >> 2)]);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) r = __fadd_rn(r, __shfl_xor_sync(kFull, r, o));
            if (__any_sync(kFull, has_max)) cnt = warp_sum(cnt);
            if (lane == 0) { sh.part[warp] = r; sh.cnt[warp] = cnt; }
        }
        TL(2);
        __syncthreads(); // (2) per-warp partial sums and maxima counts
        TL(3);
        float lse;
        if (AESMC_X_REDUNDANT_TAIL || warp == NW - 1) { // fold the partials and evaluate the scalar tail (division, log1p, log)
            float s = sh.part[lane & (NW - 1)];
            int m = sh.cnt[lane & (NW - 1)];
#pragma unroll
            for (int o = 1; o < NW; o <<= 1) {
                s = __fadd_rn(s, __shfl_xor_sync(kFull, s, o));
                m += __shfl_xor_sync(kFull, m, o);
            }
            float v;
            if (AESMC_X_ABLATE & 4) {
                v = __fadd_rn(__fmul_rn(s, 0.01f), vmax);
            } else if (m == 1) { // a single maximum (the usual case): s / 1 = s and log(1) = +0 are exact no-ops
                v = __fadd_rn(fd_log1pf(s), vmax);
            } else {
                const float mf = (float)m;
                if (s != 0.0f) s = __fdiv_rn(s, mf);
                v = __fadd_rn(__fadd_rn(fd_log1pf(s), np_logf(mf)), vmax);
            }
            lse = v;
            if (tid == NT - 32) {
                if (!AESMC_X_REDUNDANT_TAIL) sh.lse = v;
                if (p.lse) p.lse[*(volatile int *)&sh.cur_row] = v;
            }
        }
        if (!AESMC_X_REDUNDANT_TAIL) {
            __syncthreads(); // (3) lse
            lse = sh.lse;
        }

        // ---- P2b: normalised weights np.exp(lw - lse) (math.py:49), striped -> blocked inside the warp --
        float w[16];
        {
            auto pass = [&](auto normal) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = lw[i];
                    float4 e;
                    np_expf_nonpos_pair_x<decltype(normal)::value>(__fsub_rn(v.x, lse), __fsub_rn(v.y, lse), e.x, e.y);
                    np_expf_nonpos_pair_x<decltype(normal)::value>(__fsub_rn(v.z, lse), __fsub_rn(v.w, lse), e.z, e.w);
                    bufW4[sl + 36 * i] = e;
                }
            };
            if (__all_sync(kFull, __fsub_rn(tmin, lse) >= -86.5f)) pass(std::true_type{}); else pass(std::false_type{});
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = bufW4[bl + i];
                w[4 * i + 0] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
        }

        // ---- P3: np.cumsum's sequential float32 chain, in parallel (exact_scan.cuh) -----------------
        // approximate prefix with an error bound -> classification of each thread's 16-particle block
        float ls = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) ls += w[j];
        float incl = ls;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float n = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) sh.wsum[warp] = incl;
#pragma unroll
        for (int i = 0; i < 4; ++i) bufM4[bl + i] = make_int4(0, 0, 0, 0); // run marks of P4 (the previous row's readers are past barrier (1))
        TL(4);
        // (4) warp totals of the approximate prefix; run marks cleared; every warp holds its weights in registers.
        // Warp 0 only ARRIVES: it needs nothing from the others here (its prefix offset is 0, it touches the marks and
        // the staged row again only after the row's total is known, which every other warp's pass of this barrier
        // precedes), and its walk of the row's first mixed blocks is the head of level 2's serial chain.
        if (kArrive4 && warp == 0) asm volatile("bar.arrive 0, %0;" ::"n"(NT) : "memory");
        else if (kArrive4) asm volatile("bar.sync 0, %0;" ::"n"(NT) : "memory");
        else __syncthreads();
        TL(5);
        if (HAS_X && !FUSED && AESMC_X_ALIAS_X) { // the weight buffer is free: stage this row's latents into it for the gather in P5
            if (kBulkX) { // one bulk copy (issued by one thread, completion on the mbarrier waited for in front of barrier (8))
                if (tid == NT - 32) bulk_copy_g2s(bufX4, p.x_in + cur_off(), (unsigned)K * 4u, &xbar);
            } else {
                const float4 *__restrict__ x4 = reinterpret_cast<const float4 *>(p.x_in + cur_off()) + gc;
#pragma unroll
                for (int i = 0; i < 4; ++i) cp_async_16(bufX4 + gc + 32 * i, x4 + 32 * i);
            }
        }
        float woff = 0.f;
        if (warp != 0) { // (warp 0 may be here before the others have stored their totals)
            float sc = sh.wsum[lane & (NW - 1)];
#pragma unroll
            for (int o = 1; o < NW; o <<= 1) {
                const float n = __shfl_up_sync(kFull, sc, o);
                if (lane >= o) sc += n;
            }
            woff = __shfl_sync(kFull, sc, (warp + 31) & 31); // inclusive sum of the warps before this one
        }
        const float p_in = woff + (incl - ls), p_out = woff + incl;
        // |chain - real prefix| <= k 2^-24 relative; the float scans above add < 64 further roundings
        const float eps = (float)(16 * (tid + 1) + 64) * 5.9604644775390625e-08f;
        const float lo = __fmul_rd(p_in, 1.0f - eps), hi = __fmul_ru(p_out, 1.0f + eps);
        int eb = 0; // biased exponent of a pure block's binade, 0 = mixed, -1 = absorbed (identity in every binade)
        if (lo >= 7.8886090522101181e-31f) {
            const int el = __float_as_int(lo) >> 23, eh = __float_as_int(hi) >> 23;
            if (el == eh) eb = el;
        }
        if (ls == 0.f || ls < __fmul_rd(lo, 1.4901161193847656e-08f)) eb = -1;
        {   // an absorbed block joins the pure run it follows inside the warp (its map is (0, 0))
            const unsigned live = __ballot_sync(kFull, eb != -1), below = live & lt_mask;
            const int e_prev = __shfl_sync(kFull, eb, below ? 31 - __clz(below) : 0);
            if (eb == -1 && below && e_prev > 0) eb = e_prev;
        }
        const float scale = __int_as_float((277 - (eb > 0 ? eb : 127)) << 23); // 2^(23 - e)
        int c0 = 0, c1 = 0;
        if (eb > 0) { // from here on a pure block holds its weights SCALED to units of the binade's ulp (exact: powers of two)
            f32x2 m01 = pack2(8388608.0f, 8388609.0f);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                w[j] = __fmul_rn(w[j], scale);
                m01 = add2(m01, splat2(w[j]));
            }
            float m0, m1;
            unpack2(m01, m0, m1);
            if (m1 < 16777216.0f) {
                c0 = __float_as_int(m0) & 0x7fffff;
                c1 = (__float_as_int(m1) & 0x7fffff) - 1;
            } else { // the block may leave the binade after all: mixed, on its unscaled weights again
                eb = 0;
                const float back = __int_as_float((254 << 23) - __float_as_int(scale)); // 1 / scale
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = __fmul_rn(w[j], back);
            }
        }
        // Level 1 (inside the warp): parity maps of consecutive non-mixed blocks compose (on the BIT PATTERN of the chain
        // value: for s = m 2^(e-23) in binade e, bits(s) + c[bits & 1] are the bits of (m + c) 2^(e-23) while m + c < 2^24
        // and exactly those of 2^(e+1) when m + c = 2^24, so runs may continue across a binade boundary that is hit
        // exactly); a mixed block cuts the run.  g = map of the run from its start through this block.
        const bool mixed = (eb == 0);
        const unsigned mixmask = __ballot_sync(kFull, mixed);
        const int kmix = __popc(mixmask & lt_mask); // mixed blocks of this warp before this one
        int g0 = c0, g1 = c1;                       // (0, 0) unless pure
        {
            const int dist = lane - (31 - __clz((mixmask | 1u) & ((2u << lane) - 1u))); // blocks since the run's first
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int q0 = __shfl_up_sync(kFull, g0, o), q1 = __shfl_up_sync(kFull, g1, o);
                if (o <= dist) compose(q0, q1, g0, g1, g0, g1);
            }
        }
        int prev0 = __shfl_up_sync(kFull, g0, 1), prev1 = __shfl_up_sync(kFull, g1, 1); // map of the run before this block
        if (lane == 0) prev0 = prev1 = 0;
        // Level 2, a chain over the warps of the CTA instead of a walker warp behind two barriers: warp w waits for the
        // exact chain value E_w entering its span (published by warp w - 1 in shared memory, tagged with the row),
        // pushes it through its own blocks and publishes E_{w+1}.  A warp without mixed blocks -- all but the first
        // one or two of a typical row -- forwards E_w through the map of its 32 blocks: three integer instructions
        // between the poll that sees E_w and the store of E_{w+1}.  Only mixed blocks are walked (16 real additions each,
        // every lane on its own registers, one shuffle per block).  seg = chain value after the last mixed block in
        // front of a lane (E_w if there is none); the replay below starts from it.
        TL(6);
        int seg;
        {
            const int tag = *(volatile int *)&sh.cur_row + 1;
            const unsigned chain_s = (unsigned)__cvta_generic_to_shared(sh.chain);
            // (every instruction between the poll that sees E_w and the store of E_{w+1} is on the row's critical path,
            // ~10 cycles each on a busy SM: the map is applied as (E + c0) + (E & 1) (c1 - c0), two levels deep; warp 0
            // polls like the others, its record (0, tag) was stored behind barrier (1); every lane stores the result)
            if (mixmask == 0u) {
                const int G0 = __shfl_sync(kFull, g0, 31), Gd = __shfl_sync(kFull, g1, 31) - G0;
                const int E = chain_wait(chain_s + 8u * warp, tag);
                chain_publish(chain_s + 8u * (warp + 1), (E + G0) + (E & 1) * Gd, tag);
                seg = E;
            } else {
                const int pd = prev1 - prev0;
                const int E = chain_wait(chain_s + 8u * warp, tag);
                int carry = E, t = 0;
                seg = E;
                for (unsigned m = mixmask; m; m &= m - 1) {
                    const int pl = __ffs(m) - 1;
                    float s = __int_as_float((carry + prev0) + (carry & 1) * pd);
#pragma unroll
                    for (int j = 0; j < 16; ++j) s = __fadd_rn(s, w[j]);
                    carry = __shfl_sync(kFull, __float_as_int(s), pl);
                    ++t;
                    if (kmix == t) seg = carry;
                }
                if (lane == 31) chain_publish(chain_s + 8u * (warp + 1), mixed ? carry : seg + ((seg & 1) ? g1 : g0), tag);
            }
        }
        TL(7);
        float s_in; // exact chain value entering this thread's block = the CDF entry of the particle before it
        {   // every thread replays its own block from its exact entry state
            int badv = 0;
            int sb = seg;
            sb += (sb & 1) ? prev1 : prev0;
            s_in = __int_as_float(sb);
            if (eb > 0) {
                int m = (sb & 0x7fffff) | 0x800000;
                if ((sb >> 23) != eb) { badv = 1; }
                float mf = __int_as_float(0x4B000000 | (m & 0x7fffff));
                const int unscale = (150 - eb) << 23;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    mf = __fadd_rn(mf, w[j]);
                    w[j] = __int_as_float(__float_as_int(mf) - unscale);
                }
                if (mf > 16777216.0f) badv = 1;
            } else {
                float s = s_in;
#pragma unroll
                for (int j = 0; j < 16; ++j) { s = __fadd_rn(s, w[j]); w[j] = s; }
            }
            if (AESMC_X_FORCE_FAIL || (p.debug_force & 1)) badv = 1;
            if (badv && !(AESMC_X_ABLATE & 16)) sh.fail = 1; // read after barrier (8)
        }
        // the row's total = the chain value leaving the last warp (every warp's marks of P4 were zeroed in front of
        // barrier (4))
        TL(8);
        float total = __int_as_float(chain_wait((unsigned)__cvta_generic_to_shared(sh.chain) + 8u * NW, *(volatile int *)&sh.cur_row + 1));
        TL(9);

        // ---- P4: closed-form offspring boundaries (inference.py:251,260-264) and run marks ---------------------
        // particle j owns the positions [c_{j-1}, c_j); the boundary of the particle in front of this thread's block
        // is recomputed from s_in (the same inputs, hence the same number, as its owner gets)
        bool force_general = false, wl_filled = false;
        volatile float wl[16]; // (volatile: the copy stays in the cold branch)
        for (int attempt = 0;; ++attempt) {
            float rcp = rcp_approx(total);
            rcp = __fmaf_rn(__fmaf_rn(-total, rcp, 1.0f), rcp, rcp);
            const bool tot_safe = total > 9.3132257e-10f && total < 2.0f;
            // the CDF is non-decreasing: the entry in front of the thread's block bounds the others from below, so one
            // test per thread decides whether the hoisted-reciprocal division is the IEEE quotient for all 17
            const bool safe = tot_safe && (tid ? s_in : w[0]) >= 7.8886090522101181e-31f;
            unsigned marks_s = (unsigned)__cvta_generic_to_shared(bufM);
            asm volatile("" : "+r"(marks_s)); // one register for the whole loop: do not rebuild the shared base per store
            const f32x2 rcp2 = splat2(rcp), ntot2 = splat2(-total), K2 = splat2(Kf), nu2 = splat2(-u32);
            const f32x2 magic = splat2(12582912.0f);
            const float tol32 = p.tol32;
            int idb = 16 * tid;
            asm volatile("" : "+r"(idb)); // (kept in a register: otherwise the thread index is re-read for every pair)
            // THE HOT LOOP HAS NO CALLS AND NO EXITS.  A call site inside it makes the 16 CDF entries live across a call,
            // and all but the callee-saved few are then spilled for every thread of every row (measured: 4.6 % of the
            // warp time waiting for those reloads).  A goto out of it (tried first) moves the reconvergence point of
            // every branch around it to its target: a lane in the 0.1 % fix-up branch then runs the rest of the loop
            // apart from its warp (measured on warp 0, whose tid != 0 test wrapped such an exit: P4 took twice as long).
            // So: a warp outside the range the hoisted reciprocal is the IEEE quotient in takes the rolled general loop
            // as a whole (warp-uniform), and a boundary that needs the reference's float64 comparison (within 2^-46 K
            // of an integer, ~1 in 10^7) only raises sh.slow: the row's marks are then redone by the general loop
            // behind barrier (8).
            int cp = 0;
            if (!force_general && __all_sync(kFull, safe)) { // one warp-uniform choice instead of a branch inside every pair
                {   // boundary of the particle in front of the block (0 for the row's first thread, whose s_in is 0)
                    const float q0 = __fmul_rn(s_in, rcp), n = __fmaf_rn(__fmaf_rn(-total, q0, s_in), rcp, q0);
                    const float tf = __fmaf_rn(n, Kf, -u32);
                    const float d = __fsub_rn(tf, __fsub_rn(__fadd_rn(tf, 12582912.0f), 12582912.0f));
                    cp = __float_as_int(__fadd_ru(tf, 12582912.0f)) - 0x4B400000;
                    if (!(fabsf(d) > tol32) && !count_positions_near_x(n, u32, &sh.ulo, Kf, cp) && tid != 0) sh.slow = 1;
                    if (tid == 0) cp = 0;
                }
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const f32x2 c2 = pack2(w[j], w[j + 1]);
                    const f32x2 q0 = mul2(c2, rcp2);
                    const f32x2 n2 = fma2(fma2(ntot2, q0, c2), rcp2, q0);
                    const f32x2 tf = fma2(n2, K2, nu2);         // cdfn * K - u, one rounding; <= K because cdfn <= 1
                    const f32x2 tm = add2(tf, magic);
                    const f32x2 d2 = sub2(tf, sub2(tm, magic)); // tf - rint(tf), exact
                    // ceil(tf): 1.5 * 2^23 + tf rounded TOWARDS +INF is exactly 1.5 * 2^23 + ceil(tf) (ulp 1 in that binade)
                    f32x2 tmu;
                    asm("add.rp.f32x2 %0, %1, %2;" : "=l"(tmu) : "l"(tf), "l"(magic));
                    float d0, d1, m0, m1;
                    unpack2(d2, d0, d1);
                    unpack2(tmu, m0, m1);
                    int ca = __float_as_int(m0) - 0x4B400000;
                    int cb = __float_as_int(m1) - 0x4B400000;
                    if ((AESMC_X_FORCE_GENERAL || (p.debug_force & 2)) && j == 6 && (tid & 63) == 1) sh.slow = 1;
                    if (!(fminf(fabsf(d0), fabsf(d1)) > tol32)) { // ~0.1 %: the exact comparison
                        float n0, n1;
                        unpack2(n2, n0, n1);
                        if (!(fabsf(d0) > tol32) && !count_positions_near_x(n0, u32, &sh.ulo, Kf, ca)) sh.slow = 1;
                        if (!(fabsf(d1) > tol32) && !count_positions_near_x(n1, u32, &sh.ulo, Kf, cb)) sh.slow = 1;
                    }
                    if (j == 14 && tid == NT - 1) cb = K; // last particle: positions up to 1.0 stay in range (Q5)
                    mark_run(marks_s, cp, ca, idb + j);
                    mark_run(marks_s, ca, cb, idb + j + 1);
                    cp = cb;
                }
            } else { // (typically the first warp of a row with collapsed weights; every warp in the redo of a "slow" row)
                // the CDF entries move to local memory in front of the calls and their registers are declared dead: live
                // across a call they would be spilled where they are produced, in the hot path, for every thread
                if (!wl_filled) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) wl[j] = w[j];
                    wl_filled = true;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = 0.f;
                int c0 = tid ? count_positions_below_filtered_x(__fdiv_rn(s_in, total), &sh.u64, &sh.ulo, u32, K, Kf, tol32) : 0;
#pragma unroll 1
                for (int j = 0; j < 16; ++j) {
                    int c = count_positions_below_filtered_x(__fdiv_rn(wl[j], total), &sh.u64, &sh.ulo, u32, K, Kf, tol32);
                    if (j == 15 && tid == NT - 1) c = K;
                    mark_run(marks_s, c0, c, idb + j);
                    c0 = c;
                }
            }
            if (kBulkX) {
                if (attempt == 0) mbar_wait(&xbar, *(volatile unsigned *)&sh.xphase);
            } else if (HAS_X && !FUSED) {
                cp_async_wait_all();
            }
            TL(10);
            __syncthreads(); // (8) run starts, staged latents and the scan's verdict visible
            TL(11);
            if (kBulkX && attempt == 0 && tid == 0) sh.xphase ^= 1u; // (read again in front of the next row's barrier (8))
            if (attempt == 0 && sh.fail) {
                // a binade bound was too optimistic (never observed): redo the row with the plain sequential chain
                __syncthreads();
                float4 *redo = AESMC_X_ALIAS_X ? reinterpret_cast<float4 *>(bufM4) : bufW4;
                if (tid == 0) {
                    float acc = 0.f;
                    for (int c = 0; c < NCH; ++c) {
                        float4 v;
                        if (AESMC_X_ALIAS_X) { // the weights are gone: recompute them from the log-weights this CTA stored
                            const float4 l = reinterpret_cast<const float4 *>(p.log_w + cur_off())[c];
                            const float lz = AESMC_X_REDUNDANT_TAIL ? lse : sh.lse;
                            v = make_float4(np_expf_nonpos(__fsub_rn(l.x, lz)), np_expf_nonpos(__fsub_rn(l.y, lz)),
                                            np_expf_nonpos(__fsub_rn(l.z, lz)), np_expf_nonpos(__fsub_rn(l.w, lz)));
                        } else {
                            v = bufW4[pad_chunk(c)];
                        }
                        v.x = acc = c ? __fadd_rn(acc, v.x) : v.x;
                        v.y = acc = __fadd_rn(acc, v.y);
                        v.z = acc = __fadd_rn(acc, v.z);
                        v.w = acc = __fadd_rn(acc, v.w);
                        redo[pad_chunk(c)] = v;
                    }
                    sh.total = acc;
                    sh.fail = 0;
                }
                __syncthreads();
                total = sh.total;
                s_in = tid ? reinterpret_cast<const float *>(redo)[pad_elem(16 * tid - 1)] : 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = redo[bl + i];
                    w[4 * i + 0] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
                }
                wl_filled = false;
                __syncthreads();
#pragma unroll
                for (int i = 0; i < 4; ++i) bufM4[bl + i] = make_int4(0, 0, 0, 0);
                __syncthreads();
                continue;
            }
            if (!force_general && sh.slow) { // (CTA-uniform) some boundary needs the float64 comparison: one more trip
                __syncthreads();               // through P4, every warp on the general loop
#pragma unroll
                for (int i = 0; i < 4; ++i) bufM4[bl + i] = make_int4(0, 0, 0, 0);
                if (tid == 0) sh.slow = 0;
                __syncthreads();
                force_general = true;
                continue;
            }
            break;
        }
        int id[16];
        {
            int run = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int4 v = bufM4[bl + i];
                id[4 * i + 0] = run = max(run, v.x);
                id[4 * i + 1] = run = max(run, v.y);
                id[4 * i + 2] = run = max(run, v.z);
                id[4 * i + 3] = run = max(run, v.w);
            }
            int inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc = max(inc, n);
            }
            if (lane == 31) sh.i2[warp] = inc;
            int excl = __shfl_up_sync(kFull, inc, 1);
            if (lane == 0) excl = 0;
            TL(12);
            __syncthreads(); // (9) warp maxima of the run ids
            {
                int sc = sh.i2[lane & (NW - 1)];
#pragma unroll
                for (int o = 1; o < NW; o <<= 1) {
                    const int n = __shfl_up_sync(kFull, sc, o);
                    if (lane >= o) sc = max(sc, n);
                }
                const int before = __shfl_sync(kFull, sc, (warp + 31) & 31);
                if (warp) excl = max(excl, before);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) id[j] = max(id[j], excl);
        }

        // ---- P5: indices out, fused ancestral gather from the staged row ---------------------------
        const size_t off5 = cur_off();
        row = *(volatile int *)&sh.next_row;
        if (!FUSED || p.idx != nullptr) { // (the fused-model step may resample without storing the ancestors)
            int4 *__restrict__ gidx4 = reinterpret_cast<int4 *>(p.idx + off5) + 4 * tid;
#pragma unroll
            for (int i = 0; i < 4; ++i) __stcs(gidx4 + i, make_int4(id[4 * i], id[4 * i + 1], id[4 * i + 2], id[4 * i + 3]));
        }
        if (HAS_X) {
            float4 *__restrict__ xo4 = reinterpret_cast<float4 *>(p.x_out + off5) + 4 * tid;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                __stcs(xo4 + i, make_float4(bufX[id[4 * i]], bufX[id[4 * i + 1]], bufX[id[4 * i + 2]], bufX[id[4 * i + 3]]));
        }
        // no barrier here: the next row touches the warp's own slice of bufW only after its barrier (1) ...
        // (bufX is restaged after barrier (1), bufM is rewritten after barrier (4)); the fused-model step writes the
        // staging row in its P1, hence:
        TL(13);
        if (FUSED) __syncthreads();
    }
}

static int g_debug_force = 0;
void smc_step_x_set_debug_force(int bits) { g_debug_force = bits; }

static int x_sm_count()
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

bool smc_step_x_supported(int64_t K, int mode, const void *idx, const void *x_in, int64_t D)
{
    static int disabled = -1;
    if (disabled < 0) {
        const char *e = getenv("AESMC_DISABLE_STEP_X");
        disabled = (e && atoi(e)) ? 1 : 0;
    }
    if (disabled || mode != AESMC_MODE_EXACT || idx == nullptr) return false;
    if (x_in != nullptr && D != 1) return false;
    return K == 1024 || K == 2048 || K == 4096 || K == 8192 || K == 16384;
}

template <int NT, bool HAS_X, bool FUSED>
static int launch_x(const XStepParams &p, int64_t B, cudaStream_t stream)
{
    constexpr int NCH = 4 * NT;
    constexpr size_t smem = (size_t)(NCH + NCH / 8) * 16 * 2 +
                            ((HAS_X && !XConfig<HAS_X, FUSED>::kAlias) ? (size_t)NCH * 16 : 0); // weights | run marks | [latents]
    auto kern = smc_step_x_kernel<NT, HAS_X, FUSED>;
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return AESMC_ERR_LAUNCH; }
        int n = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, NT, smem);
        per_sm = n < 1 ? 1 : n;
    }
    long long grid = (long long)x_sm_count() * per_sm;
    if (grid > B) grid = B;
    kern<<<(unsigned)grid, NT, smem, stream>>>(p);
    count_launch();
    return check_launch("smc_step_x_kernel");
}

template <bool HAS_X, bool FUSED>
static int dispatch_x(const XStepParams &p, int64_t B, int64_t K, cudaStream_t stream)
{
    switch (K) {
    case 1024: return launch_x<64, HAS_X, FUSED>(p, B, stream);
    case 2048: return launch_x<128, HAS_X, FUSED>(p, B, stream);
    case 4096: return launch_x<256, HAS_X, FUSED>(p, B, stream);
    case 8192: return launch_x<512, HAS_X, FUSED>(p, B, stream);
    case 16384: return launch_x<1024, HAS_X, FUSED>(p, B, stream);
    }
    set_error("smc_step_x: unsupported K=%lld", (long long)K);
    return AESMC_ERR_UNSUPPORTED;
}

int launch_smc_step_x(const float *a, const float *b, const float *c, const double *u, int64_t B, int64_t K,
                      float *log_w, float *lse, int32_t *idx, const float *x_in, float *x_out, int32_t *flags,
                      cudaStream_t stream)
{
    XStepParams p = {};
    p.a = a; p.b = b; p.c = c; p.u = u; p.B = (int)B; p.log_w = log_w; p.lse = lse; p.idx = idx;
    p.x_in = x_in; p.x_out = x_out; p.flags = flags;
    p.tol32 = (float)K * 1.1920928955078125e-07f + 5.9604644775390625e-08f; // K*2^-23 + 2^-24
    p.debug_force = g_debug_force;
    return x_in != nullptr ? dispatch_x<true, false>(p, B, K, stream) : dispatch_x<false, false>(p, B, K, stream);
}

// the fused scalar linear-Gaussian step (aesmc_smc_step_lg_f32) in exact mode with resampling: x_out required
bool smc_step_x_lg_supported(int64_t K, int mode, const void *x_out, const void *u)
{
    return smc_step_x_supported(K, mode, u /* resampling requested */, nullptr, 1) && x_out != nullptr;
}

int launch_smc_step_x_lg(const XStepParams &proto, int64_t B, int64_t K, cudaStream_t stream)
{
    XStepParams p = proto;
    p.B = (int)B;
    p.tol32 = (float)K * 1.1920928955078125e-07f + 5.9604644775390625e-08f;
    p.debug_force = g_debug_force;
    return dispatch_x<true, true>(p, B, K, stream);
}

} // namespace aesmc

#if AESMC_X_TIMELINE
extern "C" int aesmc_debug_timeline(long long *out_host, int reset) // profiling builds only (not declared in the public header)
{
    if (reset) {
        static long long zeros[32 * 16];
        return (int)cudaMemcpyToSymbol(aesmc::g_timeline, zeros, sizeof(zeros));
    }
    return (int)cudaMemcpyFromSymbol(out_host, aesmc::g_timeline, sizeof(long long) * 32 * 16);
}
#endif
