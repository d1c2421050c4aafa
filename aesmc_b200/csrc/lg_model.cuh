// lg_model.cuh -- the scalar linear-Gaussian model evaluated inside the step kernels: Philox4x32-10
// + Box-Muller noise and torch.distributions.Normal.log_prob in torch's float32 arithmetic, two values per
// packed instruction.  Shared by smc_step_reg.cu and smc_step_x.cu.
#pragma once
#include "common.cuh"

namespace aesmc {

struct LgAffine { float mult, off, scale, two_var, log_scale; }; // loc = mult * x + off; 2 var and log scale precomputed by torch

// 5 consecutive floats of a parameter block in DEVICE memory (training: the parameters change every step and are
// never copied to the host; the block is tiny and stays in L1)
__device__ __forceinline__ LgAffine lg_load_affine(const float *q)
{
    LgAffine a;
    a.mult = __ldg(q); a.off = __ldg(q + 1); a.scale = __ldg(q + 2); a.two_var = __ldg(q + 3); a.log_scale = __ldg(q + 4);
    return a;
}
__device__ __forceinline__ bool lg_same(const LgAffine &a, const LgAffine &b)
{
    return a.mult == b.mult && a.off == b.off && a.scale == b.scale && a.two_var == b.two_var && a.log_scale == b.log_scale;
}

// Philox4x32-10 counter-based generator (Salmon et al. 2011): 4 x 32 random bits per (key, counter).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
// four standard normals from one Philox block (Box-Muller on two uniform pairs)
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long stream, unsigned long long index)
{
    const uint4 r = philox4x32_10(make_uint4((unsigned)index, (unsigned)(index >> 32), (unsigned)stream, (unsigned)(stream >> 32)),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const float u0 = ((float)(r.x >> 8) + 0.5f) * 5.9604644775390625e-08f, u1 = (float)(r.y >> 8) * 5.9604644775390625e-08f;
    const float u2 = ((float)(r.z >> 8) + 0.5f) * 5.9604644775390625e-08f, u3 = (float)(r.w >> 8) * 5.9604644775390625e-08f;
    const float m0 = sqrtf(-2.0f * __logf(u0)), m1 = sqrtf(-2.0f * __logf(u2));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u1, &s0, &c0);
    __sincosf(6.283185307179586f * u3, &s1, &c1);
    return make_float4(m0 * c0, m0 * s0, m1 * c1, m1 * s1);
}
// a * s with the product ROUNDED before anything else uses it: scalar mul.rn.f32 is never contracted,
// whereas ptxas merges mul.rn.f32x2 (and fma2(a, b, -0), which it folds back to a multiply) with a
// following add.rn.f32x2 into one FFMA2 -- torch rounds the product first
__device__ __forceinline__ f32x2 mul2_sep(f32x2 a, float s)
{
    float a0, a1;
    unpack2(a, a0, a1);
    return pack2(__fmul_rn(a0, s), __fmul_rn(a1, s));
}
// torch.distributions.Normal.log_prob in float32, operation for operation, two values at a time:
// -((value - loc) ** 2) / (2 * var) - log_scale - log(sqrt(2 pi)).  The IEEE division by the scalar 2*var
// uses the hoisted reciprocal + Markstein correction (identical to __fdiv_rn inside its safe range,
// __fdiv_rn itself outside).
__device__ __forceinline__ f32x2 normal_log_prob2(f32x2 value, f32x2 loc, float two_var, float rcp_tv, float log_scale, float c)
{
    const f32x2 d = sub2(value, loc);
    const f32x2 n = neg2(mul2(d, d));
    const f32x2 y = splat2(rcp_tv);
    const f32x2 q0 = mul2(n, y);
    f32x2 q = fma2(fma2(splat2(-two_var), q0, n), y, q0);
    float n0, n1;
    unpack2(n, n0, n1);
    const bool tv_ok = two_var > 9.3132257e-10f && two_var < 1.0737418e9f; // (2^-30, 2^30)
    const float a0 = fabsf(n0), a1 = fabsf(n1);
    if (!(tv_ok && fminf(a0, a1) >= 7.8886090522101181e-31f && fmaxf(a0, a1) <= 1.2676506e30f)) // outside [2^-100, 2^100]
        q = pack2(__fdiv_rn(n0, two_var), __fdiv_rn(n1, two_var));
    return sub2(sub2(q, splat2(log_scale)), splat2(c));
}


} // namespace aesmc
