// step_x.cuh -- launch interface of the second-generation exact row kernel (smc_step_x.cu)
#pragma once
#include "common.cuh"
#include "lg_model.cuh"

namespace aesmc {

struct XStepParams {
    const float *a, *b, *c;
    const double *u;
    int B;
    float *log_w, *lse;
    int32_t *idx;
    const float *x_in;
    float *x_out;
    int32_t *flags;
    float tol32;
    int debug_force;      // test hook (aesmc_debug_force_rare_paths): bit 0 every row fails the scan's verification (sequential redo),
                          // bit 1 every row raises the float64-comparison flag (marks redone by the general loop)
    // ---- fused scalar linear-Gaussian model (FUSED instances, aesmc_smc_step_lg_f32): the user model's sampling
    // and its three log-densities are evaluated in P1 instead of being read from HBM
    const float *x_prev;  // [B,K] resampled latents of the previous step, NULL at t = 0
    const float *y;       // [B] observation of this step
    const float *noise;   // [B,K] injected standard normals (tests), NULL -> Philox4x32-10
    const float *q_off;   // [B] per-row proposal offset (observation-dependent), NULL -> q.off
    float *x_new;         // [B,K] out, the newly proposed latents (nullable)
    LgAffine t, e, q;     // transition | initial, emission, proposal
    const float *params_dev; // non-NULL: the same 15 floats (t | e | q) are read from DEVICE memory instead
    float half_log_2pi;
    int q_same_t;         // proposal == transition (bootstrap): log q is the same number as log p(x | x_prev)
    unsigned long long seed, stream_offset;
    const unsigned long long *seed_dev; // non-NULL: the Philox key is read from device memory (CUDA-graph replays)
};

void smc_step_x_set_debug_force(int bits);
bool smc_step_x_supported(int64_t K, int mode, const void *idx, const void *x_in, int64_t D);
int launch_smc_step_x(const float *a, const float *b, const float *c, const double *u, int64_t B, int64_t K,
                      float *log_w, float *lse, int32_t *idx, const float *x_in, float *x_out, int32_t *flags,
                      cudaStream_t stream);
bool smc_step_x_lg_supported(int64_t K, int mode, const void *x_out, const void *u);
int launch_smc_step_x_lg(const XStepParams &proto, int64_t B, int64_t K, cudaStream_t stream);

} // namespace aesmc
