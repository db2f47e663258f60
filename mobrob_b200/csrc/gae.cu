// GAE (K7): RolloutBuffer.compute_returns_and_advantage [SB3 2.0.0; reached from
// src/mobrob/rl_control/ppo.py:73-74 via PPO.learn].  One thread per environment walks the
// time axis backwards; arrays are time-major [T][N] so every step is a coalesced row access.
//
// The arithmetic follows numpy's dtype flow through SB3's loop exactly (no FMA anywhere):
//   * `1.0 - dones` with a bool `dones` is float64, so the LAST step's delta is formed in
//     float64 and `last_gae_lam` is a float64 array from then on;
//   * for every earlier step delta = (r + ((g * nv) * nnt)) - v is float32 (python-float
//     gamma is cast to float32 by the float32 operand), the coefficient (gamma*lam) * nnt is
//     float32, and  A = delta + coeff * A_next  is evaluated in float64;
//   * advantages[t] = float32(A), returns = advantages + values in float32.
// Loads do not depend on the recurrence, so they are issued GAE_PF steps ahead.
#include "common.cuh"

namespace mr {

constexpr int GAE_PF = 8;

__global__ void __launch_bounds__(128)
gae_kernel(const float* __restrict__ rew, const float* __restrict__ val,
           const float* __restrict__ ep_start, const float* __restrict__ last_val,
           const uint8_t* __restrict__ last_done, float gamma, float gl,
           float* __restrict__ adv, float* __restrict__ ret, int64_t T, int64_t N) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float next_val = last_val[n];
    float next_nt = 0.f;
    double a = 0.0;
    bool last = true;
    float r_buf[GAE_PF], v_buf[GAE_PF], s_buf[GAE_PF];
    int64_t t = T - 1;
    while (t >= 0) {
        const int chunk = (int)min((int64_t)GAE_PF, t + 1);
#pragma unroll
        for (int k = 0; k < GAE_PF; ++k) {
            if (k < chunk) {
                const int64_t idx = (t - k) * N + n;
                r_buf[k] = rew[idx];
                v_buf[k] = val[idx];
                s_buf[k] = ep_start[idx];
            }
        }
#pragma unroll
        for (int k = 0; k < GAE_PF; ++k) {
            if (k < chunk) {
                const int64_t idx = (t - k) * N + n;
                const float gv = __fmul_rn(gamma, next_val);
                if (last) {
                    const double nnt = last_done[n] ? 0.0 : 1.0;
                    a = __dsub_rn(__dadd_rn((double)r_buf[k], __dmul_rn((double)gv, nnt)),
                                  (double)v_buf[k]);
                    last = false;
                } else {
                    const float delta = __fsub_rn(__fadd_rn(r_buf[k], __fmul_rn(gv, next_nt)), v_buf[k]);
                    const float coeff = __fmul_rn(gl, next_nt);
                    a = __dadd_rn((double)delta, __dmul_rn((double)coeff, a));
                }
                const float af = (float)a;
                adv[idx] = af;
                ret[idx] = __fadd_rn(af, v_buf[k]);
                next_val = v_buf[k];
                next_nt = __fsub_rn(1.0f, s_buf[k]);
            }
        }
        t -= chunk;
    }
}

}  // namespace mr

using namespace mr;

extern "C" int mr_gae(const float* rew, const float* val, const float* ep_start,
                      const float* last_val, const uint8_t* last_done, double gamma, double lam,
                      float* adv, float* ret, int64_t T, int64_t N, void* stream) {
    MR_REQUIRE(rew && val && ep_start && last_val && last_done && adv && ret, "NULL argument");
    if (T <= 0 || N <= 0) return MR_OK;
    // python-float scalars become float32 when they meet a float32 array
    const float g = (float)gamma;
    const float gl = (float)(gamma * lam);
    gae_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(rew, val, ep_start, last_val,
                                                                 last_done, g, gl, adv, ret, T, N);
    MR_CHECK_LAUNCH();
    return MR_OK;
}
