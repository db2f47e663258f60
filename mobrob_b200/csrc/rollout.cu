// Fused rollout (K4 + K1 + K6 + K11): replaces [SB3 2.0.0] OnPolicyAlgorithm.collect_rollouts over
// DummyVecEnv(Monitor(TimeLimit(PointEnv))) -- policy forward, Gaussian sampling, log-prob,
// action clip, env step, reward, goal-reached / time-limit flags, auto-reset, time-out
// bootstrap, Monitor episode statistics and RolloutBuffer.add -- for T steps in ONE launch.
// Reference call chain: examples/train.py:42-46 -> src/mobrob/rl_control/ppo.py:73-74 ->
// PPO.learn -> collect_rollouts -> EnvWrapper.step (src/mobrob/envs/wrapper.py:156-171).
//
// Environments are independent for the whole rollout and the parameters are frozen, so
// there is no grid-wide synchronisation: a warp owns 8 environments for all T steps.  The
// body state stays in fp64 registers of lanes 0..7, the parameters stay in shared memory,
// the two MLP towers run as the warp micro-GEMM of mlp.cuh, and the only HBM traffic is the
// rollout-buffer rows (80 B per env-step for the point robot).
#include "env_state.cuh"
#include "mlp.cuh"

namespace mr {

constexpr int RO_WARPS = 4;
constexpr int RO_E = 8;

struct RolloutArgs {
    PointState st;
    EnvCfg cfg;
    const float* params;
    int O;
    int64_t T, N;
    float* last_obs;       // [N][O]  in: obs the rollout starts from; out: obs it ended on
    float* last_starts;    // [N]     in/out: _last_episode_starts
    float* obs;            // [T][N][O]
    float* act;            // [T][N][2]  unclipped actions
    float* rew;            // [T][N]
    float* starts;         // [T][N]
    float* val;            // [T][N]
    float* logp;           // [T][N]
    float* last_val;       // [N]  V(final obs)
    uint8_t* last_done;    // [N]
    const float* eps;      // [T][N][2] host-supplied N(0,1) draws, or NULL -> Philox
    uint64_t seed;
    uint64_t noise_offset; // rollout counter * T (Philox counter word)
    int64_t env_offset;    // global index of env 0 (multi-GPU sharding)
    float gamma;
    double* ep_r;          // episode ring (Monitor): returns
    int32_t* ep_l;         //                          lengths
    unsigned long long* ep_count;  // total finished episodes (ring head)
    int ring_cap;
};

__global__ void __launch_bounds__(RO_WARPS * 32) point_rollout_kernel(RolloutArgs A) {
    extern __shared__ __align__(16) float smem[];
    const int O = A.O;
    SmemW W = stage_weights(smem, A.params, O);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* obsT = smem + smem_w_floats(O) + warp * (MAX_OBS * RO_E + 128 * RO_E);
    float* hbuf = obsT + MAX_OBS * RO_E;
    __syncthreads();

    const int64_t n0 = ((int64_t)blockIdx.x * RO_WARPS + warp) * RO_E;  // first env of this warp
    if (n0 >= A.N) return;
    const int n_live = (int)min((int64_t)RO_E, A.N - n0);
    const bool phys = lane < n_live;
    const int64_t n = n0 + lane;
    const float sig0 = expf(W.logstd[0]), sig1 = expf(W.logstd[1]);
    const float gamma = A.gamma;

    PointHot h;
    float start_flag = 0.f;
    if (phys) {
        h = A.st.load(n);
        start_flag = A.last_starts[n];
    }
    for (int idx = lane; idx < RO_E * O; idx += 32) {
        int e = idx / O, k = idx - e * O;
        obsT[k * RO_E + e] = e < n_live ? A.last_obs[n0 * O + idx] : 0.f;
    }
    __syncwarp();

    const int e_of = lane / 3, j_of = lane - 3 * e_of;
    bool done_flag = false;
    for (int64_t t = 0; t < A.T; ++t) {
        // RolloutBuffer.add(obs): the tile's 8 rows are contiguous in the [T][N][O] buffer
        {
            float* dst = A.obs + (t * A.N + n0) * O;
            for (int idx = lane; idx < n_live * O; idx += 32) {
                int e = idx / O, k = idx - e * O;
                dst[idx] = obsT[k * RO_E + e];
            }
        }
        float out = warp_mlp_forward<RO_E>(W, O, obsT, hbuf, lane);
        // sample: a = mu + sigma * eps ; log-prob in torch's evaluation order
        float a = out, lp = 0.f;
        if (lane < 3 * RO_E && e_of < n_live && j_of < 2) {
            const int64_t ne = n0 + e_of;
            float z;
            if (A.eps) {
                z = A.eps[(t * A.N + ne) * 2 + j_of];
            } else {
                const uint64_t g = (uint64_t)(A.env_offset + ne);
                const uint64_t c = A.noise_offset + (uint64_t)t;
                uint4 r = philox4x32(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)c, (uint32_t)(c >> 32)),
                                     make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32)));
                float2 zz = normal2(r.x, r.y);
                z = j_of == 0 ? zz.x : zz.y;
            }
            const float sig = j_of == 0 ? sig0 : sig1;
            a = __fadd_rn(out, __fmul_rn(sig, z));
            lp = normal_logprob(a, out, sig);
        }
        const int src = 3 * (lane & 7);
        const float a0 = __shfl_sync(0xffffffffu, a, src);
        const float a1 = __shfl_sync(0xffffffffu, a, src + 1);
        const float v = __shfl_sync(0xffffffffu, out, src + 2);
        const float logp = __fadd_rn(__shfl_sync(0xffffffffu, lp, src), __shfl_sync(0xffffffffu, lp, src + 1));

        StepResult r;
        r.trunc = false;
        float o_new[point::OBS], o_term[point::OBS];
        if (phys) {
            const int64_t row = t * A.N + n;
            reinterpret_cast<float2*>(A.act)[row] = make_float2(a0, a1);
            A.val[row] = v;
            A.logp[row] = logp;
            A.starts[row] = start_flag;
            r = point_env_step(h, A.st.cold, n, a0, a1, A.cfg, o_new, o_term);
            if (r.done) {
                unsigned long long slot = atomicAdd(A.ep_count, 1ull) % (unsigned long long)A.ring_cap;
                A.ep_r[slot] = r.ep_r;
                A.ep_l[slot] = r.ep_l;
            }
            start_flag = r.done ? 1.f : 0.f;
            done_flag = r.done;
        }
        // time-out bootstrap: reward += gamma * V(terminal_observation) where truncated
        const unsigned trunc_mask = __ballot_sync(0xffffffffu, phys && r.trunc);
        float rew = phys ? r.rew : 0.f;
        if (trunc_mask) {
            __syncwarp();
            if (phys) {
#pragma unroll
                for (int k = 0; k < point::OBS; ++k) obsT[k * RO_E + lane] = r.trunc ? o_term[k] : 0.f;
            }
            __syncwarp();
            const float tv_all = warp_mlp_forward<RO_E>(W, O, obsT, hbuf, lane);
            const float tv = __shfl_sync(0xffffffffu, tv_all, src + 2);
            if (phys && r.trunc) rew = __fadd_rn(rew, __fmul_rn(gamma, tv));
        }
        if (phys) {
            A.rew[t * A.N + n] = rew;
#pragma unroll
            for (int k = 0; k < point::OBS; ++k) obsT[k * RO_E + lane] = o_new[k];
        }
        __syncwarp();
    }

    // values of the final observation, carry-over state
    const float out = warp_mlp_forward<RO_E>(W, O, obsT, hbuf, lane);
    const float v_last = __shfl_sync(0xffffffffu, out, 3 * (lane & 7) + 2);
    if (phys) {
        A.last_val[n] = v_last;
        A.last_done[n] = done_flag ? 1 : 0;
        A.last_starts[n] = start_flag;
        A.st.store(n, h);
    }
    for (int idx = lane; idx < n_live * O; idx += 32) {
        int e = idx / O, k = idx - e * O;
        A.last_obs[n0 * O + idx] = obsT[k * RO_E + e];
    }
}

}  // namespace mr

using namespace mr;

extern "C" int mr_rollout(mr_env* env, const float* params, int64_t T, float* last_obs,
                          float* last_starts, float* obs, float* act, float* rew, float* starts,
                          float* val, float* logp, float* last_val, uint8_t* last_done,
                          const float* eps, uint64_t seed, uint64_t noise_offset,
                          int64_t env_offset, double gamma, double* ep_r, int32_t* ep_l,
                          unsigned long long* ep_count, int ring_cap, void* stream) {
    MR_REQUIRE(env && params && last_obs && last_starts && obs && act && rew && starts && val &&
                   logp && last_val && last_done && ep_r && ep_l && ep_count,
               "NULL argument");
    MR_REQUIRE(env->kind == MR_ENV_POINT, "fused rollout is built for the point env");
    MR_REQUIRE(T > 0 && ring_cap > 0, "T and ring_cap must be positive");
    RolloutArgs A;
    A.st = env->point;
    A.cfg = env->cfg;
    A.params = params;
    A.O = point::OBS;
    A.T = T;
    A.N = env->n;
    A.last_obs = last_obs; A.last_starts = last_starts;
    A.obs = obs; A.act = act; A.rew = rew; A.starts = starts; A.val = val; A.logp = logp;
    A.last_val = last_val; A.last_done = last_done;
    A.eps = eps; A.seed = seed; A.noise_offset = noise_offset; A.env_offset = env_offset;
    A.gamma = (float)gamma;
    A.ep_r = ep_r; A.ep_l = ep_l; A.ep_count = ep_count; A.ring_cap = ring_cap;
    const size_t smem = (smem_w_floats(A.O) + RO_WARPS * (MAX_OBS * RO_E + 128 * RO_E)) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        MR_CUDA(cudaFuncSetAttribute(point_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     100 * 1024));
        attr_set = true;
    }
    const int64_t warps = (env->n + RO_E - 1) / RO_E;
    const int blocks = (int)((warps + RO_WARPS - 1) / RO_WARPS);
    point_rollout_kernel<<<blocks, RO_WARPS * 32, smem, (cudaStream_t)stream>>>(A);
    MR_CHECK_LAUNCH();
    return MR_OK;
}
