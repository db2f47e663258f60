// Per-environment step / reset logic of the point goal env, shared by the stand-alone
// VecEnv kernels (env.cu) and the fused rollout kernel (rollout.cu).
//
// Restates, per env: EnvWrapper.step/reward_fn/reached/reset (src/mobrob/envs/wrapper.py:
// 137-207), PointEnv.set_pos (wrapper.py:301-305), [GYM] TimeLimit (wrapper.py:568-569),
// [SB3] Monitor + DummyVecEnv auto-reset (src/mobrob/rl_control/ppo.py:37-48).
#pragma once

#include "point_dyn.cuh"

namespace mr {

struct EnvCfg {
    int time_limit;         // <= 0: no TimeLimit wrapper
    int terminate_on_goal;  // EnvWrapper(terminate_on_goal=...)
    point::K pk;            // integrator constants of the point robot (constant-bank operands)
};

// Arrays touched only by resets (and by the reference-view export).
struct EnvCold {
    uint64_t* pcg_init;    // [N][4] init_space stream
    uint64_t* pcg_goal;    // [N][4] goal_space stream
    int64_t* engine_seed;  // [N] Engine._seed
    float2* body_xy;       // [N] model.body_pos[robot][:2]
    double* psi0;          // [N] start heading (body quat)
    int32_t* counts;       // [N][2] (#resets, #full resets)
    const float* spaces;   // [8] init low (x, y), init high, goal low, goal high: EnvWrapper.init_space / goal_space
                           // (defaults wrapper.py:250-264; reset_init_space / reset_goal_space, wrapper.py:209-219)
};

struct PointHot {
    point::Dyn d;
    float cx, cz;  // data.ctrl (clipped action); survives goal-only resets
    float gx, gy;  // goal (float32, as drawn by Box.sample)
    int elapsed;   // TimeLimit._elapsed_steps == Monitor episode length
    double ep_ret; // Monitor episode return
};

struct StepResult {
    float rew;
    bool done, trunc, reach;
    double ep_r;
    int ep_l;
};

constexpr double REACH_RADIUS = 0.3;  // wrapper.py:203
constexpr double REACH_BONUS = 5.0;   // wrapper.py:151-152

__device__ __forceinline__ Pcg64 load_pcg(const uint64_t* p, int64_t i) {
    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p + 4 * i);
    ulonglong2 a = q[0], b = q[1];
    return Pcg64{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void store_pcg(uint64_t* p, int64_t i, const Pcg64& g) {
    // the increment never changes
    *reinterpret_cast<ulonglong2*>(p + 4 * i) = make_ulonglong2(g.hi, g.lo);
}

// EnvWrapper.reset for one env.  full: re-place the robot (first reset, or goal not reached).
__device__ inline void point_reset(PointHot& h, const EnvCold& cold, int64_t i, bool full) {
    if (full) {
        // Engine.reset() (wrapper.py:190) then PointEnv.set_pos -> Engine.reset() again
        // (wrapper.py:302): the heading that survives is the one drawn with _seed + 2.
        int64_t seed = cold.engine_seed[i] + 2;
        cold.engine_seed[i] = seed;
        Pcg64 g = load_pcg(cold.pcg_init, i);
        // init_space.sample(): default extents / 2 (wrapper.py:250-256); float32 bounds, float64 arithmetic
        float x = (float)g.uniform((double)cold.spaces[0], (double)cold.spaces[2]);
        float y = (float)g.uniform((double)cold.spaces[1], (double)cold.spaces[3]);
        store_pcg(cold.pcg_init, i, g);
        double heading = engine_heading((uint32_t)seed);
        h.d.px = (double)x;
        h.d.py = (double)y;
        h.d.psi = heading;
        h.d.vx = h.d.vy = h.d.om = 0.0;  // new MjSim: qpos = qvel = ctrl = 0
        h.cx = h.cz = 0.f;
        cold.body_xy[i] = make_float2(x, y);
        cold.psi0[i] = heading;
        cold.counts[2 * i + 1] += 1;
    }
    Pcg64 g = load_pcg(cold.pcg_goal, i);
    h.gx = (float)g.uniform((double)cold.spaces[4], (double)cold.spaces[6]);  // goal_space: default extents (wrapper.py:258-264)
    h.gy = (float)g.uniform((double)cold.spaces[5], (double)cold.spaces[7]);
    store_pcg(cold.pcg_goal, i, g);
    h.elapsed = 0;
    h.ep_ret = 0.0;
    cold.counts[2 * i] += 1;
}

// One VecEnv step of one env.  obs receives the row the VecEnv returns (post-reset when
// done); term_obs receives info["terminal_observation"] when done.
__device__ inline StepResult point_env_step(PointHot& h, const EnvCold& cold, int64_t i,
                                            float a0, float a1, const EnvCfg& cfg, float* obs,
                                            float* term_obs) {
    StepResult r;
    h.cx = fminf(fmaxf(a0, -1.f), 1.f);  // engine.py:1401-1405
    h.cz = fminf(fmaxf(a1, -1.f), 1.f);
    const double prevx = h.d.px, prevy = h.d.py;  // _prev_pos == position before the step
    double hc, hs;  // cos / sin of the heading after the step
    point::substeps(cfg.pk, h.d, (double)h.cx, (double)h.cz, hc, hs);
    const double gx = (double)h.gx, gy = (double)h.gy;
    const double dprev = point::dist2(gx, gy, prevx, prevy);
    const double dcur = point::dist2(gx, gy, h.d.px, h.d.py);
    r.reach = dcur < REACH_RADIUS;
    double reward = __dsub_rn(dprev, dcur);
    if (r.reach) reward = __dadd_rn(reward, REACH_BONUS);
    h.elapsed += 1;
    const bool term = r.reach && cfg.terminate_on_goal;
    const bool tl = cfg.time_limit > 0 && h.elapsed >= cfg.time_limit;
    r.done = term || tl;
    r.trunc = tl && !term;
    h.ep_ret = __dadd_rn(h.ep_ret, reward);
    r.rew = (float)reward;
    r.ep_r = h.ep_ret;
    r.ep_l = h.elapsed;
    point::sensors_cs(cfg.pk, h.d, hc, hs, (double)h.cx, (double)h.cz, h.gx, h.gy, obs, dcur);
    if (r.done) {
#pragma unroll
        for (int k = 0; k < point::OBS; ++k) term_obs[k] = obs[k];
        point_reset(h, cold, i, !r.reach);
        point::sensors(h.d, (double)h.cx, (double)h.cz, h.gx, h.gy, obs);
    }
    return r;
}

}  // namespace mr
