// HBM layout of the batched environment state: one slab per mr_env, struct-of-arrays,
// every array 256-byte aligned, env index fastest so a warp's loads are 128/256-byte
// coalesced.  "Hot" arrays are read+written by every step; "cold" arrays only by resets.
#pragma once

#include "env_car.cuh"
#include "env_point.cuh"

namespace mr {

constexpr int POINT_STATE_DIM = 15;

struct PointState {
    int64_t n;
    // hot: 6 x f64 + goal float2 + i32 + f64 = 68 B read; + ctrl float2 - goal = 68 B written per env-step
    double *px, *py, *psi, *vx, *vy, *om;
    float2* ctrl;
    float2* goal;
    int32_t* elapsed;
    double* ep_ret;
    EnvCold cold;

    static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }
    static size_t slab_bytes(int64_t n) {
        size_t b = 0;
        b += 7 * align_up(n * 8);   // px py psi vx vy om ep_ret
        b += 2 * align_up(n * 8);   // ctrl goal
        b += align_up(n * 4);       // elapsed
        b += 2 * align_up(n * 32);  // pcg_init pcg_goal
        b += align_up(n * 8);       // engine_seed
        b += align_up(n * 8);       // body_xy
        b += align_up(n * 8);       // psi0
        b += align_up(n * 8);       // counts
        return b;
    }
    void carve(void* slab, int64_t n_) {
        n = n_;
        char* p = static_cast<char*>(slab);
        auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes); return r; };
        px = (double*)take(n * 8); py = (double*)take(n * 8); psi = (double*)take(n * 8);
        vx = (double*)take(n * 8); vy = (double*)take(n * 8); om = (double*)take(n * 8);
        ep_ret = (double*)take(n * 8);
        ctrl = (float2*)take(n * 8); goal = (float2*)take(n * 8);
        elapsed = (int32_t*)take(n * 4);
        cold.pcg_init = (uint64_t*)take(n * 32); cold.pcg_goal = (uint64_t*)take(n * 32);
        cold.engine_seed = (int64_t*)take(n * 8);
        cold.body_xy = (float2*)take(n * 8);
        cold.psi0 = (double*)take(n * 8);
        cold.counts = (int32_t*)take(n * 8);
    }

    __device__ __forceinline__ PointHot load(int64_t i) const {
        PointHot h;
        h.d.px = px[i]; h.d.py = py[i]; h.d.psi = psi[i];
        h.d.vx = vx[i]; h.d.vy = vy[i]; h.d.om = om[i];
        float2 c = ctrl[i], g = goal[i];
        h.cx = c.x; h.cz = c.y; h.gx = g.x; h.gy = g.y;
        h.elapsed = elapsed[i];
        h.ep_ret = ep_ret[i];
        return h;
    }
    // env-step variants: data.ctrl is overwritten by the action before anything reads it, and the
    // goal only changes in a reset -- 8 B less read and 8 B less written per env-step.
    __device__ __forceinline__ PointHot load_step(int64_t i) const {
        PointHot h;
        h.d.px = px[i]; h.d.py = py[i]; h.d.psi = psi[i];
        h.d.vx = vx[i]; h.d.vy = vy[i]; h.d.om = om[i];
        const float2 g = goal[i];
        h.cx = 0.f; h.cz = 0.f; h.gx = g.x; h.gy = g.y;
        h.elapsed = elapsed[i];
        h.ep_ret = ep_ret[i];
        return h;
    }
    __device__ __forceinline__ void store_step(int64_t i, const PointHot& h, bool was_reset) const {
        px[i] = h.d.px; py[i] = h.d.py; psi[i] = h.d.psi;
        vx[i] = h.d.vx; vy[i] = h.d.vy; om[i] = h.d.om;
        ctrl[i] = make_float2(h.cx, h.cz);
        if (was_reset) goal[i] = make_float2(h.gx, h.gy);
        elapsed[i] = h.elapsed;
        ep_ret[i] = h.ep_ret;
    }
    __device__ __forceinline__ void store(int64_t i, const PointHot& h) const {
        px[i] = h.d.px; py[i] = h.d.py; psi[i] = h.d.psi;
        vx[i] = h.d.vx; vy[i] = h.d.vy; om[i] = h.d.om;
        ctrl[i] = make_float2(h.cx, h.cz);
        goal[i] = make_float2(h.gx, h.gy);
        elapsed[i] = h.elapsed;
        ep_ret[i] = h.ep_ret;
    }
};

}  // namespace mr

struct mr_env {
    int kind;
    int64_t n;
    int device;
    mr::EnvCfg cfg;
    void* slab;
    size_t slab_bytes;
    mr::PointState point;
    mr::CarSoA car;
    mr::car::Consts carK;
    void* scratch = nullptr;  // lazily allocated by mr_rollout_unfused
};
