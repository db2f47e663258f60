// MlpPolicy (SB3 ActorCriticPolicy, net_arch pi=[64,64] vf=[64,64], tanh) as a warp-level
// micro-GEMM: a warp pushes a tile of E samples through both towers at once; lane l owns
// hidden units {l, l+32} of the policy tower and {l, l+32} of the value tower, so one
// LDS.128 fetches the four weights a lane needs for one input feature, and the tile's
// activations are broadcast reads.  All fp32 FMA (parity with torch CPU to ~1e-6).
//
// Parameter vector = the 13 state-dict tensors of data/policies/*.zip, flattened in order.
#pragma once

#include "common.cuh"

namespace mr {

#ifndef MR_MLP_UNROLL1
#define MR_MLP_UNROLL1 2
#endif
#ifndef MR_MLP_UNROLL2
#define MR_MLP_UNROLL2 8
#endif
constexpr int MLP_UNROLL1 = MR_MLP_UNROLL1, MLP_UNROLL2 = MR_MLP_UNROLL2;   // k-loops of the two hidden layers
constexpr int HID = 64;
constexpr int ACT = 2;
constexpr int MAX_OBS = 32;

struct ParamLayout {
    int O;
    int logstd, pw1, pb1, pw2, pb2, vw1, vb1, vw2, vb2, aw, ab, cw, cb, total;
};

__host__ __device__ inline ParamLayout make_layout(int O) {
    ParamLayout L;
    L.O = O;
    int p = 0;
    L.logstd = p; p += ACT;
    L.pw1 = p; p += HID * O;
    L.pb1 = p; p += HID;
    L.pw2 = p; p += HID * HID;
    L.pb2 = p; p += HID;
    L.vw1 = p; p += HID * O;
    L.vb1 = p; p += HID;
    L.vw2 = p; p += HID * HID;
    L.vb2 = p; p += HID;
    L.aw = p; p += ACT * HID;
    L.ab = p; p += ACT;
    L.cw = p; p += HID;
    L.cb = p; p += 1;
    L.total = p;
    return L;
}

// Packed shared-memory image of the parameters (float offsets).
struct SmemW {
    const float4* w1p;  // [O][32]   {pi W1[l][k], pi W1[l+32][k], vf W1[l][k], vf W1[l+32][k]}
    const float4* w2p;  // [64][32]  same packing for the second layer
    const float4* b1p;  // [32]
    const float4* b2p;  // [32]
    const float* headw; // [3][HEAD_STRIDE]  action_net row 0, row 1, value_net row (rows 4 banks apart)
    const float* headb; // [4]       ab0 ab1 cb 0
    const float* logstd;// [4]
};

// head weight rows 68 floats apart: with 64 the three rows sat in the same banks and the head loop's LDS.128 (12 lanes,
// three rows) took 6 wavefronts each -- all of the rollout kernel's 40 M excess shared-memory wavefronts
constexpr int HEAD_STRIDE = 68;
__host__ __device__ inline int smem_w_floats(int O) { return 128 * O + 8192 + 128 + 128 + 208 + 4 + 4; }

// One weight matrix of both towers, [64][K] row-major in global memory -> the packed image dst[k][lane]{4}.
// A warp reads an 8-row x 4-column patch per pass (8 full sectors; the first form of this loop read one
// column of 32 rows, i.e. 32 sectors for 128 bytes, and the L1's sector rate -- not latency -- made staging
// 40 us of a 45 us policy_forward launch) and scatters it with 4-way bank conflicts at worst.
__device__ __forceinline__ void stage_matrix(float* dst, const float* __restrict__ pi_w, const float* __restrict__ vf_w, int K) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const int kb_n = (K + 3) >> 2;
    const int n_items = 2 * 8 * kb_n;   // (tower, block of 8 rows, block of 4 columns)
#pragma unroll 8
    for (int it = warp; it < n_items; it += n_warps) {
        const int kb = it % kb_n, ub = (it / kb_n) & 7, t = it / (8 * kb_n);
        const int u = 8 * ub + (lane >> 2), k = 4 * kb + (lane & 3);
        if (k < K) dst[k * 128 + (u & 31) * 4 + (u >> 5) + 2 * t] = __ldcg((t ? vf_w : pi_w) + u * K + k);
    }
}

// All threads of the block cooperate; caller must __syncthreads() afterwards.  Loads go through
// L2 (ld.cg): parameters may have been updated by a kernel that ran just before on another SM.
__device__ inline SmemW stage_weights(float* smem, const float* __restrict__ params, int O) {
    const ParamLayout L = make_layout(O);
    float* w1 = smem;
    float* w2 = w1 + 128 * O;
    float* b1 = w2 + 8192;
    float* b2 = b1 + 128;
    float* hw = b2 + 128;
    float* hb = hw + 208;
    float* ls = hb + 4;
    stage_matrix(w1, params + L.pw1, params + L.vw1, O);
    stage_matrix(w2, params + L.pw2, params + L.vw2, HID);
    for (int idx = threadIdx.x; idx < 128; idx += blockDim.x) {
        int lane = idx >> 2, j = idx & 3;
        int u = lane + ((j & 1) ? 32 : 0);
        b1[idx] = __ldcg(params + ((j & 2) ? L.vb1 : L.pb1) + u);
        b2[idx] = __ldcg(params + ((j & 2) ? L.vb2 : L.pb2) + u);
    }
    for (int idx = threadIdx.x; idx < 192; idx += blockDim.x)
        hw[(idx >> 6) * HEAD_STRIDE + (idx & 63)] = idx < 128 ? __ldcg(params + L.aw + idx) : __ldcg(params + L.cw + idx - 128);
    if (threadIdx.x < 4) {
        hb[threadIdx.x] = threadIdx.x < 2 ? __ldcg(params + L.ab + threadIdx.x)
                                          : (threadIdx.x == 2 ? __ldcg(params + L.cb) : 0.f);
        ls[threadIdx.x] = threadIdx.x < 2 ? __ldcg(params + L.logstd + threadIdx.x) : 0.f;
    }
    SmemW W;
    W.w1p = reinterpret_cast<const float4*>(w1);
    W.w2p = reinterpret_cast<const float4*>(w2);
    W.b1p = reinterpret_cast<const float4*>(b1);
    W.b2p = reinterpret_cast<const float4*>(b2);
    W.headw = hw;
    W.headb = hb;
    W.logstd = ls;
    return W;
}

// tanh for the hidden layers: (1 - t) / (1 + t), t = 2^(-2 log2(e) |x|), on the two MUFU units
// (ex2, rcp) -- 8 instructions against ~20 for tanhf, which spends the difference on RELATIVE
// accuracy near zero.  Absolute error <= 1.5e-7 over the whole range (tests/test_ppo_update_gpu.py
// holds the gradients to 1e-5 of torch's), which is what the activations need.
__device__ __forceinline__ float tanh_fast(float x) {
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(x) * -2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + t));
    return copysignf((1.f - t) * r, x);
}

// Warp-collective forward of E samples.
//   obsT  [O][E]       shared, features of the tile, sample index fastest
//   hbuf  [2*64][E]    shared scratch owned by this warp
// Returns, in lane 3*e + j (j = 0, 1: action mean; j = 2: value), the head output of sample e.
template <int E>
__device__ __forceinline__ float warp_mlp_forward(const SmemW& W, int O, const float* obsT,
                                                  float* hbuf, int lane) {
    static_assert(E % 4 == 0, "tile must be a multiple of 4 samples");
    float acc[4][E];
    {
        float4 b = W.b1p[lane];
#pragma unroll
        for (int e = 0; e < E; ++e) { acc[0][e] = b.x; acc[1][e] = b.y; acc[2][e] = b.z; acc[3][e] = b.w; }
    }
#pragma unroll MLP_UNROLL1
    for (int k = 0; k < O; ++k) {
        float4 w = W.w1p[k * 32 + lane];
        float x[E];
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
            float4 v = reinterpret_cast<const float4*>(obsT + k * E)[q];
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            acc[0][e] = fmaf(w.x, x[e], acc[0][e]);
            acc[1][e] = fmaf(w.y, x[e], acc[1][e]);
            acc[2][e] = fmaf(w.z, x[e], acc[2][e]);
            acc[3][e] = fmaf(w.w, x[e], acc[3][e]);
        }
    }
    auto store_tanh = [&]() {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float* row = hbuf + ((j >> 1) * 64 + lane + ((j & 1) ? 32 : 0)) * E;
#pragma unroll
            for (int q = 0; q < E / 4; ++q)
                reinterpret_cast<float4*>(row)[q] =
                    make_float4(tanh_fast(acc[j][4 * q]), tanh_fast(acc[j][4 * q + 1]),
                                tanh_fast(acc[j][4 * q + 2]), tanh_fast(acc[j][4 * q + 3]));
        }
    };
    store_tanh();
    __syncwarp();
    {
        float4 b = W.b2p[lane];
#pragma unroll
        for (int e = 0; e < E; ++e) { acc[0][e] = b.x; acc[1][e] = b.y; acc[2][e] = b.z; acc[3][e] = b.w; }
    }
#pragma unroll MLP_UNROLL2
    for (int k = 0; k < HID; ++k) {
        float4 w = W.w2p[k * 32 + lane];
        float hp[E], hv[E];
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
            float4 a = reinterpret_cast<const float4*>(hbuf + k * E)[q];
            float4 b = reinterpret_cast<const float4*>(hbuf + (64 + k) * E)[q];
            hp[4 * q] = a.x; hp[4 * q + 1] = a.y; hp[4 * q + 2] = a.z; hp[4 * q + 3] = a.w;
            hv[4 * q] = b.x; hv[4 * q + 1] = b.y; hv[4 * q + 2] = b.z; hv[4 * q + 3] = b.w;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            acc[0][e] = fmaf(w.x, hp[e], acc[0][e]);
            acc[1][e] = fmaf(w.y, hp[e], acc[1][e]);
            acc[2][e] = fmaf(w.z, hv[e], acc[2][e]);
            acc[3][e] = fmaf(w.w, hv[e], acc[3][e]);
        }
    }
    __syncwarp();
    store_tanh();
    __syncwarp();
    float out = 0.f;
    if (lane < 3 * E) {
        int e = lane / 3, j = lane - 3 * e;
        const float* hw = W.headw + j * HEAD_STRIDE;
        const float* h = hbuf + (j == 2 ? 64 * E : 0) + e;
        // eight partial sums: one 64-term chain of dependent FMAs (each behind two shared-memory loads) was 12 % of the
        // rollout kernel's stall samples
        float s[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = 0.f;
        s[0] = W.headb[j];
#pragma unroll
        for (int u = 0; u < HID; u += 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = fmaf(hw[u + i], h[(u + i) * E], s[i]);
        }
        out = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    }
    __syncwarp();
    return out;
}

// torch.distributions.Normal(mu, sigma).log_prob(a) for one action dimension, evaluated in
// fp32 in torch's order: -((a - mu)^2) / (2 sigma^2) - log(sigma) - log(sqrt(2 pi)).
__device__ __forceinline__ float normal_logprob(float a, float mu, float sigma) {
    float d = __fsub_rn(a, mu);
    float var = __fmul_rn(sigma, sigma);
    float q = __fdiv_rn(__fmul_rn(d, d), __fmul_rn(2.f, var));
    return __fsub_rn(__fsub_rn(-q, logf(sigma)), 0.9189385332046727f);
}

}  // namespace mr
