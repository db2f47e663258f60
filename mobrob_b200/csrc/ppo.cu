// PPO update (K8 + K9): replaces the inner loop of [SB3 2.0.0] PPO.train -- evaluate_actions,
// per-minibatch advantage normalisation, clipped surrogate, value MSE, entropy bonus,
// backward, clip_grad_norm_, Adam -- reached in the reference through
// src/mobrob/rl_control/ppo.py:73-74 (PPOCtrl.learn -> PPO.learn).  No autograd: the
// gradients of the two 64-64 tanh towers are formed analytically.
//
// ppo_grad_kernel: one CTA (8 warps) per 64-sample tile, persistent over tiles.
//   * forward and backward-data run warp-private on 8 samples per warp, lanes own hidden
//     units (same micro-GEMM as mlp.cuh), activations live in shared memory as
//     [tower][unit][sample] with a 68-float row stride (conflict-free for every access below);
//   * weight gradients are CTA-level register-tiled GEMMs over the tile's 64 samples
//     (dW2: 4x4x2 accumulators per thread, samples consumed four at a time with LDS.128);
//   * each CTA writes one partial gradient; ppo_reduce_kernel sums the partials in a fixed
//     order (deterministic), adam_kernel clips by global norm and applies torch's Adam.
#include "mlp.cuh"
#include "ppo_tc.cuh"

#include <stdlib.h>
#include <string.h>

namespace mr {

constexpr int PG_WARPS = 8;
constexpr int PG_THREADS = PG_WARPS * 32;
constexpr int PG_E = 8;                    // samples per warp
constexpr int PG_S = PG_WARPS * PG_E;      // 64 samples per tile
constexpr int PG_SP = PG_S + 4;            // padded row stride (floats)
constexpr int STAT_SLOTS = 16;             // tail of the gradient vector
// tail layout: 0 policy_loss, 1 value_loss, 2 clip_fraction, 3 approx_kl, 4 sample count

__host__ __device__ inline int grad_stride(int O) { return ((make_layout(O).total + 3) & ~3) + STAT_SLOTS; }
__host__ __device__ inline int stat_base(int O) { return (make_layout(O).total + 3) & ~3; }

struct GradArgs {
    const float* params;
    const float* obs;       // [T][N][O]
    const float* act;       // [T][N][2]
    const float* old_logp;  // [T][N]
    const float* adv;       // [T][N]
    const float* ret;       // [T][N]
    const int64_t* perm;    // [mb_size] env-major sample ids (n * T + t)
    const int32_t* rows;    // tensor-core kernels: the same samples as time-major buffer rows (t * N + n)
    int64_t mb_size;
    const double* mb_stats; // (sum adv, sum adv^2, count) of the GLOBAL minibatch
    int64_t N, T;
    float clip_range, ent_coef, vf_coef;
    int normalize_adv;
    float* partials;        // [gridDim.x][grad_stride]
};

struct GradSmem {
    SmemW W;
    float4* w2b;  // [u][lane] backward pack of W2
    float *X, *H1, *H2, *Z1, *dOut, *red;
};

// Carve the dynamic shared memory and (re)stage the parameters: all threads of the CTA cooperate;
// the caller must __syncthreads() before using them.
template <int O_PAD>
__device__ __forceinline__ GradSmem grad_stage(float* smem, const float* __restrict__ params, int O,
                                               bool first) {
    const int tid = threadIdx.x;
    GradSmem S;
    float* p = smem;
    float* wf = p;                                  p += smem_w_floats(O);
    S.w2b = reinterpret_cast<float4*>(p);           p += 8192;
    S.X = p;                                        p += O_PAD * PG_SP;
    S.H1 = p;                                       p += 128 * PG_SP;
    S.H2 = p;                                       p += 128 * PG_SP;   // becomes dZ2
    S.Z1 = p;                                       p += 128 * PG_SP;   // dZ1
    S.dOut = p;                                     p += PG_S * 4;
    S.red = p;                                      /* PG_WARPS * 32 * 16 floats */
    // the activation buffers are idle here: use them for the raw parameter image
    static_assert(2 * 128 * PG_SP >= 2 * HID * (MAX_OBS + 1) + 2 * HID * (HID + 1) + 1024, "raw image fits");
    stage_raw(S.H1, params, O);
    __syncthreads();
    S.W = pack_from_raw(wf, reinterpret_cast<float*>(S.w2b), S.H1, O);
    if (first)
        for (int idx = tid; idx < O_PAD * PG_SP; idx += PG_THREADS) S.X[idx] = 0.f;
    return S;
}

// One minibatch on this CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; writes the CTA's
// partial gradient (all grad_stride(O) - STAT_SLOTS + 4 used entries) to `out`.
template <int O_PAD>
__device__ __forceinline__ void grad_minibatch(const GradArgs& A, int O, const GradSmem& S,
                                               float* __restrict__ out) {
    const ParamLayout L = make_layout(O);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const SmemW& W = S.W;
    const float4* w2b = S.w2b;
    float *X = S.X, *H1 = S.H1, *H2 = S.H2, *Z1 = S.Z1, *dOut = S.dOut, *red = S.red;

    // ---- minibatch constants ----------------------------------------------------------------
    const double cnt = A.mb_stats[2];
    float adv_mean = 0.f, adv_std = 1.f;
    const bool do_norm = A.normalize_adv && cnt > 1.0;
    if (do_norm) {
        double m = A.mb_stats[0] / cnt;
        double var = (A.mb_stats[1] - A.mb_stats[0] * m) / (cnt - 1.0);
        adv_mean = (float)m;
        adv_std = (float)sqrt(fmax(var, 0.0));
    }
    const float inv_b = (float)(1.0 / cnt);
    const float sig0 = expf(W.logstd[0]), sig1 = expf(W.logstd[1]);

    // ---- persistent accumulators ---------------------------------------------------------------
    float gW2[2][4][4];   // [tower][u = tu + 16 i][k = tk + 16 j]
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) gW2[a][i][j] = 0.f;
    constexpr int KH = O_PAD / 2;
    float gW1[KH];        // tower = tid >> 7, unit = tid & 63, k in [half * KH, half * KH + KH)
#pragma unroll
    for (int k = 0; k < KH; ++k) gW1[k] = 0.f;
    float gb1[4] = {0, 0, 0, 0}, gb2[4] = {0, 0, 0, 0};  // lane-owned units, this warp's samples
    float gWh[6] = {0, 0, 0, 0, 0, 0};  // aW[0][l], aW[0][l+32], aW[1][l], aW[1][l+32], cW[l], cW[l+32]
    float g_head = 0.f;   // lane (e, j): d loss / d head bias j
    float g_ls = 0.f;     // lane (e, j<2): d loss / d log_std j (sample part)
    float st_pl = 0.f, st_vl = 0.f, st_cf = 0.f, st_kl = 0.f;

    const int tu = tid >> 4, tk = tid & 15;
    const int w1_tower = tid >> 7, w1_unit = tid & 63, w1_half = (tid >> 6) & 1;
    const int col0 = warp * PG_E;

    const int64_t n_tiles = (A.mb_size + PG_S - 1) / PG_S;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- gather: this warp's 8 samples ----------------------------------------------------
        const int64_t s_base = tile * PG_S + col0;
        int64_t my_row = -1;  // lane e < 8 holds the buffer row of sample e
        if (lane < PG_E && s_base + lane < A.mb_size) {
            int64_t id = A.perm[s_base + lane];
            int64_t n = id / A.T, t = id - n * A.T;
            my_row = t * A.N + n;
        }
        for (int it = 0; it < (PG_E * O + 31) / 32; ++it) {  // warp-uniform trip count (shuffles inside)
            const int idx = it * 32 + lane;
            const bool in = idx < PG_E * O;
            const int e = in ? idx / O : 0, k = idx - e * O;
            const int64_t row = __shfl_sync(0xffffffffu, my_row, e);
            if (in) X[k * PG_SP + col0 + e] = row >= 0 ? A.obs[row * O + k] : 0.f;
        }
        const int e_of = lane / 3, j_of = lane - 3 * e_of;
        const int64_t row_e = __shfl_sync(0xffffffffu, my_row, e_of < PG_E ? e_of : 0);
        const bool live = lane < 3 * PG_E && row_e >= 0;
        float a_j = 0.f, oldlp = 0.f, adv = 0.f, ret = 0.f;
        if (live) {
            if (j_of < 2) a_j = A.act[row_e * 2 + j_of];
            oldlp = A.old_logp[row_e];
            adv = A.adv[row_e];
            ret = A.ret[row_e];
        }
        __syncwarp();

        // ---- forward layer 1 -----------------------------------------------------------------------
        float acc[4][PG_E];
        {
            float4 b = W.b1p[lane];
#pragma unroll
            for (int e = 0; e < PG_E; ++e) { acc[0][e] = b.x; acc[1][e] = b.y; acc[2][e] = b.z; acc[3][e] = b.w; }
        }
        for (int k = 0; k < O; ++k) {
            float4 w = W.w1p[k * 32 + lane];
            float4 x0 = *reinterpret_cast<const float4*>(X + k * PG_SP + col0);
            float4 x1 = *reinterpret_cast<const float4*>(X + k * PG_SP + col0 + 4);
            float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int e = 0; e < PG_E; ++e) {
                acc[0][e] = fmaf(w.x, x[e], acc[0][e]);
                acc[1][e] = fmaf(w.y, x[e], acc[1][e]);
                acc[2][e] = fmaf(w.z, x[e], acc[2][e]);
                acc[3][e] = fmaf(w.w, x[e], acc[3][e]);
            }
        }
        auto row_of = [&](int j) { return (j >> 1) * 64 + lane + ((j & 1) ? 32 : 0); };
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float* dst = H1 + row_of(j) * PG_SP + col0;
            reinterpret_cast<float4*>(dst)[0] = make_float4(tanhf(acc[j][0]), tanhf(acc[j][1]), tanhf(acc[j][2]), tanhf(acc[j][3]));
            reinterpret_cast<float4*>(dst)[1] = make_float4(tanhf(acc[j][4]), tanhf(acc[j][5]), tanhf(acc[j][6]), tanhf(acc[j][7]));
        }
        __syncwarp();

        // ---- forward layer 2 -----------------------------------------------------------------------
        {
            float4 b = W.b2p[lane];
#pragma unroll
            for (int e = 0; e < PG_E; ++e) { acc[0][e] = b.x; acc[1][e] = b.y; acc[2][e] = b.z; acc[3][e] = b.w; }
        }
#pragma unroll 4
        for (int k = 0; k < HID; ++k) {
            float4 w = W.w2p[k * 32 + lane];
            const float* hp = H1 + k * PG_SP + col0;
            const float* hv = H1 + (64 + k) * PG_SP + col0;
            float4 p0 = reinterpret_cast<const float4*>(hp)[0], p1 = reinterpret_cast<const float4*>(hp)[1];
            float4 v0 = reinterpret_cast<const float4*>(hv)[0], v1 = reinterpret_cast<const float4*>(hv)[1];
            float a[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            float b[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int e = 0; e < PG_E; ++e) {
                acc[0][e] = fmaf(w.x, a[e], acc[0][e]);
                acc[1][e] = fmaf(w.y, a[e], acc[1][e]);
                acc[2][e] = fmaf(w.z, b[e], acc[2][e]);
                acc[3][e] = fmaf(w.w, b[e], acc[3][e]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < PG_E; ++e) acc[j][e] = tanhf(acc[j][e]);  // acc now holds h2
            float* dst = H2 + row_of(j) * PG_SP + col0;
            reinterpret_cast<float4*>(dst)[0] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            reinterpret_cast<float4*>(dst)[1] = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
        }
        __syncwarp();

        // ---- heads, loss, d loss / d head outputs (lane = 3 e + j) -----------------------------------
        float head = 0.f;
        if (lane < 3 * PG_E) {
            const float* hw = W.headw + j_of * 64;
            const float* h = H2 + (j_of == 2 ? 64 * PG_SP : 0) + col0 + e_of;
            float s = W.headb[j_of];
#pragma unroll 8
            for (int u = 0; u < HID; ++u) s = fmaf(hw[u], h[u * PG_SP], s);
            head = s;
        }
        float lp_j = 0.f, sig = 1.f, diff = 0.f;
        if (live && j_of < 2) {
            sig = j_of == 0 ? sig0 : sig1;
            lp_j = normal_logprob(a_j, head, sig);
            diff = __fsub_rn(a_j, head);
        }
        // logp = lp(j=0) + lp(j=1), broadcast to the sample's three lanes
        const int lane0 = 3 * e_of;
        float lp0 = __shfl_sync(0xffffffffu, lp_j, lane0 < 32 ? lane0 : 0);
        float lp1 = __shfl_sync(0xffffffffu, lp_j, lane0 + 1 < 32 ? lane0 + 1 : 0);
        float d_out = 0.f;
        if (live) {
            const float logp = __fadd_rn(lp0, lp1);
            const float log_ratio = logp - oldlp;
            const float ratio = expf(log_ratio);
            float adv_n = adv;
            if (do_norm) adv_n = __fdiv_rn(adv - adv_mean, adv_std + 1e-8f);
            const float lo = 1.f - A.clip_range, hi = 1.f + A.clip_range;
            const float pl1 = adv_n * ratio;
            const float pl2 = adv_n * fminf(fmaxf(ratio, lo), hi);
            const float g_logp = (pl1 <= pl2) ? -adv_n * ratio * inv_b : 0.f;
            if (j_of < 2) {
                const float inv_var = 1.f / (sig * sig);
                d_out = g_logp * diff * inv_var;                       // d/d mu_j
                g_ls += g_logp * (diff * diff * inv_var - 1.f);        // d/d log_std_j
            } else {
                const float dv = head - ret;
                d_out = A.vf_coef * 2.f * dv * inv_b;                  // d/d V
                st_vl += dv * dv;
                st_pl += -fminf(pl1, pl2);
                st_cf += (fabsf(ratio - 1.f) > A.clip_range) ? 1.f : 0.f;
                st_kl += (ratio - 1.f) - log_ratio;
            }
            g_head += d_out;
        }
        if (lane < 3 * PG_E) dOut[(col0 + e_of) * 4 + j_of] = d_out;
        __syncwarp();

        // ---- backward through the heads: dZ2 (overwrites H2 in place; acc still holds h2) -------
        {
            const float aw00 = W.headw[lane], aw01 = W.headw[lane + 32];
            const float aw10 = W.headw[64 + lane], aw11 = W.headw[96 + lane];
            const float cw0 = W.headw[128 + lane], cw1 = W.headw[160 + lane];
            float dz[4][PG_E];
#pragma unroll
            for (int e = 0; e < PG_E; ++e) {
                const float4 d = *reinterpret_cast<const float4*>(dOut + (col0 + e) * 4);
                const float h0 = acc[0][e], h1 = acc[1][e], h2 = acc[2][e], h3 = acc[3][e];
                gWh[0] = fmaf(d.x, h0, gWh[0]); gWh[1] = fmaf(d.x, h1, gWh[1]);
                gWh[2] = fmaf(d.y, h0, gWh[2]); gWh[3] = fmaf(d.y, h1, gWh[3]);
                gWh[4] = fmaf(d.z, h2, gWh[4]); gWh[5] = fmaf(d.z, h3, gWh[5]);
                dz[0][e] = (d.x * aw00 + d.y * aw10) * (1.f - h0 * h0);
                dz[1][e] = (d.x * aw01 + d.y * aw11) * (1.f - h1 * h1);
                dz[2][e] = d.z * cw0 * (1.f - h2 * h2);
                dz[3][e] = d.z * cw1 * (1.f - h3 * h3);
                gb2[0] += dz[0][e]; gb2[1] += dz[1][e]; gb2[2] += dz[2][e]; gb2[3] += dz[3][e];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float* dst = H2 + row_of(j) * PG_SP + col0;
                reinterpret_cast<float4*>(dst)[0] = make_float4(dz[j][0], dz[j][1], dz[j][2], dz[j][3]);
                reinterpret_cast<float4*>(dst)[1] = make_float4(dz[j][4], dz[j][5], dz[j][6], dz[j][7]);
            }
        }
        __syncwarp();

        // ---- backward data: dH1 = dZ2 * W2, dZ1 = dH1 * (1 - h1^2) ---------------------------------
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < PG_E; ++e) acc[j][e] = 0.f;
#pragma unroll 4
        for (int u = 0; u < HID; ++u) {
            float4 w = w2b[u * 32 + lane];
            const float* dp = H2 + u * PG_SP + col0;
            const float* dv = H2 + (64 + u) * PG_SP + col0;
            float4 p0 = reinterpret_cast<const float4*>(dp)[0], p1 = reinterpret_cast<const float4*>(dp)[1];
            float4 v0 = reinterpret_cast<const float4*>(dv)[0], v1 = reinterpret_cast<const float4*>(dv)[1];
            float a[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            float b[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int e = 0; e < PG_E; ++e) {
                acc[0][e] = fmaf(w.x, a[e], acc[0][e]);
                acc[1][e] = fmaf(w.y, a[e], acc[1][e]);
                acc[2][e] = fmaf(w.z, b[e], acc[2][e]);
                acc[3][e] = fmaf(w.w, b[e], acc[3][e]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* hsrc = H1 + row_of(j) * PG_SP + col0;
            float4 h0 = reinterpret_cast<const float4*>(hsrc)[0], h1 = reinterpret_cast<const float4*>(hsrc)[1];
            float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int e = 0; e < PG_E; ++e) {
                acc[j][e] *= (1.f - h[e] * h[e]);
                gb1[j] += acc[j][e];
            }
            float* dst = Z1 + row_of(j) * PG_SP + col0;
            reinterpret_cast<float4*>(dst)[0] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            reinterpret_cast<float4*>(dst)[1] = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
        }
        __syncthreads();

        // ---- weight gradients over the tile's 64 samples (CTA-level register tiles) ------------------
#pragma unroll
        for (int tower = 0; tower < 2; ++tower) {
            const float* dzb = H2 + tower * 64 * PG_SP;
            const float* hb = H1 + tower * 64 * PG_SP;
#pragma unroll 2
            for (int s4 = 0; s4 < PG_S / 4; ++s4) {
                float4 dz[4], h[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    dz[i] = *reinterpret_cast<const float4*>(dzb + (tu + 16 * i) * PG_SP + 4 * s4);
                    h[i] = *reinterpret_cast<const float4*>(hb + (tk + 16 * i) * PG_SP + 4 * s4);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float g = gW2[tower][i][j];
                        g = fmaf(dz[i].x, h[j].x, g);
                        g = fmaf(dz[i].y, h[j].y, g);
                        g = fmaf(dz[i].z, h[j].z, g);
                        g = fmaf(dz[i].w, h[j].w, g);
                        gW2[tower][i][j] = g;
                    }
            }
        }
        {
            const float* dzr = Z1 + (w1_tower * 64 + w1_unit) * PG_SP;
            const float* xb = X + (w1_half * KH) * PG_SP;
#pragma unroll 2
            for (int s4 = 0; s4 < PG_S / 4; ++s4) {
                const float4 dz = *reinterpret_cast<const float4*>(dzr + 4 * s4);
#pragma unroll
                for (int k = 0; k < KH; ++k) {
                    const float4 x = *reinterpret_cast<const float4*>(xb + k * PG_SP + 4 * s4);
                    gW1[k] = fmaf(dz.x, x.x, gW1[k]);
                    gW1[k] = fmaf(dz.y, x.y, gW1[k]);
                    gW1[k] = fmaf(dz.z, x.z, gW1[k]);
                    gW1[k] = fmaf(dz.w, x.w, gW1[k]);
                }
            }
        }
        __syncthreads();
    }

    // ---- write this CTA's partial gradient ------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            out[L.pw2 + (tu + 16 * i) * HID + tk + 16 * j] = gW2[0][i][j];
            out[L.vw2 + (tu + 16 * i) * HID + tk + 16 * j] = gW2[1][i][j];
        }
#pragma unroll
    for (int k = 0; k < KH; ++k) {
        int kk = w1_half * KH + k;
        if (kk < O) out[(w1_tower ? L.vw1 : L.pw1) + w1_unit * O + kk] = gW1[k];
    }
    // lane-owned sums: 14 floats per lane per warp -> cross-warp reduction in shared memory
    {
        float* r = red + (warp * 32 + lane) * 16;
        r[0] = gb1[0]; r[1] = gb1[1]; r[2] = gb1[2]; r[3] = gb1[3];
        r[4] = gb2[0]; r[5] = gb2[1]; r[6] = gb2[2]; r[7] = gb2[3];
        r[8] = gWh[0]; r[9] = gWh[1]; r[10] = gWh[2]; r[11] = gWh[3]; r[12] = gWh[4]; r[13] = gWh[5];
    }
    // per-sample-lane sums (lane = 3 e + j): reduce over e with shuffles (fixed order)
    auto sum_over_e = [&](float v) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < PG_E; ++e) s += __shfl_sync(0xffffffffu, v, (3 * e + (lane % 3)) & 31);
        return s;
    };
    float s_head = sum_over_e(g_head), s_ls = sum_over_e(g_ls);
    float s_pl = sum_over_e(st_pl), s_vl = sum_over_e(st_vl), s_cf = sum_over_e(st_cf), s_kl = sum_over_e(st_kl);
    __shared__ float red2[PG_WARPS][12];
    if (lane < 3) {
        red2[warp][lane] = s_head;          // d ab0, d ab1, d cb
        red2[warp][3 + lane] = s_ls;        // d log_std0, d log_std1, (unused)
    }
    if (lane == 2) { red2[warp][6] = s_pl; red2[warp][7] = s_vl; red2[warp][8] = s_cf; red2[warp][9] = s_kl; }
    __syncthreads();
    for (int it = tid; it < 32 * 14; it += PG_THREADS) {
        int ln = it / 14, q = it - ln * 14;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PG_WARPS; ++w) s += red[(w * 32 + ln) * 16 + q];
        int dst;
        if (q < 4) dst = ((q & 2) ? L.vb1 : L.pb1) + ln + ((q & 1) ? 32 : 0);
        else if (q < 8) dst = (((q - 4) & 2) ? L.vb2 : L.pb2) + ln + (((q - 4) & 1) ? 32 : 0);
        else if (q < 12) dst = L.aw + ((q - 8) >> 1) * HID + ln + (((q - 8) & 1) ? 32 : 0);
        else dst = L.cw + ln + ((q - 12) ? 32 : 0);
        out[dst] = s;
    }
    if (tid < 10) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PG_WARPS; ++w) s += red2[w][tid];
        const int sb = stat_base(O);
        if (tid < 2) out[L.ab + tid] = s;
        else if (tid == 2) out[L.cb] = s;
        else if (tid < 5) out[L.logstd + tid - 3] = s;   // entropy term added in the reduce kernel
        else if (tid == 5) { /* unused */ }
        else out[sb + tid - 6] = s;                      // policy_loss, value_loss, clip_frac, approx_kl sums
    }
}

template <int O_PAD>
__global__ void __launch_bounds__(PG_THREADS, 1) ppo_grad_kernel(GradArgs A, int O) {
    extern __shared__ __align__(16) float smem[];
    GradSmem S = grad_stage<O_PAD>(smem, A.params, O, true);
    __syncthreads();
    grad_minibatch<O_PAD>(A, O, S, A.partials + (size_t)blockIdx.x * grad_stride(O));
}

// Tensor-core version (ppo_tc.cuh): even CTAs own the policy tower, odd CTAs the value tower.
template <int KP>
__global__ void __launch_bounds__(tc::THREADS, 1) ppo_grad_tc_kernel(GradArgs A, int O) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    MR_TR(0);
    tc::Ctx C = tc::make_ctx(smem_raw, blockIdx.x & 1);
    tc::setup(C);
    tc::stage<KP>(C, A.params, O);
    const tc::Sched S{A.mb_size, A.mb_size, 1, (int)(blockIdx.x >> 1), (int)(gridDim.x >> 1)};
    tc::Pipe<KP> Q;
    tc::pipe_start<KP>(Q, A, S, O, threadIdx.x & 127, threadIdx.x >> 7, C.tower == 0);
    __syncthreads();
    const tc::MbConst MK = tc::mb_const(A.mb_stats, 0, A.normalize_adv);
    tc::minibatch<KP>(C, A, S, MK, 0, Q, O, A.partials + (size_t)blockIdx.x * grad_stride(O));
    tc::teardown(C);
}

// grad[p] = sum over CTAs; finishes the scalar terms.  One thread per parameter; the n_parts
// loads are independent, so they are issued 8 deep into 4 interleaved accumulators that are
// combined in a fixed order (deterministic for a given n_parts).
__global__ void __launch_bounds__(128)
ppo_reduce_kernel(const float* __restrict__ partials, int n_parts, int O, float ent_coef,
                  const double* __restrict__ mb_stats, float* __restrict__ grad, float rank_share, int towers) {
    const int stride = grad_stride(O);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= stride) return;
    const float* src = partials + p;
    if (towers) {  // tensor-core kernel: CTA b holds tower (b & 1) only
        src += (size_t)tc::param_tower(p, make_layout(O)) * stride;
        n_parts >>= 1;
        // consecutive parts of one tower are 2 rows apart
        const int st2 = 2 * stride;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int b = 0;
        for (; b + 8 <= n_parts; b += 8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = __ldg(src + (size_t)(b + q) * st2);
            a0 += v[0]; a1 += v[1]; a2 += v[2]; a3 += v[3];
            a0 += v[4]; a1 += v[5]; a2 += v[6]; a3 += v[7];
        }
        for (; b < n_parts; ++b) a0 += __ldg(src + (size_t)b * st2);
        float s = (a0 + a1) + (a2 + a3);
        const ParamLayout L = make_layout(O);
        const int sb = stat_base(O);
        if (p >= L.logstd && p < L.logstd + ACT) s -= ent_coef * rank_share;
        if (p >= sb && p < sb + 4) s *= (float)(1.0 / mb_stats[2]);
        if (p >= sb + 4 || (p >= L.total && p < sb)) s = 0.f;
        grad[p] = s;
        return;
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int b = 0;
    for (; b + 8 <= n_parts; b += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __ldg(src + (size_t)(b + q) * stride);
        a0 += v[0]; a1 += v[1]; a2 += v[2]; a3 += v[3];
        a0 += v[4]; a1 += v[5]; a2 += v[6]; a3 += v[7];
    }
    for (; b < n_parts; ++b) a0 += __ldg(src + (size_t)b * stride);
    float s = (a0 + a1) + (a2 + a3);
    const ParamLayout L = make_layout(O);
    const int sb = stat_base(O);
    if (p >= L.logstd && p < L.logstd + ACT) s -= ent_coef * rank_share;  // d(-ent_coef * mean entropy)/d log_std
    if (p >= sb && p < sb + 4) s *= (float)(1.0 / mb_stats[2]);
    if (p >= sb + 4 || (p >= L.total && p < sb)) s = 0.f;
    grad[p] = s;
}

struct AdamArgs {
    float* params;
    float* exp_avg;
    float* exp_avg_sq;
    const float* grad;
    int64_t* step;      // [0] Adam step count (state["step"]); [1] launch-internal ticket (starts at 0)
    float lr, beta1, beta2, eps, max_grad_norm;
    float* info;        // [8]: total_norm, clip_coef, step, 0, then the gradient's stats tail
    int n_params;       //      (policy_loss, value_loss, clip_fraction, approx_kl)
};

constexpr int ADAM_THREADS = 256;

// clip_grad_norm_ + torch.optim.Adam (single-tensor path, torch 2.0.1 arithmetic order).
// The vector is only 42-48 KB, so the kernel is latency-bound: one parameter per thread,
// every CTA recomputes the global norm from L2 in the same fixed order (bit-identical across
// CTAs, no inter-CTA barrier), and the last CTA to finish bumps the step counter.
__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(AdamArgs A) {
    __shared__ double s_part[ADAM_THREADS / 32];
    __shared__ double s_bc[2];
    const int tid = threadIdx.x;
    const int64_t step = A.step[0] + 1;
    if (tid == 0) {
        s_bc[0] = 1.0 - pow((double)A.beta1, (double)step);
        s_bc[1] = sqrt(1.0 - pow((double)A.beta2, (double)step));
    }
    // global grad norm: thread t sums elements t, t + 256, ... (independent loads)
    double sq = 0.0;
    {
        constexpr int U = 8;
        int p = tid;
        for (; p + (U - 1) * ADAM_THREADS < A.n_params; p += U * ADAM_THREADS) {
            float v[U];
#pragma unroll
            for (int q = 0; q < U; ++q) v[q] = __ldg(A.grad + p + q * ADAM_THREADS);
#pragma unroll
            for (int q = 0; q < U; ++q) sq += (double)v[q] * (double)v[q];
        }
        for (; p < A.n_params; p += ADAM_THREADS) {
            float v = __ldg(A.grad + p);
            sq += (double)v * (double)v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((tid & 31) == 0) s_part[tid >> 5] = sq;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < ADAM_THREADS / 32; ++w) tot += s_part[w];
    const float total_norm = (float)sqrt(tot);
    const float coef = fminf(A.max_grad_norm / (total_norm + 1e-6f), 1.0f);
    const float neg_step_size = (float)(-(double)A.lr / s_bc[0]);
    const float bc2_sqrt = (float)s_bc[1];
    const float omb1 = 1.f - A.beta1, omb2 = 1.f - A.beta2;
    const int p = blockIdx.x * ADAM_THREADS + tid;
    if (p < A.n_params) {
        const float g = __fmul_rn(A.grad[p], coef);
        const float m = __fadd_rn(__fmul_rn(A.exp_avg[p], A.beta1), __fmul_rn(g, omb1));
        const float v = __fadd_rn(__fmul_rn(A.exp_avg_sq[p], A.beta2), __fmul_rn(__fmul_rn(g, g), omb2));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), A.eps);
        A.params[p] = __fadd_rn(A.params[p], __fdiv_rn(__fmul_rn(neg_step_size, m), denom));
        A.exp_avg[p] = m;
        A.exp_avg_sq[p] = v;
    }
    if (blockIdx.x == 0 && tid == 0 && A.info) {
        A.info[0] = total_norm; A.info[1] = coef; A.info[2] = (float)step; A.info[3] = 0.f;
        const int sb = (A.n_params + 3) & ~3;
        for (int q = 0; q < 4; ++q) A.info[4 + q] = A.grad[sb + q];
    }
    // last CTA out increments the step (every CTA has read it by then) and re-arms the ticket
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        unsigned long long* ticket = reinterpret_cast<unsigned long long*>(A.step + 1);
        if (atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1) {
            A.step[0] = step;
            *ticket = 0ull;
        }
    }
}

// Per-minibatch advantage sums of one epoch's permutation: stats[mb] = (sum, sum of squares,
// count) in float64 (shifted by the first element to keep the variance well conditioned is not
// needed at float64).  One CTA per minibatch.
__global__ void __launch_bounds__(1024)
adv_stats_kernel(const float* __restrict__ adv, const int64_t* __restrict__ perm,
                 int64_t n_samples, int64_t batch, int64_t N, int64_t T,
                 double* __restrict__ stats) {
    const int64_t mb = blockIdx.x;
    const int64_t s0 = mb * batch, s1 = min(n_samples, s0 + batch);
    double s = 0.0, q = 0.0;
    constexpr int U = 4;   // independent gathers in flight per thread
    for (int64_t i0 = s0 + threadIdx.x; i0 < s1; i0 += (int64_t)U * blockDim.x) {
        float a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + (int64_t)u * blockDim.x;
            a[u] = 0.f;
            if (i < s1) {
                const int64_t id = perm[i];
                const int64_t n = id / T, t = id - n * T;
                a[u] = __ldg(adv + t * N + n);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s += (double)a[u];
            q += (double)a[u] * (double)a[u];
        }
    }
    __shared__ double sh[2][32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0, tq = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ts += sh[0][w]; tq += sh[1][w]; }
        stats[3 * mb] = ts;
        stats[3 * mb + 1] = tq;
        stats[3 * mb + 2] = (double)(s1 - s0);
    }
}

// env-major sample ids (RolloutBuffer.swap_and_flatten order, n * T + t) -> rows of the time-major
// buffers (t * N + n), once per epoch, so that the gather in the tensor-core kernels needs no division
__global__ void __launch_bounds__(256) perm_to_rows_kernel(const int64_t* __restrict__ perm, int64_t n_samples,
                                                           int64_t N, int64_t T, int32_t* __restrict__ rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) return;
    const unsigned id = (unsigned)perm[i];
    const unsigned n = id / (unsigned)T, t = id - n * (unsigned)T;
    rows[i] = (int32_t)((int64_t)t * N + n);
}

// Device-side index stream for RolloutBuffer.get when the permutation need not come from the host:
// out[i] = P(i), P a keyed bijection of [0, n) -- four rounds of (odd multiply, xor-shift, add key)
// on the enclosing power-of-two domain, cycle-walked back into range.  One thread per index, no
// sort (torch.randperm costs 0.17 ms per epoch at n = 1.2e6; this is a 10 MB store).
struct PermKey {
    uint32_t mul[4], add[4];
    int bits;
};
__device__ __forceinline__ uint32_t perm_mix(uint32_t x, const PermKey& K) {
    const uint32_t mask = K.bits >= 32 ? 0xFFFFFFFFu : ((1u << K.bits) - 1u);
    const int sh = (K.bits + 1) >> 1;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        x = (x * K.mul[r] + K.add[r]) & mask;
        x ^= x >> sh;
    }
    return x;
}
__global__ void __launch_bounds__(256) device_perm_kernel(int64_t* __restrict__ out, int64_t n, PermKey K) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = (uint32_t)i;
    do { x = perm_mix(x, K); } while ((int64_t)x >= n);
    out[i] = (int64_t)x;
}

// ---------------------------------------------------------------------------------------------
// Persistent epoch kernel: every minibatch of one epoch in ONE cooperative launch (one CTA per SM).
// Per minibatch: [stage parameters] -> forward/backward (grad_minibatch) -> grid barrier ->
// each CTA reduces its slice of the gradient over all CTAs' partials [-> one-shot all-reduce over
// NVLink peer memory: push the slice into every peer's inbox as tagged 8-byte packets, spin on the
// tags, sum in rank order] -> slice
// sum of squares -> grid barrier -> global-norm clip + Adam on the slice -> grid barrier.
// Replaces 3 launches (+ an NCCL call) per minibatch; everything is summed in a fixed order, so
// parameters stay bit-identical across ranks.
struct PeerXchg {
    int world, rank;
    int* err;                           // set to 1 when a peer's packets did not arrive in time
    unsigned long long* inbox;          // local  [2][world][stride] packets {seq << 32 | float bits}
    unsigned long long* peer_inbox[8];  // peers' inboxes (NVLink-mapped)
};

struct EpochArgs {
    GradArgs G;
    int64_t n_samples, batch;
    const double* stats;       // [n_mb][3] (global minibatch sums)
    const float* rank_share;   // [n_mb] local count / global count, NULL -> 1
    float *params, *exp_avg, *exp_avg_sq;
    int64_t* step;
    float lr, beta1, beta2, eps, max_grad_norm;
    float* grad;               // [stride] last reduced gradient
    float* info;               // [n_mb][8] or NULL
    double* sq;                // [n_cta] scratch (v1 kernel)
    float* acc;                // [3][2 * TL.size] rotating gradient accumulators (zero at launch)
    float* gsum;               // [2 * TL.size] gradient summed over ranks (world > 1)
    unsigned* barrier;         // [1], zero at launch
    unsigned seq0;             // exchange sequence number before this launch
    PeerXchg X;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, unsigned n_cta) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += n_cta;
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_gpu(counter) < target) {}
    }
    __syncthreads();
}

template <int O_PAD>
__global__ void __launch_bounds__(PG_THREADS, 1) ppo_epoch_kernel(EpochArgs E, int O) {
    constexpr bool TC = false;   // fp32 CUDA-core cross-check path; the product path is ppo_epoch_tc_kernel
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_grp[PG_THREADS];
    __shared__ double s_dred[PG_WARPS];
    __shared__ double s_scal[4];
    const ParamLayout L = make_layout(O);
    const int tid = threadIdx.x;
    const int G = gridDim.x, c = blockIdx.x;
    const int stride = grad_stride(O), sb = stat_base(O), n_params = L.total;
    const int slice = (((stride + G - 1) / G) + 3) & ~3;
    const int groups = PG_THREADS / slice;            // part-groups summing in parallel
    const int p0 = c * slice;
    const int64_t n_mb = (E.n_samples + E.batch - 1) / E.batch;
    unsigned target = 0;
    int64_t step = E.step[0];
    const float omb1 = 1.f - E.beta1, omb2 = 1.f - E.beta2;
    // partial gradients of parameter p: every CTA (SIMT) or the CTAs of p's tower (tensor-core path)
    const int n_src = TC ? G >> 1 : G;
    const size_t src_step = TC ? 2 * (size_t)stride : (size_t)stride;

    for (int64_t m = 0; m < n_mb; ++m) {
        GradArgs A = E.G;
        A.params = E.params;
        A.perm = E.G.perm + m * E.batch;
        A.mb_size = min(E.batch, E.n_samples - m * E.batch);
        A.mb_stats = E.stats + 3 * m;
        {
            GradSmem S = grad_stage<O_PAD>(smem, E.params, O, m == 0);
            __syncthreads();
            grad_minibatch<O_PAD>(A, O, S, A.partials + (size_t)c * stride);
        }
        grid_barrier(E.barrier, target, G);
        MR_TR(4);

        // ---- phase B: this CTA's slice of the gradient --------------------------------------------
        ++step;
        if (tid == PG_THREADS - 1) {  // bias corrections, off the critical path
            s_scal[0] = 1.0 - pow((double)E.beta1, (double)step);
            s_scal[1] = sqrt(1.0 - pow((double)E.beta2, (double)step));
        }
        const int j = tid % slice, g = tid / slice;
        const int p = p0 + j;
        float part = 0.f;
        if (g < groups && p < stride) {
            const float* src = A.partials + p + (TC ? (size_t)tc::param_tower(p, L) * stride : 0);
            float a0 = 0.f, a1 = 0.f;
            int b = g;
            for (; b + 7 * groups < n_src; b += 8 * groups) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = __ldcg(src + (size_t)(b + q * groups) * src_step);
                a0 += v[0]; a1 += v[1]; a0 += v[2]; a1 += v[3];
                a0 += v[4]; a1 += v[5]; a0 += v[6]; a1 += v[7];
            }
            for (; b < n_src; b += groups) a0 += __ldcg(src + (size_t)b * src_step);
            part = a0 + a1;
        }
        s_grp[tid] = part;
        __syncthreads();
        float val = 0.f;
        const bool own = tid < slice && p < stride;
        if (own) {
            for (int q = 0; q < groups; ++q) val += s_grp[q * slice + j];
            const float share = E.rank_share ? E.rank_share[m] : 1.f;
            if (p >= L.logstd && p < L.logstd + ACT) val -= E.G.ent_coef * share;
            if (p >= sb && p < sb + 4) val *= (float)(1.0 / A.mb_stats[2]);
            if (p >= sb + 4 || (p >= n_params && p < sb)) val = 0.f;
        }
        if (E.X.world > 1) {
            // One-shot all-reduce over NVLink peer memory, low-latency protocol: every element
            // travels as one 8-byte store {value, sequence tag}; the receiver spins on the tag of
            // the element itself, so there is no fence, no flag and no second round trip.
            // Two parity slots: slot (seq & 1) is rewritten at seq + 2, which a rank can only
            // reach after every peer has pushed seq + 1, i.e. after it consumed seq.
            const unsigned seq = E.seq0 + (unsigned)m + 1u;
            const int slot = seq & 1u;
            if (own) {
                const unsigned long long pkt =
                    ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(val);
                for (int r = 0; r < E.X.world; ++r)
                    if (r != E.X.rank)
                        __stcg(E.X.peer_inbox[r] + ((size_t)slot * E.X.world + E.X.rank) * stride + p, pkt);
                float tot = 0.f;
                for (int r = 0; r < E.X.world; ++r) {  // rank order: bit-identical on every rank
                    if (r == E.X.rank) { tot += val; continue; }
                    const unsigned long long* src = E.X.inbox + ((size_t)slot * E.X.world + r) * stride + p;
                    unsigned long long v;
                    do {
                        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
                    } while ((unsigned)(v >> 32) != seq);
                    tot += __uint_as_float((unsigned)v);
                }
                val = tot;
            }
        }
        double sq = 0.0;
        if (own) {
            E.grad[p] = val;
            if (p < n_params) sq = (double)val * (double)val;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if ((tid & 31) == 0) s_dred[tid >> 5] = sq;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < PG_WARPS; ++w) t += s_dred[w];
            E.sq[c] = t;
        }
        MR_TR(5);
        grid_barrier(E.barrier, target, G);
        MR_TR(6);

        // ---- phase C: global-norm clip + Adam on the slice ----------------------------------------------
        if (tid < 32) {
            double t = 0.0;
            for (int b = tid; b < G; b += 32) t += __ldcg(E.sq + b);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (tid == 0) s_scal[2] = t;
        }
        __syncthreads();
        const float total_norm = (float)sqrt(s_scal[2]);
        const float coef = fminf(E.max_grad_norm / (total_norm + 1e-6f), 1.0f);
        const float neg_step_size = (float)(-(double)E.lr / s_scal[0]);
        const float bc2_sqrt = (float)s_scal[1];
        if (own && p < n_params) {
            const float gq = __fmul_rn(val, coef);
            const float mq = __fadd_rn(__fmul_rn(E.exp_avg[p], E.beta1), __fmul_rn(gq, omb1));
            const float vq = __fadd_rn(__fmul_rn(E.exp_avg_sq[p], E.beta2), __fmul_rn(__fmul_rn(gq, gq), omb2));
            const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vq), bc2_sqrt), E.eps);
            E.params[p] = __fadd_rn(E.params[p], __fdiv_rn(__fmul_rn(neg_step_size, mq), denom));
            E.exp_avg[p] = mq;
            E.exp_avg_sq[p] = vq;
        }
        if (E.info) {
            float* row = E.info + 8 * m;
            if (c == 0 && tid == 0) { row[0] = total_norm; row[1] = coef; row[2] = (float)step; row[3] = 0.f; }
            if (own && p >= sb && p < sb + 4) row[4 + p - sb] = val;
        }
        MR_TR(7);
        grid_barrier(E.barrier, target, G);
        MR_TR(8);
    }
    if (c == 0 && tid == 0) E.step[0] = step;
}

// Tensor-core epoch kernel.  Per minibatch: tiles (tc::minibatch, rows prefetched one tile ahead)
// -> grid barrier -> 16-byte slice reduction of the partial gradients [-> NVLink all-reduce] ->
// grid barrier -> global-norm clip + Adam on the slice -> grid barrier -> tc::restage.
// Every hand-off through L2 is one wave of independent 8/16-byte loads: with all 148 SMs pulling
// at once a round trip costs 1400-2900 cycles (tools/ubench/l2_handoff.cu), so the phases are
// shaped to pay it once each.  (Tried and measured slower: every CTA applying Adam to its whole
// tower to save the third barrier -- 4 x the L2 traffic, 15 us instead of 10 us per minibatch.)
template <int KP>
__global__ void __launch_bounds__(tc::THREADS, 1) ppo_epoch_tc_v1_kernel(EpochArgs E, int O) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ float4 s_grp[tc::THREADS];
    __shared__ float4 s_x[8 * 32];   // peers' packets of the own threads (tid < q4 <= 32)
    const ParamLayout L = make_layout(O);
    const int tid = threadIdx.x;
    const int G = gridDim.x, c = blockIdx.x;
    const int stride = grad_stride(O), sb = stat_base(O), n_params = L.total;
    const int slice = (((stride + G - 1) / G) + 3) & ~3;
    const int q4 = slice >> 2;                          // 16-byte lanes per slice (<= 32, checked on the host)
    const int groups = tc::THREADS / q4;                // part-groups summing in parallel
    const int p0 = c * slice;
    const int n_mb = (int)((E.n_samples + E.batch - 1) / E.batch);
    const int n_src = G >> 1;
    const int rot = c % n_src;
    unsigned target = 0;
    int64_t step = E.step[0];
    double b1pow = pow((double)E.beta1, (double)step), b2pow = pow((double)E.beta2, (double)step);

    MR_TR(0);
    tc::Ctx C = tc::make_ctx(smem_raw, c & 1);
    tc::setup(C);
    GradArgs A = E.G;
    const tc::Sched S{E.n_samples, E.batch, n_mb, c >> 1, G >> 1};
    tc::Pipe<KP> Q;
    tc::pipe_start<KP>(Q, A, S, O, tid & 127, tid >> 7, C.tower == 0);

    tc::stage<KP>(C, E.params, O);
    __syncthreads();

    // this CTA's slice of the fp32 master copy and the Adam moments lives in registers of the
    // `own` threads for the whole epoch (nobody else writes it): no load on the critical path
    const int own_p = p0 + 4 * (tid % q4);
    const bool own_t = tid < q4 && own_p < stride && own_p < n_params;
    // (the arrays hold n_params floats, which is not a multiple of 4: the last quad is partial)
    auto load4 = [&](const float* a) {
        if (own_p + 3 < n_params) return *reinterpret_cast<const float4*>(a + own_p);
        float4 r = make_float4(a[own_p], 0.f, 0.f, 0.f);
        if (own_p + 1 < n_params) r.y = a[own_p + 1];
        if (own_p + 2 < n_params) r.z = a[own_p + 2];
        return r;
    };
    auto store4 = [&](float* a, const float4& r) {
        if (own_p + 3 < n_params) { *reinterpret_cast<float4*>(a + own_p) = r; return; }
        a[own_p] = r.x;
        if (own_p + 1 < n_params) a[own_p + 1] = r.y;
        if (own_p + 2 < n_params) a[own_p + 2] = r.z;
    };
    float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), v4 = m4, w4 = m4;
    if (own_t) {
        m4 = load4(E.exp_avg);
        v4 = load4(E.exp_avg_sq);
        w4 = load4(E.params);
    }

    tc::MbConst MK = tc::mb_const(E.stats, 0, A.normalize_adv);
    for (int m = 0; m < n_mb; ++m) {
        // inputs of the reduction that do not depend on other CTAs: fetched ahead of the barrier
        const float share = E.rank_share ? __ldg(E.rank_share + m) : 1.f;
        const float inv_cnt = MK.inv_b;
        MR_TR(2);
        tc::minibatch<KP>(C, A, S, MK, m, Q, O, A.partials + (size_t)c * stride);
        MR_TR(3);
        grid_barrier(E.barrier, target, G);
        MR_TR(4);
        if (m + 1 < n_mb) MK = tc::mb_const(E.stats, m + 1, A.normalize_adv);   // lands under the exchange below

        // ---- this CTA's slice of the gradient: sum over the CTAs of each parameter's tower -----------
        ++step;
        b1pow *= (double)E.beta1;   // beta^step, carried from one pow() per launch
        b2pow *= (double)E.beta2;
        const int j = tid % q4, g = tid / q4;
        const int p = p0 + 4 * j;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < groups && p < stride) {
            const int t0 = tc::param_tower(p, L), t3 = tc::param_tower(p + 3, L);
            const bool s1 = tc::param_tower(p + 1, L) != 0, s2 = tc::param_tower(p + 2, L) != 0;
            const size_t step4 = 2 * (size_t)stride;
            // All of a thread's loads are issued before the first add (RW at a time; the sums keep their
            // order).  Written as a plain load-add loop the compiler kept ONE load in flight per thread, and
            // the scoreboard stall at each add made the 5-6 loads of a thread 5-6 dependent L2 round trips:
            // 3.3 us per minibatch for this phase.
            constexpr int RW = 6;
            if (t0 == t3 && s1 == (t0 != 0) && s2 == (t0 != 0)) {
                const float* src = A.partials + (size_t)t0 * stride + p;
                for (int base = g; base < n_src; base += RW * groups) {
                    float4 v[RW];
#pragma unroll
                    for (int it = 0; it < RW; ++it) {
                        const int b0 = base + it * groups;
                        if (b0 < n_src) {
                            int bb = b0 + rot;   // CTAs start on different rows: no L2 hot spot
                            bb = bb >= n_src ? bb - n_src : bb;
                            v[it] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)bb * step4));
                        }
                    }
#pragma unroll
                    for (int it = 0; it < RW; ++it) {
                        if (base + it * groups < n_src) { acc.x += v[it].x; acc.y += v[it].y; acc.z += v[it].z; acc.w += v[it].w; }
                    }
                }
            } else {  // the quad straddles a tower boundary: pick per element
                const float* src = A.partials + p;
                for (int base = g; base < n_src; base += RW * groups) {
                    float4 v0[RW], v1[RW];
#pragma unroll
                    for (int it = 0; it < RW; ++it) {
                        const int b0 = base + it * groups;
                        if (b0 < n_src) {
                            int bb = b0 + rot;
                            bb = bb >= n_src ? bb - n_src : bb;
                            v0[it] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)bb * step4));
                            v1[it] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)bb * step4 + stride));
                        }
                    }
#pragma unroll
                    for (int it = 0; it < RW; ++it) {
                        if (base + it * groups < n_src) {
                            acc.x += t0 ? v1[it].x : v0[it].x;
                            acc.y += s1 ? v1[it].y : v0[it].y;
                            acc.z += s2 ? v1[it].z : v0[it].z;
                            acc.w += t3 ? v1[it].w : v0[it].w;
                        }
                    }
                }
            }
        }
        MR_TR(30);
        s_grp[tid] = acc;
        __syncthreads();
        MR_TR(31);
        const bool own = tid < q4 && p < stride;
        float val[4] = {0.f, 0.f, 0.f, 0.f};
        if (own) {
            for (int q = 0; q < groups; ++q) {
                const float4 v = s_grp[q * q4 + j];
                val[0] += v.x; val[1] += v.y; val[2] += v.z; val[3] += v.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int pe = p + e;
                if (pe >= L.logstd && pe < L.logstd + ACT) val[e] -= E.G.ent_coef * share;
                if (pe >= sb && pe < sb + 4) val[e] *= inv_cnt;
                if (pe >= sb + 4 || (pe >= n_params && pe < sb)) val[e] = 0.f;
            }
        }
        if (E.X.world > 1 && own) {
            // one-shot all-reduce over NVLink peer memory, tagged 8-byte packets (see ppo_epoch_kernel)
            const unsigned seq = E.seq0 + (unsigned)m + 1u;
            const int slot = seq & 1u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const unsigned long long pkt =
                    ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(val[e]);
                for (int r = 0; r < E.X.world; ++r)
                    if (r != E.X.rank)
                        __stcg(E.X.peer_inbox[r] + ((size_t)slot * E.X.world + E.X.rank) * stride + p + e, pkt);
            }
            // Pull: the four packets of every peer, four peers per wave of 16-byte loads (the packets land
            // in local memory; polling them one after the other cost world x 4 dependent L2 round trips:
            // 13.6 us per minibatch at 8 GPUs).  Payloads wait in shared memory (not in 32 registers) and
            // are summed in rank order: bit-identical on every rank.
            const unsigned long long* src0 = E.X.inbox + (size_t)slot * E.X.world * stride + p;
            unsigned pending = ((1u << E.X.world) - 1u) & ~(1u << E.X.rank);
            while (pending) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    ulonglong2 lo[4], hi[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int r = 4 * half + k;
                        if (pending >> r & 1u) {
                            const unsigned long long* q = src0 + (size_t)r * stride;
                            asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo[k].x), "=l"(lo[k].y) : "l"(q) : "memory");
                            asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(hi[k].x), "=l"(hi[k].y) : "l"(q + 2) : "memory");
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int r = 4 * half + k;
                        if ((pending >> r & 1u) && (unsigned)(lo[k].x >> 32) == seq && (unsigned)(lo[k].y >> 32) == seq &&
                            (unsigned)(hi[k].x >> 32) == seq && (unsigned)(hi[k].y >> 32) == seq) {
                            s_x[r * 32 + tid] = make_float4(__uint_as_float((unsigned)lo[k].x), __uint_as_float((unsigned)lo[k].y),
                                                            __uint_as_float((unsigned)hi[k].x), __uint_as_float((unsigned)hi[k].y));
                            pending &= ~(1u << r);
                        }
                    }
                }
            }
            float tot[4] = {0.f, 0.f, 0.f, 0.f};
            for (int r = 0; r < E.X.world; ++r) {
                const float4 v = r == E.X.rank ? make_float4(val[0], val[1], val[2], val[3]) : s_x[r * 32 + tid];
                tot[0] += v.x; tot[1] += v.y; tot[2] += v.z; tot[3] += v.w;
            }
            val[0] = tot[0]; val[1] = tot[1]; val[2] = tot[2]; val[3] = tot[3];
        }
        MR_TR(32);
        double sq = 0.0;
        if (own) {
            *reinterpret_cast<float4*>(E.grad + p) = make_float4(val[0], val[1], val[2], val[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (p + e < n_params) sq += (double)val[e] * (double)val[e];
        }
        if (tid < 32) {   // own threads all sit in warp 0 (q4 <= 32)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if (tid == 0) E.sq[c] = sq;
        }
        MR_TR(5);
        grid_barrier(E.barrier, target, G);
        MR_TR(6);

        // ---- global-norm clip + Adam on the slice (torch's single-tensor arithmetic, as adam_kernel) ---
        if (tid < 32) {
            double t = 0.0;
            for (int b = tid; b < G; b += 32) t += __ldcg(E.sq + b);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            t = __shfl_sync(0xffffffffu, t, 0);
            const float total_norm = (float)sqrt(t);
            const float coef = fminf(E.max_grad_norm / (total_norm + 1e-6f), 1.0f);
            if (own && p < n_params) {   // p is 4-aligned; n_params is not: the last quad is partial
                const float neg_step_size = (float)(-(double)E.lr / (1.0 - b1pow));
                const float bc2_sqrt = (float)sqrt(1.0 - b2pow);
                const float omb1 = 1.f - E.beta1, omb2 = 1.f - E.beta2;
                float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float gq = __fmul_rn(val[e], coef);
                    const float mq = __fadd_rn(__fmul_rn(mm[e], E.beta1), __fmul_rn(gq, omb1));
                    const float vq = __fadd_rn(__fmul_rn(vv[e], E.beta2), __fmul_rn(__fmul_rn(gq, gq), omb2));
                    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vq), bc2_sqrt), E.eps);
                    const float nw = __fadd_rn(ww[e], __fdiv_rn(__fmul_rn(neg_step_size, mq), denom));
                    const bool ok = p + e < n_params;
                    mm[e] = ok ? mq : mm[e]; vv[e] = ok ? vq : vv[e]; ww[e] = ok ? nw : ww[e];
                }
                m4 = make_float4(mm[0], mm[1], mm[2], mm[3]);
                v4 = make_float4(vv[0], vv[1], vv[2], vv[3]);
                w4 = make_float4(ww[0], ww[1], ww[2], ww[3]);
                store4(E.params, w4);   // read by every CTA of the tower (restage)
            }
            if (E.info) {
                float* row = E.info + 8 * m;
                if (c == 0 && tid == 0) { row[0] = total_norm; row[1] = coef; row[2] = (float)step; row[3] = 0.f; }
                if (own && p == sb) { row[4] = val[0]; row[5] = val[1]; row[6] = val[2]; row[7] = val[3]; }
            }
        }
        MR_TR(7);
        grid_barrier(E.barrier, target, G);
        MR_TR(8);
        if (m + 1 < n_mb) {
            tc::restage<KP>(C, E.params, O);
            __syncthreads();   // misc floats are read at the top of the next minibatch
        }
        MR_TR(37);
    }
    if (own_t) {
        store4(E.exp_avg, m4);
        store4(E.exp_avg_sq, v4);
    }
    if (c == 0 && tid == 0) E.step[0] = step;
    tc::teardown(C);
}

// ---------------------------------------------------------------------------------------------
// Tensor-core epoch kernel, second form (round 2).  What changed against ppo_epoch_tc_v1_kernel is
// everything BETWEEN the tiles of consecutive minibatches -- 47 % of the v1 kernel's time:
//   v1: 148 partial vectors -> L2 | barrier | every CTA pulls its 72-float slice from 74 partials |
//       barrier | clip + Adam on the slice | barrier | every CTA pulls its tower's new parameters.
//   now: every CTA adds its partial into ONE accumulator with a single bulk reduction
//       (cp.reduce.async.bulk .add.f32, shared -> L2, 22 KB) | ONE barrier | every CTA reads the
//       whole reduced gradient (one wave of 12 float4 loads per thread), derives the clip
//       coefficient itself and applies Adam to ITS tower, whose fp32 master copy and moments live in
//       its shared memory for the whole epoch (74-fold redundant arithmetic instead of two more
//       grid-wide hand-offs); the updated W2 goes straight from the Adam registers into the fp16
//       operand panel.
// Three accumulators rotate (minibatch m uses m % 3; after barrier j the CTAs clear their slices of
// buffer (j - 1) % 3, whose readers all passed barrier j - 1 ... j).  The order in which the L2
// adds the 74 contributions is not fixed, so a single-GPU run is reproducible to rounding only
// (1e-7 of the gradient); every CTA reads the same sums, so the towers' copies -- and with more
// than one rank, the ranks -- stay bit-identical.
// World > 1: after the barrier the owner of each accumulator slice pushes it to every peer (tagged
// 8-byte packets over NVLink, as in v1), sums the peers' slices in rank order into `gsum`, and a
// second barrier publishes the global sum.
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const float* ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
                 "r"(umma::smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.global;" ::: "memory");   // async-proxy writes before the generic release below
}

template <int KP>
__global__ void __launch_bounds__(tc::THREADS, 1) ppo_epoch_tc_kernel(EpochArgs E, int O) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ double s_sq[tc::WARPS];
    __shared__ float4 s_x[8 * 32];   // peers' packets of the slice threads (tid < q4 <= 32)
    const ParamLayout L = make_layout(O);
    const tc::TowerLayout TL = tc::make_tl(O);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, c = blockIdx.x;
    const int n_mb = (int)((E.n_samples + E.batch - 1) / E.batch);
    const int acc_floats = 2 * TL.size;
    const int nq = acc_floats >> 2;        // quads of the accumulator
    const int q4 = (nq + G - 1) / G;       // quads per CTA slice (<= 32, checked on the host)
    const int my_quad = c * q4 + tid;
    const bool slice_t = tid < q4 && my_quad < nq;
    unsigned target = 0;
    int64_t step = E.step[0];
    double b1pow = pow((double)E.beta1, (double)step), b2pow = pow((double)E.beta2, (double)step);

    MR_TR(0);
    tc::Ctx C = tc::make_ctx(smem_raw, c & 1);
    tc::setup(C);
    float* st_w = reinterpret_cast<float*>(C.base + tc::OFF_STATE);
    float* st_m = st_w + TL.size;
    float* st_v = st_m + TL.size;
    float* stage = reinterpret_cast<float*>(C.base + tc::OFF_H1);   // H1 + dZ panels are idle between minibatches
    C.p_b2 = st_w + TL.b2;
    C.p_hw = st_w + TL.hw;
    C.p_hs = st_w + TL.hs;
    GradArgs A = E.G;
    const tc::Sched S{E.n_samples, E.batch, n_mb, c >> 1, G >> 1};
    tc::Pipe<KP> Q;
    tc::pipe_start<KP>(Q, A, S, O, tid & 127, tid >> 7, C.tower == 0);

    // the tower's parameters and Adam moments: flat global arrays -> shared memory, TL order
    for (int i = tid; i < TL.size; i += tc::THREADS) {
        const int p = tc::tl_to_flat(i, C.tower, TL, L, O);
        st_w[i] = p >= 0 ? E.params[p] : 0.f;
        st_m[i] = p >= 0 ? E.exp_avg[p] : 0.f;
        st_v[i] = p >= 0 ? E.exp_avg_sq[p] : 0.f;
    }
    __syncthreads();
    tc::panel_w1_from_tl<KP>(C, st_w, TL, O);
    tc::panel_w2_from_tl(C, st_w, TL);
    __syncthreads();

    tc::MbConst MK = tc::mb_const(E.stats, 0, A.normalize_adv);
    for (int m = 0; m < n_mb; ++m) {
        const float inv_cnt = MK.inv_b;
        float* accb = E.acc + (size_t)(m % 3) * acc_floats;
        // every CTA has passed barrier m - 1, i.e. nobody reads the accumulator of minibatch m - 2 any more
        if (m >= 2 && slice_t)
            __stcg(reinterpret_cast<float4*>(E.acc + (size_t)((m + 1) % 3) * acc_floats) + my_quad, make_float4(0.f, 0.f, 0.f, 0.f));
        MR_TR(2);
        const tc::TileAcc T = tc::tiles<KP>(C, A, S, MK, m, Q, O);
        if (T.any) {
            tc::stage_partial<KP>(C, T, MK, O, TL, stage);
            umma::fence_proxy_async();   // the staged block (generic writes) -> the bulk reduction's reads
            __syncthreads();
            MR_TR(22);
            if (tid == 0) {
                bulk_reduce_add_f32(accb + (size_t)C.tower * TL.size, stage, (uint32_t)TL.size * 4u);
                bulk_wait_all();
            }
        }
        if (m + 1 < n_mb) MK = tc::mb_const(E.stats, m + 1, A.normalize_adv);   // loads land under the barrier
        MR_TR(3);
        grid_barrier(E.barrier, target, G);
        MR_TR(4);
        ++step;
        b1pow *= (double)E.beta1;   // beta^step, carried from one pow() per launch
        b2pow *= (double)E.beta2;

        const float* src = accb;
        if (E.X.world > 1) {
            // one-shot all-reduce of this CTA's accumulator slice over NVLink peer memory: tagged 8-byte
            // packets {seq, float} into every peer's inbox, summed in rank order (bit-identical everywhere)
            if (slice_t) {
                const float4 mine = __ldcg(reinterpret_cast<const float4*>(accb) + my_quad);
                const float val[4] = {mine.x, mine.y, mine.z, mine.w};
                const unsigned seq = E.seq0 + (unsigned)m + 1u;
                const int slot = seq & 1u;
                const size_t p = 4 * (size_t)my_quad;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned long long pkt = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(val[e]);
                    for (int r = 0; r < E.X.world; ++r)
                        if (r != E.X.rank)
                            __stcg(E.X.peer_inbox[r] + ((size_t)slot * E.X.world + E.X.rank) * acc_floats + p + e, pkt);
                }
                // pull: the four packets of every peer, four peers per wave of 16-byte loads; payloads wait in
                // shared memory.  A peer that never delivers (dead rank) ends the wait after ~4 s: the flag is
                // raised and the epoch finishes on what arrived, so the GPU is released instead of hanging.
                const unsigned long long* src0 = E.X.inbox + (size_t)slot * E.X.world * acc_floats + p;
                unsigned pending = ((1u << E.X.world) - 1u) & ~(1u << E.X.rank);
                long long t0 = 0;
                unsigned spins = 0;
                while (pending) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        ulonglong2 lo[4], hi[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int r = 4 * half + k;
                            if (pending >> r & 1u) {
                                const unsigned long long* q = src0 + (size_t)r * acc_floats;
                                asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo[k].x), "=l"(lo[k].y) : "l"(q) : "memory");
                                asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(hi[k].x), "=l"(hi[k].y) : "l"(q + 2) : "memory");
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int r = 4 * half + k;
                            if ((pending >> r & 1u) && (unsigned)(lo[k].x >> 32) == seq && (unsigned)(lo[k].y >> 32) == seq &&
                                (unsigned)(hi[k].x >> 32) == seq && (unsigned)(hi[k].y >> 32) == seq) {
                                s_x[r * 32 + tid] = make_float4(__uint_as_float((unsigned)lo[k].x), __uint_as_float((unsigned)lo[k].y),
                                                                __uint_as_float((unsigned)hi[k].x), __uint_as_float((unsigned)hi[k].y));
                                pending &= ~(1u << r);
                            }
                        }
                    }
                    if (pending && (++spins & 0xFFFu) == 0) {
                        long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > 4000000000ll) {
                            for (int r = 0; r < E.X.world; ++r)
                                if (pending >> r & 1u) s_x[r * 32 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (E.X.err) *E.X.err = 1;
                            pending = 0;
                        }
                    }
                }
                float tot[4] = {0.f, 0.f, 0.f, 0.f};
                for (int r = 0; r < E.X.world; ++r) {
                    const float4 v = r == E.X.rank ? mine : s_x[r * 32 + tid];
                    tot[0] += v.x; tot[1] += v.y; tot[2] += v.z; tot[3] += v.w;
                }
                __stcg(reinterpret_cast<float4*>(E.gsum) + my_quad, make_float4(tot[0], tot[1], tot[2], tot[3]));
            }
            MR_TR(5);
            grid_barrier(E.barrier, target, G);
            MR_TR(6);
            src = E.gsum;
        }

        // ---- the whole reduced gradient: clip coefficient, then Adam on the own tower -------------------
        tc::BlockRegs gOwn, gOth;
        tc::load_block(src + (size_t)C.tower * TL.size, TL, gOwn);
        tc::load_block(src + (size_t)(C.tower ^ 1) * TL.size, TL, gOth);
        MR_TR(30);
        double sq = tc::finish_block(gOwn, TL, C.tower == 0, E.G.ent_coef) + tc::finish_block(gOth, TL, C.tower != 0, E.G.ent_coef);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) s_sq[warp] = sq;
        __syncthreads();
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < tc::WARPS; ++w) tot += s_sq[w];
        const float total_norm = (float)sqrt(tot);
        tc::AdamK K;
        K.coef = fminf(E.max_grad_norm / (total_norm + 1e-6f), 1.0f);
        K.beta1 = E.beta1; K.beta2 = E.beta2; K.omb1 = 1.f - E.beta1; K.omb2 = 1.f - E.beta2;
        K.neg_step_size = (float)(-(double)E.lr / (1.0 - b1pow));
        K.bc2_sqrt = (float)sqrt(1.0 - b2pow);
        K.eps = E.eps;
        MR_TR(31);
        tc::adam_block(C, gOwn, K, TL, st_w, st_m, st_v);
        MR_TR(7);
        // outputs of the API, off the other CTAs' critical path: statistics row (CTA 0), last gradient (CTAs 0, 1)
        if (c < 2) {
            const int q_st = (TL.st - TL.w1) >> 2;
            if (E.info && c == 0) {
                float* row = E.info + 8 * m;
                if (tid == 0) { row[0] = total_norm; row[1] = K.coef; row[2] = (float)step; row[3] = 0.f; }
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (tid + k * tc::THREADS == q_st) {   // policy: loss, clipped count, kl; value: squared error
                        row[4] = gOwn.r[k].x * inv_cnt; row[5] = gOth.r[k].x * inv_cnt;
                        row[6] = gOwn.r[k].y * inv_cnt; row[7] = gOwn.r[k].z * inv_cnt;
                    }
            }
            if (m == n_mb - 1) {
                const int sb = stat_base(O);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int idx = tid + k * tc::THREADS;
                    const int p = (C.tower ? L.vw2 : L.pw2) + (idx >> 3) * HID + 8 * (idx & 7);
                    const float4 a = gOwn.o[k][0], b = gOwn.o[k][1];
                    const float v8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) E.grad[p + e] = v8[e];
                    const int qi = idx;
                    const float r4[4] = {gOwn.r[k].x, gOwn.r[k].y, gOwn.r[k].z, gOwn.r[k].w};
                    if (qi < ((TL.size - TL.w1) >> 2)) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int pf = tc::tl_to_flat(TL.w1 + 4 * qi + e, C.tower, TL, L, O);
                            if (pf >= 0) E.grad[pf] = r4[e];
                        }
                        if (qi == q_st) {
                            if (C.tower == 0) { E.grad[sb] = r4[0] * inv_cnt; E.grad[sb + 2] = r4[1] * inv_cnt; E.grad[sb + 3] = r4[2] * inv_cnt; }
                            else E.grad[sb + 1] = r4[0] * inv_cnt;
                        }
                    }
                }
            }
        }
        __syncthreads();   // the tower's new fp32 values (other threads' Adam) -> W1 panel, head vectors
        if (m + 1 < n_mb) tc::panel_w1_from_tl<KP>(C, st_w, TL, O);
        MR_TR(37);
    }
    // tower state back to the flat arrays (every CTA of a tower holds the same bits: one writes)
    if (c < 2) {
        for (int i = tid; i < TL.size; i += tc::THREADS) {
            const int p = tc::tl_to_flat(i, C.tower, TL, L, O);
            if (p >= 0) {
                E.params[p] = st_w[i];
                E.exp_avg[p] = st_m[i];
                E.exp_avg_sq[p] = st_v[i];
            }
        }
    }
    if (c == 0 && tid == 0) E.step[0] = step;
    tc::teardown(C);
}

// The tensor-core kernels are the product path; MR_PPO_SIMT=1 selects the fp32 CUDA-core kernels
// (kept as a cross-check of the tcgen05 path in the tests).
static bool use_tc() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MR_PPO_SIMT");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

static size_t grad_smem_bytes(int O, int O_PAD) {
    size_t f = smem_w_floats(O) + 8192 + (size_t)O_PAD * PG_SP + 3 * 128 * PG_SP + PG_S * 4 +
               PG_WARPS * 32 * 16;
    return f * sizeof(float);
}

}  // namespace mr

using namespace mr;

extern "C" {

int mr_ppo_grad_stride(int obs_dim) { return grad_stride(obs_dim); }
int mr_ppo_num_params(int obs_dim) { return make_layout(obs_dim).total; }
int mr_ppo_max_parts(void) { return sm_count(); }
int mr_ppo_epoch_scratch_floats(int obs_dim) { return 4 * 2 * tc::make_tl(obs_dim).size + 64; }

int mr_ppo_adv_stats(const float* adv, const int64_t* perm, int64_t n_samples, int64_t batch_size,
                     int64_t N, int64_t T, double* stats, void* stream) {
    MR_REQUIRE(adv && perm && stats, "NULL argument");
    MR_REQUIRE(batch_size > 0 && n_samples > 0, "empty batch");
    int n_mb = ceil_div(n_samples, batch_size);
    adv_stats_kernel<<<n_mb, 1024, 0, (cudaStream_t)stream>>>(adv, perm, n_samples, batch_size, N, T, stats);
    MR_CHECK_LAUNCH();
    return MR_OK;
}

int mr_device_permutation(uint64_t seed, uint64_t stream_id, int64_t n, int64_t* out, void* stream) {
    MR_REQUIRE(out != nullptr, "NULL argument");
    MR_REQUIRE(n > 0 && n < (int64_t(1) << 31), "n out of range");
    PermKey K;
    K.bits = 1;
    while ((int64_t(1) << K.bits) < n) ++K.bits;
    uint64_t x = seed ^ (stream_id * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
    for (int r = 0; r < 4; ++r) {   // splitmix64 key schedule
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        K.mul[r] = (uint32_t)z | 1u;            // odd: invertible modulo 2^bits
        K.mul[r] = (K.mul[r] & ~6u) | 4u;       // = 5 (mod 8): full-period style multiplier
        K.add[r] = (uint32_t)(z >> 32);
    }
    device_perm_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(out, n, K);
    MR_CHECK_LAUNCH();
    return MR_OK;
}

int mr_ppo_grad_partials(const float* params, int obs_dim, const float* obs, const float* act,
                         const float* old_logp, const float* adv, const float* ret, const int64_t* perm,
                         int32_t* rows, int64_t mb_size, const double* mb_stats, int64_t N, int64_t T,
                         float clip_range, float ent_coef, float vf_coef, int normalize_adv, float* partials,
                         int* n_parts, void* stream) {
    MR_REQUIRE(params && obs && act && old_logp && adv && ret && perm && rows && mb_stats && partials,
               "NULL argument");
    MR_REQUIRE(obs_dim > 0 && obs_dim <= MAX_OBS, "obs_dim out of range");
    MR_REQUIRE(mb_size > 0, "empty minibatch");
    GradArgs A{params, obs, act, old_logp, adv, ret, perm, nullptr, mb_size, mb_stats, N, T,
               clip_range, ent_coef, vf_coef, normalize_adv, partials};
    const int o_pad = obs_dim <= 16 ? 16 : 32;
    const size_t smem = grad_smem_bytes(obs_dim, o_pad);
    static OncePerDevice once;
    if (once.first()) {
        MR_CUDA(cudaFuncSetAttribute(ppo_grad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        MR_CUDA(cudaFuncSetAttribute(ppo_grad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        MR_CUDA(cudaFuncSetAttribute(ppo_grad_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        MR_CUDA(cudaFuncSetAttribute(ppo_grad_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    }
    const int max_parts = mr_ppo_max_parts();
    cudaStream_t s = (cudaStream_t)stream;
    int grid;
    if (use_tc()) {
        MR_REQUIRE(obs_dim < 32 && max_parts >= 2, "tensor-core path needs obs_dim < 32");
        MR_REQUIRE(N * T < (int64_t(1) << 31), "tensor-core path indexes samples with 32 bits");
        perm_to_rows_kernel<<<ceil_div(mb_size, 256), 256, 0, s>>>(perm, mb_size, N, T, rows);
        MR_CHECK_LAUNCH();
        A.rows = rows;
        // obs_dim + 1 (bias column) padded to the 16-bit MMA K of 16
        const int kp = obs_dim + 1 <= 16 ? 16 : 32;
        const int64_t tiles = (mb_size + tc::TILE - 1) / tc::TILE;
        grid = (int)std::min<int64_t>(2 * tiles, max_parts & ~1);
        if (kp == 16) ppo_grad_tc_kernel<16><<<grid, tc::THREADS, tc::SMEM_BYTES, s>>>(A, obs_dim);
        else ppo_grad_tc_kernel<32><<<grid, tc::THREADS, tc::SMEM_BYTES, s>>>(A, obs_dim);
    } else {
        const int64_t tiles = (mb_size + PG_S - 1) / PG_S;
        grid = (int)std::min<int64_t>(tiles, max_parts);
        if (o_pad == 16) ppo_grad_kernel<16><<<grid, PG_THREADS, smem, s>>>(A, obs_dim);
        else ppo_grad_kernel<32><<<grid, PG_THREADS, smem, s>>>(A, obs_dim);
    }
    MR_CHECK_LAUNCH();
    if (n_parts) *n_parts = grid;
    return MR_OK;
}

int mr_ppo_grad(const float* params, int obs_dim, const float* obs, const float* act,
                const float* old_logp, const float* adv, const float* ret, const int64_t* perm,
                int32_t* rows, int64_t mb_size, const double* mb_stats, int64_t N, int64_t T,
                float clip_range, float ent_coef, float vf_coef, int normalize_adv, float rank_share,
                float* partials, float* grad, void* stream) {
    MR_REQUIRE(grad, "NULL argument");
    int grid = 0;
    int rc = mr_ppo_grad_partials(params, obs_dim, obs, act, old_logp, adv, ret, perm, rows, mb_size, mb_stats,
                                  N, T, clip_range, ent_coef, vf_coef, normalize_adv, partials, &grid, stream);
    if (rc != MR_OK) return rc;
    const int stride = grad_stride(obs_dim);
    ppo_reduce_kernel<<<ceil_div(stride, 128), 128, 0, (cudaStream_t)stream>>>(
        partials, grid, obs_dim, ent_coef, mb_stats, grad, rank_share, use_tc() ? 1 : 0);
    MR_CHECK_LAUNCH();
    return MR_OK;
}

int mr_adam_step(float* params, float* exp_avg, float* exp_avg_sq, const float* grad, int n_params,
                 int64_t* step, float lr, float beta1, float beta2, float eps, float max_grad_norm,
                 float* info, void* stream) {
    MR_REQUIRE(params && exp_avg && exp_avg_sq && grad && step, "NULL argument");
    AdamArgs A{params, exp_avg, exp_avg_sq, grad, step, lr, beta1, beta2, eps, max_grad_norm, info, n_params};
    adam_kernel<<<ceil_div(n_params, ADAM_THREADS), ADAM_THREADS, 0, (cudaStream_t)stream>>>(A);
    MR_CHECK_LAUNCH();
    return MR_OK;
}

int mr_ppo_train_epoch(float* params, float* exp_avg, float* exp_avg_sq, int64_t* step, int obs_dim,
                       const float* obs, const float* act, const float* old_logp, const float* adv,
                       const float* ret, const int64_t* perm, int32_t* rows, int64_t n_samples,
                       int64_t batch_size, const double* stats, int64_t N, int64_t T, float clip_range,
                       float ent_coef, float vf_coef, int normalize_adv, float lr, float beta1, float beta2,
                       float eps, float max_grad_norm, float* partials, float* grad, float* info, void* stream) {
    MR_REQUIRE(batch_size > 0 && n_samples > 0, "empty batch");
    const int n_params = make_layout(obs_dim).total;
    const int64_t n_mb = (n_samples + batch_size - 1) / batch_size;
    for (int64_t mb = 0; mb < n_mb; ++mb) {
        const int64_t s0 = mb * batch_size;
        const int64_t sz = std::min(batch_size, n_samples - s0);
        int rc = mr_ppo_grad(params, obs_dim, obs, act, old_logp, adv, ret, perm + s0, rows, sz, stats + 3 * mb,
                             N, T, clip_range, ent_coef, vf_coef, normalize_adv, 1.0f, partials, grad, stream);
        if (rc != MR_OK) return rc;
        rc = mr_adam_step(params, exp_avg, exp_avg_sq, grad, n_params, step, lr, beta1, beta2, eps,
                          max_grad_norm, info ? info + 8 * mb : nullptr, stream);
        if (rc != MR_OK) return rc;
    }
    return MR_OK;
}

#ifdef MR_TRACE
// debug build only: copy out and clear the phase trace of CTA `cta` (0 or 1); returns the count
int mr_trace_read(int cta, unsigned long long* out, int cap) {
    unsigned n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, mr::g_trace_n, sizeof(unsigned), cta * sizeof(unsigned));
    if ((int)n > cap) n = cap;
    cudaMemcpyFromSymbol(out, mr::g_trace, n * sizeof(unsigned long long), (size_t)cta * 4096 * sizeof(unsigned long long));
    unsigned z = 0;
    cudaMemcpyToSymbol(mr::g_trace_n, &z, sizeof(unsigned), cta * sizeof(unsigned));
    return (int)n;
}
#endif

struct mr_xchg {
    int world, rank, device, n_cta;
    int64_t stride;        // packets per (slot, rank): the larger of the two epoch kernels' vectors
    void* base;            // local allocation: inbox packets, then the status word
    size_t bytes;
    void* peer_base[8];
    unsigned seq;          // exchange sequence number (host mirror)
};

static int64_t xchg_stride(int obs_dim) {
    return std::max<int64_t>(grad_stride(obs_dim), 2 * (int64_t)tc::make_tl(obs_dim).size);
}
static size_t xchg_inbox_bytes(int world, int64_t stride) {
    return (size_t)2 * world * stride * sizeof(unsigned long long);
}

int mr_xchg_create(int world, int rank, int device, int obs_dim, mr_xchg** out, uint8_t* h_handle_out) {
    MR_REQUIRE(out && h_handle_out, "NULL argument");
    MR_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "bad world/rank");
    DeviceGuard guard(device);   // the caller's current device is left as it was
    mr_xchg* x = new mr_xchg();
    x->world = world; x->rank = rank; x->device = device;
    x->n_cta = mr_ppo_max_parts();
    x->stride = xchg_stride(obs_dim);
    x->bytes = xchg_inbox_bytes(world, x->stride) + 256;
    x->seq = 0;
    for (int r = 0; r < 8; ++r) x->peer_base[r] = nullptr;
    cudaError_t e = cudaMalloc(&x->base, x->bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); delete x; return MR_ERR_ALLOC; }
    MR_CUDA(cudaMemset(x->base, 0, x->bytes));
    cudaIpcMemHandle_t h;
    MR_CUDA(cudaIpcGetMemHandle(&h, x->base));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(h_handle_out, &h, 64);
    x->peer_base[rank] = x->base;
    *out = x;
    return MR_OK;
}

int mr_xchg_connect(mr_xchg* x, const uint8_t* h_all_handles) {
    MR_REQUIRE(x && h_all_handles, "NULL argument");
    DeviceGuard guard(x->device);
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, h_all_handles + 64 * r, 64);
        MR_CUDA(cudaIpcOpenMemHandle(&x->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    return MR_OK;
}

static int* xchg_status_word(mr_xchg* x) {
    return reinterpret_cast<int*>(static_cast<char*>(x->base) + xchg_inbox_bytes(x->world, x->stride));
}

int mr_xchg_status(mr_xchg* x, int* timed_out) {
    MR_REQUIRE(x && timed_out, "NULL argument");
    DeviceGuard guard(x->device);
    MR_CUDA(cudaMemcpy(timed_out, xchg_status_word(x), sizeof(int), cudaMemcpyDeviceToHost));
    return MR_OK;
}

void mr_xchg_destroy(mr_xchg* x) {
    if (!x) return;
    DeviceGuard guard(x->device);
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->peer_base[r]) cudaIpcCloseMemHandle(x->peer_base[r]);
    cudaFree(x->base);
    delete x;
}

// MR_PPO_EPOCH=v1 selects the round-1 form of the tensor-core epoch kernel (148 partial vectors,
// three grid barriers per minibatch; deterministic summation order) for A/B runs.
static bool use_epoch_v1() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MR_PPO_EPOCH");
        v = (e && e[0] == 'v' && e[1] == '1') ? 1 : 0;
    }
    return v == 1;
}

int mr_ppo_epoch_fused(float* params, float* exp_avg, float* exp_avg_sq, int64_t* step, int obs_dim,
                       const float* obs, const float* act, const float* old_logp, const float* adv,
                       const float* ret, const int64_t* perm, int32_t* rows, int64_t n_samples,
                       int64_t batch_size, const double* stats, const float* rank_share, int64_t N, int64_t T,
                       float clip_range, float ent_coef, float vf_coef, int normalize_adv, float lr,
                       float beta1, float beta2, float eps, float max_grad_norm, float* partials,
                       float* grad, float* info, mr_xchg* xchg, void* stream) {
    MR_REQUIRE(params && exp_avg && exp_avg_sq && step && obs && act && old_logp && adv && ret && perm &&
                   rows && stats && partials && grad, "NULL argument");
    MR_REQUIRE(obs_dim > 0 && obs_dim <= MAX_OBS, "obs_dim out of range");
    MR_REQUIRE(batch_size > 0 && n_samples > 0, "empty batch");
    cudaStream_t s = (cudaStream_t)stream;
    const int n_cta = mr_ppo_max_parts();
    const int stride = grad_stride(obs_dim);
    const bool tcp = use_tc();
    const bool v1 = tcp && use_epoch_v1();
    const tc::TowerLayout TL = tc::make_tl(obs_dim);
    const int acc_floats = 2 * TL.size;
    MR_REQUIRE((int64_t)n_cta * stride >= mr_ppo_epoch_scratch_floats(obs_dim) && stride >= 2 * n_cta + 64,
               "partials scratch too small for the epoch kernel");
    // Everything the launch synchronises through lives in the CALLER's scratch (`partials`,
    // (mr_ppo_max_parts() + 1) rows of grad_stride floats), so two updaters (different streams) never
    // share a barrier word.  Current kernel: [3 accumulators | global sum | barrier word]; v1 / SIMT
    // kernels: one partial vector per CTA in the first n_cta rows, [sq doubles | barrier word] in the
    // extra row.
    EpochArgs E;
    E.G = GradArgs{params, obs, act, old_logp, adv, ret, perm, nullptr, 0, stats, N, T,
                   clip_range, ent_coef, vf_coef, normalize_adv, partials};
    E.n_samples = n_samples; E.batch = batch_size; E.stats = stats; E.rank_share = rank_share;
    E.params = params; E.exp_avg = exp_avg; E.exp_avg_sq = exp_avg_sq; E.step = step;
    E.lr = lr; E.beta1 = beta1; E.beta2 = beta2; E.eps = eps; E.max_grad_norm = max_grad_norm;
    E.grad = grad; E.info = info;
    E.acc = nullptr; E.gsum = nullptr; E.sq = nullptr;
    if (tcp && !v1) {
        E.acc = partials;
        E.gsum = partials + 3 * (size_t)acc_floats;
        E.barrier = reinterpret_cast<unsigned*>(partials + 4 * (size_t)acc_floats);
        MR_CUDA(cudaMemsetAsync(partials, 0, (4 * (size_t)acc_floats + 16) * sizeof(float), s));
        MR_CUDA(cudaMemsetAsync(grad, 0, (size_t)stride * sizeof(float), s));
        MR_REQUIRE((acc_floats / 4 + n_cta - 1) / n_cta <= 32, "accumulator slice per CTA too wide");
    } else {
        float* tail = partials + (size_t)n_cta * stride;   // stride is a multiple of 4: 16-byte aligned
        E.sq = reinterpret_cast<double*>(tail);
        E.barrier = reinterpret_cast<unsigned*>(tail + 2 * (size_t)n_cta + 16);
        MR_CUDA(cudaMemsetAsync(E.barrier, 0, sizeof(unsigned), s));
    }
    E.X.world = 1; E.X.rank = 0; E.X.inbox = nullptr; E.X.err = nullptr;
    E.seq0 = 0;
    const int64_t n_mb = (n_samples + batch_size - 1) / batch_size;
    if (xchg && xchg->world > 1) {
        MR_REQUIRE(xchg->stride == xchg_stride(obs_dim) && xchg->n_cta == n_cta, "exchange buffer mismatch");
        E.X.world = xchg->world; E.X.rank = xchg->rank;
        for (int r = 0; r < xchg->world; ++r)
            E.X.peer_inbox[r] = reinterpret_cast<unsigned long long*>(xchg->peer_base[r]);
        E.X.inbox = E.X.peer_inbox[xchg->rank];
        E.X.err = xchg_status_word(xchg);
        E.seq0 = xchg->seq;
        xchg->seq += (unsigned)n_mb;
    }
    MR_REQUIRE(!tcp || ((n_cta & 1) == 0 && obs_dim < 32), "tensor-core path needs an even CTA count and obs_dim < 32");
    const int o_pad = tcp ? (obs_dim + 1 <= 16 ? 16 : 32) : (obs_dim <= 16 ? 16 : 32);
    const size_t smem_epoch = (size_t)tc::SMEM_BYTES + 3 * (size_t)tc::make_tl(MAX_OBS - 1).size * sizeof(float);
    const size_t smem = tcp ? (v1 ? (size_t)tc::SMEM_BYTES : (size_t)tc::SMEM_BYTES + 3 * (size_t)TL.size * sizeof(float))
                            : grad_smem_bytes(obs_dim, o_pad);
    static OncePerDevice once;
    if (once.first()) {
        MR_CUDA(cudaFuncSetAttribute(ppo_epoch_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        MR_CUDA(cudaFuncSetAttribute(ppo_epoch_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        MR_CUDA(cudaFuncSetAttribute(ppo_epoch_tc_v1_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        MR_CUDA(cudaFuncSetAttribute(ppo_epoch_tc_v1_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        MR_CUDA(cudaFuncSetAttribute(ppo_epoch_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_epoch));
        MR_CUDA(cudaFuncSetAttribute(ppo_epoch_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_epoch));
    }
    if (tcp) {
        MR_REQUIRE(n_samples < (int64_t(1) << 31) && N * T < (int64_t(1) << 31),
                   "tensor-core path indexes samples with 32 bits");
        perm_to_rows_kernel<<<ceil_div(n_samples, 256), 256, 0, s>>>(perm, n_samples, N, T, rows);
        MR_CHECK_LAUNCH();
        E.G.rows = rows;
        if (v1) MR_REQUIRE((((stride + n_cta - 1) / n_cta + 3) >> 2) <= 32, "gradient slice per CTA too wide");
    }
    int O = obs_dim;
    void* args[] = {&E, &O};
    const void* fn = tcp ? (v1 ? (o_pad == 16 ? (const void*)ppo_epoch_tc_v1_kernel<16> : (const void*)ppo_epoch_tc_v1_kernel<32>)
                               : (o_pad == 16 ? (const void*)ppo_epoch_tc_kernel<16> : (const void*)ppo_epoch_tc_kernel<32>))
                         : (o_pad == 16 ? (const void*)ppo_epoch_kernel<16> : (const void*)ppo_epoch_kernel<32>);
    MR_CUDA(cudaLaunchCooperativeKernel(fn, dim3(n_cta), dim3(tcp ? tc::THREADS : PG_THREADS), args, smem, s));
    mr::count_launch();
    return MR_OK;
}

}  // extern "C"
