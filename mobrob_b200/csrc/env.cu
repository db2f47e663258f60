// Batched goal environment resident in HBM: state slab, stand-alone VecEnv kernels
// (step / reset / obs / state import-export).  One thread per environment, SoA state,
// observation rows staged through shared memory so the [N][O] store is fully coalesced.
#include "env_state.cuh"

#include <stdarg.h>

#include <atomic>

namespace mr {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------
constexpr int STEP_THREADS = 128;
constexpr int OBS_PAD = point::OBS + 1;  // 15-float rows: conflict-free thread-per-row writes

__device__ __forceinline__ void store_rows_coalesced(float* __restrict__ dst, const float* smem,
                                                     int64_t block_start, int rows, int64_t n) {
    // dst rows [block_start, block_start + rows) are one contiguous span of rows * OBS floats
    const int total = rows * point::OBS;
    float* base = dst + block_start * point::OBS;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int r = idx / point::OBS, k = idx - r * point::OBS;
        base[idx] = smem[r * OBS_PAD + k];
    }
}

__global__ void __launch_bounds__(STEP_THREADS)
point_step_kernel(PointState st, EnvCfg cfg, const float2* __restrict__ act,
                  float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ done,
                  uint8_t* __restrict__ trunc, float* __restrict__ term_obs,
                  double* __restrict__ ep_ret, int32_t* __restrict__ ep_len) {
    __shared__ float s_obs[STEP_THREADS * OBS_PAD];
    const int64_t block_start = (int64_t)blockIdx.x * STEP_THREADS;
    const int64_t i = block_start + threadIdx.x;
    if (i < st.n) {
        PointHot h = st.load(i);
        float2 a = act[i];
        float tobs[point::OBS];
        StepResult r = point_env_step(h, st.cold, i, a.x, a.y, cfg,
                                      &s_obs[threadIdx.x * OBS_PAD], tobs);
        st.store(i, h);
        rew[i] = r.rew;
        done[i] = r.done ? 1 : 0;
        trunc[i] = r.trunc ? 1 : 0;
        if (r.done) {
            if (term_obs) {
#pragma unroll
                for (int k = 0; k < point::OBS; ++k) term_obs[i * point::OBS + k] = tobs[k];
            }
            if (ep_ret) ep_ret[i] = r.ep_r;
            if (ep_len) ep_len[i] = r.ep_l;
        }
    }
    __syncthreads();
    int rows = (int)min((int64_t)STEP_THREADS, st.n - block_start);
    store_rows_coalesced(obs, s_obs, block_start, rows, st.n);
}

__global__ void __launch_bounds__(STEP_THREADS)
point_reset_kernel(PointState st, const uint8_t* __restrict__ mask, int first,
                   float* __restrict__ obs) {
    const int64_t i = (int64_t)blockIdx.x * STEP_THREADS + threadIdx.x;
    if (i >= st.n) return;
    if (mask && !mask[i]) return;
    PointHot h = st.load(i);
    bool reach = point::dist2((double)h.gx, (double)h.gy, h.d.px, h.d.py) < REACH_RADIUS;
    point_reset(h, st.cold, i, first || !reach);
    st.store(i, h);
    if (obs) {
        float o[point::OBS];
        point::sensors(h.d, (double)h.cx, (double)h.cz, h.gx, h.gy, o);
#pragma unroll
        for (int k = 0; k < point::OBS; ++k) obs[i * point::OBS + k] = o[k];
    }
}

__global__ void __launch_bounds__(STEP_THREADS)
point_obs_kernel(PointState st, float* __restrict__ obs) {
    __shared__ float s_obs[STEP_THREADS * OBS_PAD];
    const int64_t block_start = (int64_t)blockIdx.x * STEP_THREADS;
    const int64_t i = block_start + threadIdx.x;
    if (i < st.n) {
        PointHot h = st.load(i);
        point::sensors(h.d, (double)h.cx, (double)h.cz, h.gx, h.gy,
                       &s_obs[threadIdx.x * OBS_PAD]);
    }
    __syncthreads();
    int rows = (int)min((int64_t)STEP_THREADS, st.n - block_start);
    store_rows_coalesced(obs, s_obs, block_start, rows, st.n);
}

// reference view: qpos(3) qvel(3) body_xy(2) psi0(1) ctrl(2) goal(2) elapsed(1) ep_ret(1)
__global__ void point_get_state_kernel(PointState st, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n) return;
    PointHot h = st.load(i);
    float2 b = st.cold.body_xy[i];
    double p0 = st.cold.psi0[i], s0, c0;
    sincos(p0, &s0, &c0);
    double dx = h.d.px - (double)b.x, dy = h.d.py - (double)b.y;
    double* o = out + i * POINT_STATE_DIM;
    o[0] = c0 * dx + s0 * dy;
    o[1] = -s0 * dx + c0 * dy;
    o[2] = h.d.psi - p0;
    o[3] = c0 * h.d.vx + s0 * h.d.vy;
    o[4] = -s0 * h.d.vx + c0 * h.d.vy;
    o[5] = h.d.om;
    o[6] = b.x; o[7] = b.y; o[8] = p0;
    o[9] = h.cx; o[10] = h.cz; o[11] = h.gx; o[12] = h.gy;
    o[13] = (double)h.elapsed; o[14] = h.ep_ret;
}

__global__ void point_set_state_kernel(PointState st, const double* __restrict__ in) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n) return;
    const double* s = in + i * POINT_STATE_DIM;
    double p0 = s[8], s0, c0;
    sincos(p0, &s0, &c0);
    PointHot h;
    h.d.px = s[6] + c0 * s[0] - s0 * s[1];
    h.d.py = s[7] + s0 * s[0] + c0 * s[1];
    h.d.psi = p0 + s[2];
    h.d.vx = c0 * s[3] - s0 * s[4];
    h.d.vy = s0 * s[3] + c0 * s[4];
    h.d.om = s[5];
    h.cx = (float)s[9]; h.cz = (float)s[10]; h.gx = (float)s[11]; h.gy = (float)s[12];
    h.elapsed = (int)s[13]; h.ep_ret = s[14];
    st.cold.body_xy[i] = make_float2((float)s[6], (float)s[7]);
    st.cold.psi0[i] = p0;
    st.store(i, h);
}

__global__ void point_get_pos_kernel(PointState st, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n) return;
    out[2 * i] = st.px[i];
    out[2 * i + 1] = st.py[i];
}

}  // namespace mr

using namespace mr;

// =======================================================================================
extern "C" {

int mr_version(void) { return 100; }
const char* mr_last_error(void) { return mr::g_err; }
uint64_t mr_launch_count(void) { return mr::g_launches.load(); }

int mr_env_create(int kind, int64_t n_envs, int device, int time_limit, int terminate_on_goal,
                  mr_env** out) {
    MR_REQUIRE(out != nullptr, "out is NULL");
    MR_REQUIRE(n_envs > 0, "n_envs must be positive");
    if (kind != MR_ENV_POINT) {
        set_error("env kind %d is not built yet (point only)", kind);
        return MR_ERR_UNSUPPORTED;
    }
    MR_CUDA(cudaSetDevice(device));
    mr_env* e = new mr_env();
    e->kind = kind;
    e->n = n_envs;
    e->device = device;
    e->cfg.time_limit = time_limit;
    e->cfg.terminate_on_goal = terminate_on_goal ? 1 : 0;
    size_t bytes = PointState::slab_bytes(n_envs);
    cudaError_t err = cudaMalloc(&e->slab, bytes);
    if (err != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(err));
        delete e;
        return MR_ERR_ALLOC;
    }
    cudaMemset(e->slab, 0, bytes);
    e->slab_bytes = bytes;
    e->point.carve(e->slab, n_envs);
    *out = e;
    return MR_OK;
}

void mr_env_destroy(mr_env* env) {
    if (!env) return;
    cudaSetDevice(env->device);
    cudaFree(env->slab);
    delete env;
}

int mr_env_obs_dim(const mr_env* env) { return env ? point::OBS : 0; }
int mr_env_state_dim(const mr_env* env) { return env ? POINT_STATE_DIM : 0; }

int mr_env_seed(mr_env* env, const uint64_t* h_pcg_init, const uint64_t* h_pcg_goal,
                const int64_t* h_engine_seed, void* stream) {
    MR_REQUIRE(env && h_pcg_init && h_pcg_goal && h_engine_seed, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    MR_CUDA(cudaSetDevice(env->device));
    const EnvCold& c = env->point.cold;
    MR_CUDA(cudaMemcpyAsync(c.pcg_init, h_pcg_init, env->n * 32, cudaMemcpyHostToDevice, s));
    MR_CUDA(cudaMemcpyAsync(c.pcg_goal, h_pcg_goal, env->n * 32, cudaMemcpyHostToDevice, s));
    MR_CUDA(cudaMemcpyAsync(c.engine_seed, h_engine_seed, env->n * 8, cudaMemcpyHostToDevice, s));
    MR_CUDA(cudaStreamSynchronize(s));  // host buffers may be freed by the caller on return
    return MR_OK;
}

int mr_env_reset(mr_env* env, const uint8_t* mask, int first, float* obs_out, void* stream) {
    MR_REQUIRE(env, "env is NULL");
    point_reset_kernel<<<ceil_div(env->n, STEP_THREADS), STEP_THREADS, 0, (cudaStream_t)stream>>>(
        env->point, mask, first, obs_out);
    MR_CHECK_LAUNCH();
    return MR_OK;
}

int mr_env_step(mr_env* env, const float* act, float* obs, float* rew, uint8_t* done,
                uint8_t* trunc, float* term_obs, double* ep_ret, int32_t* ep_len, void* stream) {
    MR_REQUIRE(env && act && obs && rew && done && trunc, "NULL argument");
    point_step_kernel<<<ceil_div(env->n, STEP_THREADS), STEP_THREADS, 0, (cudaStream_t)stream>>>(
        env->point, env->cfg, reinterpret_cast<const float2*>(act), obs, rew, done, trunc,
        term_obs, ep_ret, ep_len);
    MR_CHECK_LAUNCH();
    return MR_OK;
}

int mr_env_get_obs(mr_env* env, float* obs_out, void* stream) {
    MR_REQUIRE(env && obs_out, "NULL argument");
    point_obs_kernel<<<ceil_div(env->n, STEP_THREADS), STEP_THREADS, 0, (cudaStream_t)stream>>>(
        env->point, obs_out);
    MR_CHECK_LAUNCH();
    return MR_OK;
}

int mr_env_get_state(mr_env* env, double* state_out, void* stream) {
    MR_REQUIRE(env && state_out, "NULL argument");
    point_get_state_kernel<<<ceil_div(env->n, 128), 128, 0, (cudaStream_t)stream>>>(env->point,
                                                                                   state_out);
    MR_CHECK_LAUNCH();
    return MR_OK;
}

int mr_env_set_state(mr_env* env, const double* state_in, void* stream) {
    MR_REQUIRE(env && state_in, "NULL argument");
    point_set_state_kernel<<<ceil_div(env->n, 128), 128, 0, (cudaStream_t)stream>>>(env->point,
                                                                                   state_in);
    MR_CHECK_LAUNCH();
    return MR_OK;
}

int mr_env_get_pos(mr_env* env, double* pos_out, void* stream) {
    MR_REQUIRE(env && pos_out, "NULL argument");
    point_get_pos_kernel<<<ceil_div(env->n, 128), 128, 0, (cudaStream_t)stream>>>(env->point,
                                                                                 pos_out);
    MR_CHECK_LAUNCH();
    return MR_OK;
}

int mr_env_get_reset_counts(mr_env* env, int32_t* out, void* stream) {
    MR_REQUIRE(env && out, "NULL argument");
    MR_CUDA(cudaMemcpyAsync(out, env->point.cold.counts, env->n * 8, cudaMemcpyDeviceToDevice,
                            (cudaStream_t)stream));
    return MR_OK;
}

}  // extern "C"
