// Stand-alone batched policy / value forward (K4, K5): one warp per tile of 8 samples,
// parameters staged once per block in shared memory, grid sized to the SM count.
#include "mlp.cuh"

namespace mr {

constexpr int PF_WARPS = 4;
constexpr int PF_E = 8;

__global__ void __launch_bounds__(PF_WARPS * 32)
policy_forward_kernel(const float* __restrict__ params, int O, const float* __restrict__ obs,
                      const float* __restrict__ eps, float* __restrict__ act,
                      float* __restrict__ logp, float* __restrict__ val, int64_t n) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SmemW W = stage_weights(smem, params, O);
    float* obsT = smem + smem_w_floats(O) + warp * (MAX_OBS * PF_E + 128 * PF_E);
    float* hbuf = obsT + MAX_OBS * PF_E;
    __syncthreads();
    const float sig0 = expf(W.logstd[0]), sig1 = expf(W.logstd[1]);
    const int64_t n_tiles = (n + PF_E - 1) / PF_E;
    for (int64_t tile = (int64_t)blockIdx.x * PF_WARPS + warp; tile < n_tiles;
         tile += (int64_t)gridDim.x * PF_WARPS) {
        const int64_t s0 = tile * PF_E;
        const int rows = (int)min((int64_t)PF_E, n - s0);
        // rows of the tile are contiguous in obs: coalesced read, transposed into obsT[k][e]
        for (int idx = lane; idx < PF_E * O; idx += 32) {
            int e = idx / O, k = idx - e * O;
            obsT[k * PF_E + e] = e < rows ? obs[s0 * O + idx] : 0.f;
        }
        __syncwarp();
        float out = warp_mlp_forward<PF_E>(W, O, obsT, hbuf, lane);
        const int e = lane / 3, j = lane - 3 * e;
        const bool live = lane < 3 * PF_E && e < rows;
        float a = out, lp = 0.f;
        if (live && j < 2) {
            const float sig = j == 0 ? sig0 : sig1;
            if (eps) a = __fadd_rn(out, __fmul_rn(sig, eps[(s0 + e) * 2 + j]));
            lp = normal_logprob(a, out, sig);
            act[(s0 + e) * 2 + j] = a;
        }
        float lp1 = __shfl_down_sync(0xffffffffu, lp, 1);
        if (live && j == 0 && logp) logp[s0 + e] = __fadd_rn(lp, lp1);
        if (live && j == 2 && val) val[s0 + e] = out;
        __syncwarp();
    }
}

}  // namespace mr

using namespace mr;

extern "C" int mr_policy_forward(const float* params, int obs_dim, const float* obs,
                                 const float* eps, float* act, float* logp, float* val,
                                 int64_t n, void* stream) {
    MR_REQUIRE(params && obs && act, "NULL argument");
    MR_REQUIRE(obs_dim > 0 && obs_dim <= MAX_OBS, "obs_dim out of range");
    if (n <= 0) return MR_OK;
    size_t smem = (smem_w_floats(obs_dim) + PF_WARPS * (MAX_OBS * PF_E + 128 * PF_E)) * sizeof(float);
    static OncePerDevice once;
    if (once.first())
        MR_CUDA(cudaFuncSetAttribute(policy_forward_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const int sms = sm_count();
    int64_t tiles = (n + PF_E - 1) / PF_E;
    int blocks = (int)std::min<int64_t>((tiles + PF_WARPS - 1) / PF_WARPS, (int64_t)sms * 2);
    policy_forward_kernel<<<blocks, PF_WARPS * 32, smem, (cudaStream_t)stream>>>(
        params, obs_dim, obs, eps, act, logp, val, n);
    MR_CHECK_LAUNCH();
    return MR_OK;
}
