// Micro-benchmark: how long does an SM wait for data another SM has just written (grid-barrier
// hand-off through L2), as in the gradient reduction / Adam phases of ppo_epoch_tc_kernel?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_handoff l2_handoff.cu && ./l2_handoff
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

constexpr int STRIDE = 10456, SLICE = 72, Q4 = SLICE / 4, THREADS = 256, GROUPS = THREADS / Q4;

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
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

// mode bit 0: writer uses 8-byte strided stores (like the TMEM epilogue) instead of coalesced float4
// mode bit 1: readers use plain (L1-cached) loads instead of ld.cg
// mode bit 2: no rewrite between rounds (data stays clean in L2)
__global__ void __launch_bounds__(THREADS, 1) k(float* partials, float* sink, unsigned* counter, long long* times, int mode, int rounds) {
    const int c = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
    unsigned target = 0;
    float* row = partials + (size_t)c * STRIDE;
    float acc_all = 0.f;
    for (int r = 0; r < rounds; ++r) {
        if (!(mode & 4) || r == 0) {
            if (mode & 1) {
                // lane l < 16 of each warp owns a 64-float row segment, 8-byte stores (16 lanes x 32 stores)
                const int warp = tid >> 5, lane = tid & 31;
                if (lane < 16) {
                    float2* dst = reinterpret_cast<float2*>(row + ((warp * 16 + lane) * 64) % (STRIDE - 64));
                    for (int q = 0; q < 32; ++q) dst[q] = make_float2(r + q, c);
                }
            } else {
                for (int i = tid; i < STRIDE / 4; i += THREADS) reinterpret_cast<float4*>(row)[i] = make_float4(r, c, i, 1.f);
            }
        }
        grid_barrier(counter, target, G);
        long long t0 = clock64();
        const int j = tid % Q4, g = tid / Q4;
        float4 acc = make_float4(0, 0, 0, 0);
        if (g < GROUPS) {
            const float* src = partials + c * SLICE + 4 * j;
            for (int b = g; b < G; b += GROUPS) {
                float4 v;
                if (mode & 2) v = __ldca(reinterpret_cast<const float4*>(src + (size_t)b * STRIDE));
                else v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)b * STRIDE));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        acc_all += acc.x + acc.y + acc.z + acc.w;
        long long t1 = clock64();
        // second pass over the same lines
        acc = make_float4(0, 0, 0, 0);
        if (g < GROUPS) {
            const float* src = partials + c * SLICE + 4 * j;
            for (int b = g; b < G; b += GROUPS) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)b * STRIDE));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        acc_all += acc.x + acc.y + acc.z + acc.w;
        long long t2 = clock64();
        if (tid == 0 && c < 4) { times[(c * rounds + r) * 2] = t1 - t0; times[(c * rounds + r) * 2 + 1] = t2 - t1; }
        grid_barrier(counter, target, G);
    }
    sink[c * THREADS + tid] = acc_all;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *partials, *sink; unsigned* counter; long long* times;
    const int rounds = 20;
    cudaMalloc(&partials, (size_t)sms * STRIDE * 4); cudaMemset(partials, 0, (size_t)sms * STRIDE * 4);
    cudaMalloc(&sink, sms * THREADS * 4); cudaMalloc(&counter, 4); cudaMalloc(&times, 4 * rounds * 2 * 8);
    long long h[4 * rounds * 2];
    for (int mode = 0; mode < 8; ++mode) {
        cudaMemset(counter, 0, 4);
        int rr = rounds;
        void* args[] = {&partials, &sink, &counter, &times, &mode, &rr};
        cudaLaunchCooperativeKernel((void*)k, dim3(sms), dim3(THREADS), args, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, times, sizeof(h), cudaMemcpyDeviceToHost);
        double a = 0, b = 0; int n = 0;
        for (int c = 0; c < 4; ++c) for (int r = 5; r < rounds; ++r) { a += h[(c * rounds + r) * 2]; b += h[(c * rounds + r) * 2 + 1]; ++n; }
        printf("mode %d (%s stores, %s loads, %s): first pass %.0f cyc, second pass %.0f cyc  (thread 0 of CTA 0-3, %d float4 loads each)\n", mode,
               (mode & 1) ? "8B strided" : "coalesced 16B", (mode & 2) ? "plain" : "ld.cg", (mode & 4) ? "clean data" : "rewritten every round", a / n, b / n, (sms + GROUPS - 1) / GROUPS);
    }
    return 0;
}
