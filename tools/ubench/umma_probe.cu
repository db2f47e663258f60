// Probe for the tcgen05 building blocks of the PPO update kernel (run on a B200 via gpurun):
// checks every operand layout / major-mode combination the kernel relies on against a CPU
// product, measures the 3xTF32 error on random fp32 data and times MMA sequences.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && ./umma_probe
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <cuda_bf16.h>

#include "../../mobrob_b200/csrc/umma.cuh"

using namespace mr::umma;

struct ProbeCfg {
    int M, N, K;
    int a_mn, b_mn;     // 1 = MN-major operand
    int swap_fields;    // MN-major: put the panel stride in SBO instead of LBO
    int split;          // 1 = 3xTF32 (hi/lo buffers, three MMAs per k-step)
    int reps;           // repeat the whole MMA sequence (timing)
    int order;          // split order: 0 = small terms first, 1 = hi*hi first
    int la, lb;         // operand layout: 0 = 128B swizzle panels, 1 = no swizzle (8x16B core matrices),
                        // 2 = 128B swizzle with 32B atoms (MN-major only)
};

__device__ __forceinline__ uint64_t desc_raw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)type << 61;
    return d;
}

// generic operand placement: R = extent along M/N, K = extent along K
__device__ __forceinline__ uint32_t elem_off2(int layout, int mn, int r, int k, int R, int K) {
    if (layout == 0) {
        if (!mn) return (uint32_t)(k >> 5) * R * 128u + panel_off(r, k & 31);
        return (uint32_t)(r >> 5) * K * 128u + panel_off(k, r & 31);
    }
    if (layout == 1) {
        // core matrix = 8 rows of 16 B.  K-major: rows = r, 4 k per row; MN-major: rows = k, 4 r per row.
        // cores of consecutive 4-element chunks are adjacent (128 B); 8-row groups follow.
        if (!mn) return (uint32_t)(r >> 3) * ((K >> 2) * 128u) + (uint32_t)(k >> 2) * 128u + (r & 7) * 16u + (k & 3) * 4u;
        return (uint32_t)(k >> 3) * ((R >> 2) * 128u) + (uint32_t)(r >> 2) * 128u + (k & 7) * 16u + (r & 3) * 4u;
    }
    // layout 2 (MN-major): panel of 32 r, rows = k of 128 B, 32-byte chunks XOR (k & 3)
    return (uint32_t)(r >> 5) * K * 128u + (uint32_t)k * 128u + (((((uint32_t)r & 31) >> 3) ^ ((uint32_t)k & 3)) << 5) +
           (r & 7) * 4u;
}
__device__ __forceinline__ uint64_t op_desc(int layout, int mn, uint32_t base, int ks, int R, int K) {
    if (layout == 0) {
        if (!mn) return desc_kmajor(base + (ks >> 2) * R * 128, ks & 3);
        return desc_mnmajor(base, ks, K * 128);
    }
    if (layout == 1) {
        if (!mn) return desc_raw(base + ks * 256, /*LBO: k chunk*/ 128, /*SBO: r group*/ (K >> 2) * 128, 0);
        return desc_raw(base + ks * ((R >> 2) * 128), /*LBO: k group*/ (R >> 2) * 128, /*SBO: r chunk*/ 128, 0);
    }
    return desc_raw(base + ks * 1024, /*LBO: panel*/ K * 128, /*SBO: 4-row group*/ 512, 1);
}

__device__ __forceinline__ uint32_t elem_off(int mn, int r, int k, int rows_k_major, int rows_mn_major) {
    // r = M/N index, k = K index
    if (!mn) return (uint32_t)(k >> 5) * rows_k_major * 128u + panel_off(r, k & 31);
    return (uint32_t)(r >> 5) * rows_mn_major * 128u + panel_off(k, r & 31);
}

__global__ void __launch_bounds__(128) probe_kernel(ProbeCfg c, const float* __restrict__ A,
                                                    const float* __restrict__ B, float* __restrict__ D,
                                                    long long* __restrict__ cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int a_panels = c.a_mn ? c.M / 32 : (c.K + 31) / 32, a_rows = c.a_mn ? c.K : c.M;
    const int b_panels = c.b_mn ? (c.N + 31) / 32 : (c.K + 31) / 32, b_rows = c.b_mn ? c.K : c.N;
    const uint32_t a_bytes = (a_panels * a_rows * 128 + 1023) & ~1023u, b_bytes = (b_panels * b_rows * 128 + 1023) & ~1023u;
    uint8_t* a_hi = base;
    uint8_t* a_lo = a_hi + a_bytes;
    uint8_t* b_hi = a_lo + a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;
    for (uint32_t i = tid; i < (2 * a_bytes + 2 * b_bytes) / 4; i += blockDim.x) ((float*)base)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < c.M * c.K; i += blockDim.x) {
        int m = i / c.K, k = i - m * c.K;
        float x = A[i], hi = x, lo = 0.f;
        if (c.split) split_tf32(x, hi, lo);
        uint32_t off = elem_off2(c.la, c.a_mn, m, k, c.M, c.K);
        *(float*)(a_hi + off) = hi;
        *(float*)(a_lo + off) = lo;
    }
    for (int i = tid; i < c.N * c.K; i += blockDim.x) {
        int n = i / c.K, k = i - n * c.K;
        float x = B[i], hi = x, lo = 0.f;
        if (c.split) split_tf32(x, hi, lo);
        uint32_t off = elem_off2(c.lb, c.b_mn, n, k, c.N, c.K);
        *(float*)(b_hi + off) = hi;
        *(float*)(b_lo + off) = lo;
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_slot;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(c.M, c.N, c.a_mn, c.b_mn);
        auto adesc = [&](uint8_t* buf, int ks) {
            if (c.la == 0 && c.a_mn && c.swap_fields) return smem_desc(smem_u32(buf) + ks * 1024, 1024, a_rows * 128);
            return op_desc(c.la, c.a_mn, smem_u32(buf), ks, c.M, c.K);
        };
        auto bdesc = [&](uint8_t* buf, int ks) {
            if (c.lb == 0 && c.b_mn && c.swap_fields) return smem_desc(smem_u32(buf) + ks * 1024, 1024, b_rows * 128);
            return op_desc(c.lb, c.b_mn, smem_u32(buf), ks, c.N, c.K);
        };
        t0 = clock64();
        for (int rep = 0; rep < c.reps; ++rep) {
            bool acc = false;
            const int nk = c.K / 8;
            if (!c.split) {
                for (int ks = 0; ks < nk; ++ks) { mma_tf32(tmem, adesc(a_hi, ks), bdesc(b_hi, ks), idesc, acc); acc = true; }
            } else if (c.order == 0) {
                for (int ks = 0; ks < nk; ++ks) { mma_tf32(tmem, adesc(a_lo, ks), bdesc(b_hi, ks), idesc, acc); acc = true; }
                for (int ks = 0; ks < nk; ++ks) mma_tf32(tmem, adesc(a_hi, ks), bdesc(b_lo, ks), idesc, true);
                for (int ks = 0; ks < nk; ++ks) mma_tf32(tmem, adesc(a_hi, ks), bdesc(b_hi, ks), idesc, true);
            } else {
                for (int ks = 0; ks < nk; ++ks) {
                    mma_tf32(tmem, adesc(a_hi, ks), bdesc(b_hi, ks), idesc, acc); acc = true;
                    mma_tf32(tmem, adesc(a_lo, ks), bdesc(b_hi, ks), idesc, true);
                    mma_tf32(tmem, adesc(a_hi, ks), bdesc(b_lo, ks), idesc, true);
                }
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
    fence_after_sync();
    for (int c0 = 0; c0 < c.N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 32; ++j)
            if (c0 + j < c.N) D[(size_t)tid * c.N + c0 + j] = v[j];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static double run(const char* name, ProbeCfg c, bool ints, unsigned seed, bool verbose_rows = false) {
    std::vector<float> A((size_t)c.M * c.K), B((size_t)c.N * c.K), D((size_t)128 * c.N, -777.f);
    srand(seed);
    auto rnd = [&]() {
        if (ints) return (float)((rand() % 9) - 4);
        return (float)((rand() / (double)RAND_MAX) * 2.0 - 1.0);
    };
    for (auto& x : A) x = rnd();
    for (auto& x : B) x = rnd();
    float *dA, *dB, *dD;
    long long* dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(c, dA, dB, dD, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%-34s CUDA ERROR %s\n", name, cudaGetErrorString(e));
        exit(1);
    }
    long long cyc = 0;
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0;
    int bad = 0;
    for (int m = 0; m < c.M; ++m) {
        const int lane = c.M == 128 ? m : (m % 16) + 32 * (m / 16);
        for (int n = 0; n < c.N; ++n) {
            double ref = 0;
            for (int k = 0; k < c.K; ++k) ref += (double)A[(size_t)m * c.K + k] * (double)B[(size_t)n * c.K + k];
            double err = fabs(ref - (double)D[(size_t)lane * c.N + n]);
            if (err > max_err) max_err = err;
            if (fabs(ref) > max_ref) max_ref = fabs(ref);
            if (err > 1e-3 * (1 + fabs(ref)) && bad < 4 && verbose_rows) {
                printf("   mismatch m=%d n=%d ref=%g got=%g\n", m, n, ref, D[(size_t)lane * c.N + n]);
                ++bad;
            }
        }
    }
    const int n_mma = c.reps * (c.K / 8) * (c.split ? 3 : 1);
    printf("%-34s M=%3d N=%3d K=%3d a_mn=%d b_mn=%d la=%d lb=%d swap=%d split=%d: max_err=%.3e (rel %.2e)  %lld cyc / %d mma = %.1f\n",
           name, c.M, c.N, c.K, c.a_mn, c.b_mn, c.la, c.lb, c.swap_fields, c.split, max_err, max_err / (max_ref + 1e-30), cyc, n_mma,
           (double)cyc / n_mma);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
    return max_err;
}

// ---------------------------------------------------------------------------------------------------
// kind::f16 with bf16 operands split three ways (x = b0 + b1 + b2, 8 significand bits each): 16-bit
// operands may be K-major or MN-major under the SAME 128-byte swizzle, so one buffer serves both.
struct Probe16 {
    int M, N, K;          // K multiple of 16
    int a_mn, b_mn;
    int terms;            // 1, 3, 5 or 6 products of the 3x3 expansion (largest first in magnitude order)
    int reps;
    int n_off;            // MN-major B: column offset of the N window inside the 64-wide panel
};

__device__ __forceinline__ uint32_t off16(int mn, int r, int k, int R, int K) {
    // panel = 64 contiguous elements (128 B) per row
    int row = mn ? k : r, col = mn ? r : k, rows = mn ? K : R;
    return (uint32_t)(col >> 6) * rows * 128u + (uint32_t)row * 128u + ((((uint32_t)(col & 63) >> 3) ^ (uint32_t)row) & 7u) * 16u +
           (col & 7) * 2u;
}
__device__ __forceinline__ uint64_t desc16(int mn, uint32_t base, int ks, int R, int K) {
    if (!mn) return smem_desc(base + (ks >> 2) * R * 128 + (ks & 3) * 32, 16, 1024);
    return smem_desc(base + ks * 2048, K * 128, 1024);
}
__device__ __forceinline__ void split3(float x, __nv_bfloat16& b0, __nv_bfloat16& b1, __nv_bfloat16& b2) {
    b0 = __float2bfloat16_rn(x);
    float r1 = x - __bfloat162float(b0);
    b1 = __float2bfloat16_rn(r1);
    float r2 = r1 - __bfloat162float(b1);
    b2 = __float2bfloat16_rn(r2);
}

__global__ void __launch_bounds__(128) probe16_kernel(Probe16 c, const float* __restrict__ A, const float* __restrict__ B,
                                                      float* __restrict__ D, long long* __restrict__ cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int NB = c.b_mn ? 64 : c.N;   // MN-major B lives in a full 64-wide panel
    const uint32_t a_bytes = ((c.a_mn ? (c.M + 63) / 64 * c.K : (c.K + 63) / 64 * c.M) * 128 + 1023) & ~1023u;
    const uint32_t b_bytes = ((c.b_mn ? c.K : (c.K + 63) / 64 * c.N) * 128 + 1023) & ~1023u;
    uint8_t* a[3] = {base, base + a_bytes, base + 2 * a_bytes};
    uint8_t* b[3] = {base + 3 * a_bytes, base + 3 * a_bytes + b_bytes, base + 3 * a_bytes + 2 * b_bytes};
    for (uint32_t i = tid; i < (3 * a_bytes + 3 * b_bytes) / 4; i += blockDim.x) ((float*)base)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < c.M * c.K; i += blockDim.x) {
        int m = i / c.K, k = i - m * c.K;
        __nv_bfloat16 x[3];
        split3(A[i], x[0], x[1], x[2]);
        uint32_t off = off16(c.a_mn, m, k, c.M, c.K);
        for (int q = 0; q < 3; ++q) *(__nv_bfloat16*)(a[q] + off) = x[q];
    }
    for (int i = tid; i < c.N * c.K; i += blockDim.x) {
        int n = i / c.K, k = i - n * c.K;
        __nv_bfloat16 x[3];
        split3(B[i], x[0], x[1], x[2]);
        uint32_t off = off16(c.b_mn, n + (c.b_mn ? c.n_off : 0), k, NB, c.K);
        for (int q = 0; q < 3; ++q) *(__nv_bfloat16*)(b[q] + off) = x[q];
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // provably warp-uniform
    if (warp_u == 0) {
        // the whole warp runs the issue loop (uniform datapath); one elected lane issues
        const bool leader = elect_one();
        // kind::f16: a/b format 1 = bf16, accumulate fp32
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                               ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
        const int nk = c.K / 16;
        uint64_t ad[3], bd[3];
        for (int q = 0; q < 3; ++q) {
            ad[q] = desc16(c.a_mn, smem_u32(a[q]), 0, c.M, c.K);
            bd[q] = desc16(c.b_mn, smem_u32(b[q]) + (c.b_mn ? c.n_off * 2 : 0), 0, NB, c.K);
        }
        const uint32_t a_inc = c.a_mn ? 2048 / 16 : 32 / 16, b_inc = c.b_mn ? 2048 / 16 : 32 / 16;
        const int pi[6] = {2, 0, 1, 1, 0, 0}, pj[6] = {0, 2, 1, 0, 1, 0};
        long long t0 = clock64();
        for (int rep = 0; rep < c.reps; ++rep) {
            uint32_t acc = 0;
            for (int t = 6 - c.terms; t < 6; ++t) {
                uint64_t da = ad[pi[t]], db = bd[pj[t]];
                for (int ks = 0; ks < nk; ++ks) {
                    if (leader) mma_f16(tmem, da, db, idesc, acc);
                    acc = 1;
                    da += a_inc; db += b_inc;
                }
            }
        }
        if (leader) mma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        if (leader) cycles[0] = clock64() - t0;
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c0 = 0; c0 < c.N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 32; ++j)
            if (c0 + j < c.N) D[(size_t)tid * c.N + c0 + j] = v[j];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static double run16(const char* name, Probe16 c, bool ints, unsigned seed, bool verbose_rows = false) {
    std::vector<float> A((size_t)c.M * c.K), B((size_t)c.N * c.K), D((size_t)128 * c.N, -777.f);
    srand(seed);
    auto rnd = [&]() {
        if (ints) return (float)((rand() % 9) - 4);
        return (float)((rand() / (double)RAND_MAX) * 2.0 - 1.0);
    };
    for (auto& x : A) x = rnd();
    for (auto& x : B) x = rnd();
    float *dA, *dB, *dD;
    long long* dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(probe16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe16_kernel<<<1, 128, smem>>>(c, dA, dB, dD, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-34s CUDA ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
    long long cyc = 0;
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0, sum_sq = 0;
    int bad = 0;
    for (int m = 0; m < c.M; ++m) {
        const int lane = c.M == 128 ? m : (m % 16) + 32 * (m / 16);
        for (int n = 0; n < c.N; ++n) {
            double ref = 0;
            for (int k = 0; k < c.K; ++k) ref += (double)A[(size_t)m * c.K + k] * (double)B[(size_t)n * c.K + k];
            double err = fabs(ref - (double)D[(size_t)lane * c.N + n]);
            sum_sq += err * err;
            if (err > max_err) max_err = err;
            if (fabs(ref) > max_ref) max_ref = fabs(ref);
            if (err > 1e-3 * (1 + fabs(ref)) && bad < 4 && verbose_rows) {
                printf("   mismatch m=%d n=%d ref=%g got=%g\n", m, n, ref, D[(size_t)lane * c.N + n]);
                ++bad;
            }
        }
    }
    const int n_mma = c.reps * (c.K / 16) * c.terms;
    printf("%-34s M=%3d N=%3d K=%3d a_mn=%d b_mn=%d terms=%d noff=%d: max_err=%.3e rms=%.2e (max rel %.2e)  %lld cyc / %d mma = %.1f\n",
           name, c.M, c.N, c.K, c.a_mn, c.b_mn, c.terms, c.n_off, max_err, sqrt(sum_sq / (c.M * c.N)), max_err / (max_ref + 1e-30),
           cyc, n_mma, (double)cyc / n_mma);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
    return max_err;
}

int main() {
    // {M, N, K, a_mn, b_mn, terms, reps, n_off}
    run16("bf16 L2 fwd K/K", {128, 64, 64, 0, 0, 1, 1, 0}, true, 2, true);
    run16("bf16 L1 fwd K/K K=16", {128, 64, 16, 0, 0, 1, 1, 0}, true, 1, true);
    run16("bf16 dH1 A K, B MN", {128, 64, 64, 0, 1, 1, 1, 0}, true, 3, true);
    run16("bf16 dW2 MN/MN M=64", {64, 64, 128, 1, 1, 1, 1, 0}, true, 4, true);
    run16("bf16 dW1 MN/MN M=64 N=16", {64, 16, 128, 1, 1, 1, 1, 0}, true, 5, true);
    run16("bf16 dW1 MN/MN M=64 N=16 off16", {64, 16, 128, 1, 1, 1, 1, 16}, true, 5, true);
    run16("bf16 dW1 MN/MN M=64 N=32", {64, 32, 128, 1, 1, 1, 1, 0}, true, 5, true);
    run16("bf16 dW MN/MN M=128 N=64", {128, 64, 128, 1, 1, 1, 1, 0}, true, 6, true);
    // precision on random fp32 data
    run16("random 1 term", {128, 64, 64, 0, 0, 1, 1, 0}, false, 9);
    run16("random 3 terms", {128, 64, 64, 0, 0, 3, 1, 0}, false, 9);
    run16("random 5 terms", {128, 64, 64, 0, 0, 5, 1, 0}, false, 9);
    run16("random 6 terms", {128, 64, 64, 0, 0, 6, 1, 0}, false, 9);
    run16("random 6 terms dW2 K=128", {64, 64, 128, 1, 1, 6, 1, 0}, false, 10);
    // timing with precomputed descriptors
    run16("time L2 x64", {128, 64, 64, 0, 0, 6, 64, 0}, true, 11);
    run16("time L2 N=128 x64", {128, 128, 64, 0, 0, 6, 64, 0}, true, 11);
    run16("time dW2 M=64 x64", {64, 64, 128, 1, 1, 6, 64, 0}, true, 12);
    run16("time dW1 M=64 N=16 x64", {64, 16, 128, 1, 1, 6, 64, 0}, true, 12);
    run16("time dH1 (B MN) x64", {128, 64, 64, 0, 1, 6, 64, 0}, true, 12);
    // tf32 reference points
    run("tf32 L2 fwd SW128 K/K", {128, 64, 64, 0, 0, 0, 0, 1, 0, 0, 0}, true, 2);
    run("tf32 3x random", {128, 64, 64, 0, 0, 0, 1, 1, 0, 0, 0}, false, 9);
    return 0;
}
