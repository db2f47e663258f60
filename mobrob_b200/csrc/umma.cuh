// sm_100a tensor-core plumbing used by the PPO update kernel: tcgen05.mma (kind::f16 on fp16 operand
// pairs in the product path, ppo_tc.cuh; kind::tf32 for the layout probes) with
// shared-memory operand descriptors, tensor-memory (TMEM) allocation and loads, mbarrier
// completion, and the shared-memory operand layout both major modes can read.
//
// Operand layout ("panel"): a panel is R rows of 128 bytes (32 tf32 or 64 fp16 values); inside every
// 8-row group the 16-byte chunks of a row are XOR-swizzled with (row & 7) -- the hardware's
// SWIZZLE_128B pattern, so panels must start on 1024-byte boundaries.  The same bytes are a valid
//   * K-major operand   : row = M/N index, the 32 values of a row = 32 consecutive K
//   * MN-major operand  : row = K index,   the 32 values of a row = 32 consecutive M/N
// which is what lets one activation buffer feed the forward GEMM (samples x units) and the
// weight-gradient GEMM (units x samples) without a transpose.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mr {
namespace umma {

constexpr uint32_t PANEL_ROW_BYTES = 128;
constexpr uint32_t ATOM_BYTES = 1024;  // 8 rows x 128 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// byte offset of 32-bit element (row, col) inside a panel, col in [0, 32)
__device__ __forceinline__ uint32_t panel_off(uint32_t row, uint32_t col) {
    return row * PANEL_ROW_BYTES + ((((col >> 2) ^ row) & 7u) << 4) + ((col & 3u) << 2);
}
// byte offset of the 16-byte chunk (row, chunk) inside a panel, chunk in [0, 8)
__device__ __forceinline__ uint32_t panel_chunk_off(uint32_t row, uint32_t chunk) {
    return row * PANEL_ROW_BYTES + (((chunk ^ row) & 7u) << 4);
}

// ---- descriptors ---------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit): start address >> 4 in [0,14), leading-dimension byte
// offset >> 4 in [16,30), stride-dimension byte offset >> 4 in [32,46), descriptor version 1 in
// [46,48), swizzle mode in [61,64) (2 = 128-byte swizzle).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// K-major operand: rows = M/N, 8-row groups 1024 B apart; advance K by 8 values = +32 B
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t panel_saddr, uint32_t kstep_in_panel) {
    return smem_desc(panel_saddr + kstep_in_panel * 32u, 16u, ATOM_BYTES);
}
// MN-major operand: rows = K; 32-wide M/N groups live in consecutive panels `panel_stride` bytes
// apart; advance K by 8 rows = +1024 B
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t panel_saddr, uint32_t kstep, uint32_t panel_stride) {
    return smem_desc(panel_saddr + kstep * ATOM_BYTES, panel_stride, ATOM_BYTES);
}

// Instruction descriptor (32 bit) for kind::tf32 with fp32 accumulation.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM ------------------------------------------------------------------------------------------
// one full warp; writes the base address (lane 0, column c) to *slot in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the tensor core's (async proxy) operand reads
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 32 consecutive columns: thread i of the warp gets lane (base lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }

// ---- MMA issue / completion ----------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate ? 1u : 0u)
        : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate));
}
// arrive on the mbarrier when every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}

// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// fp32 -> (hi, lo) with hi = the value the tensor core sees (low 13 mantissa bits dropped) and
// lo = the exact remainder; x*y ~= hi*hi' + hi*lo' + lo*hi' to ~2^-21 relative ("3xTF32")
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

}  // namespace umma
}  // namespace mr
