// tc_primitives.cuh -- tcgen05 / tensor-memory / UMMA-descriptor helpers shared by the tensor-core policy kernels
// (actor_tc_kernel.cuh, actor_tc_wide_kernel.cuh, policy_step_tc_kernel.cuh).  Inline device functions only.
#pragma once
#include "msched_common.cuh"

namespace msched {

// ---- tcgen05 / tensor-memory primitives (PTX ISA 8.6+, sm_100a) ---------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor, no swizzle, K-major: core matrices adjacent in K are lbo bytes
// apart, 8-row groups adjacent in M/N are sbo bytes apart (both in 16-byte units), version 1
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32, A and B tf32, both K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// bounded mbarrier wait: a lost completion traps instead of hanging the device
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity)
{
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// stage an [n_real][k_real] torch weight matrix as hi / lo B operands ([Npad][Kpad], zero padded)
__device__ __forceinline__ void stage_weight(const float *__restrict__ w, int n_real, int k_real, int Npad, int Kpad,
                                             unsigned char *dstHi, unsigned char *dstLo, float scale = 1.f)
{
    const int panel = Npad * 16;  // bytes of one K chunk (4 floats) of all rows
#pragma unroll 4
    for (int i = threadIdx.x; i < Npad * Kpad; i += blockDim.x) {
        const int n = i / Kpad, k = i - n * Kpad;
        const float v = (n < n_real && k < k_real) ? w[n * k_real + k] * scale : 0.f;
        const float hi = tf32_hi(v);
        const int off = (k >> 2) * panel + n * 16 + (k & 3) * 4;
        *reinterpret_cast<float *>(dstHi + off) = hi;
        *reinterpret_cast<float *>(dstLo + off) = v - hi;
    }
}

}  // namespace msched
