// pmctf_umma.cuh -- thin inline-PTX layer over the sm_100a 5th-generation tensor core (tcgen05):
// shared-memory matrix descriptors, the kind::i8 instruction descriptor, MMA issue, TMEM
// allocation / loads, mbarrier completion.  Only what the lifting kernels need.
//
// Operand layout used throughout (K-major, no swizzle; "INTERLEAVE" canonical layout): a matrix
// row is 16 contiguous bytes (one 16-element int8 K-chunk); row r of an operand lives at
//      start + (r % 8) * 16 + (r / 8) * SBO + chunk * LBO          (bytes)
// and one kind::i8 MMA consumes K = 32 = two chunks.  With SBO = 128 the rows of a chunk are a plain
// linear array of 16-byte records, which is what makes the implicit-GEMM convolution possible:
// a pixel's 16 channels are one record, and a filter tap is just a different start address.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmctf {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (sm_100: version field = 1, SWIZZLE_NONE, base offset 0)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// 32-bit instruction descriptor: D = s32, B = signed int8, A = signed (default) or unsigned int8, both K-major, M = 128, N = n
__host__ __device__ __forceinline__ constexpr uint32_t idesc_s8(uint32_t n, bool a_unsigned = false)
{
    return (2u << 4) | ((a_unsigned ? 0u : 1u) << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_s8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

// one lane of a converged warp (the values feeding the MMA stay in uniform registers this way)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// all previously issued MMAs of this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void commit(uint32_t mbar_saddr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMEM allocation: executed by one full warp; the base address is written to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread i of the warp reads lane
// 32*(warp%4)+i; the lane base must be encoded in bits 31:16 of the address)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// zero 16 consecutive 32-bit columns of this warp's 32 lanes (thread i owns lane 32*(warp%4)+i, as for the loads)
__device__ __forceinline__ void tmem_zero16(uint32_t taddr)
{
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t mbar_saddr, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// TMA bulk copy (1-D, no tensor map): global -> shared, completion counted in bytes on an mbarrier.  dst / src 16-byte aligned,
// bytes a multiple of 16.  Issued by one thread after mbar_expect_tx() announced the total.
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar_saddr, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void *src_global, uint32_t bytes, uint32_t mbar_saddr)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr),
                 "l"(__cvta_generic_to_global(src_global)), "r"(bytes), "r"(mbar_saddr)
                 : "memory");
}

// bounded wait (never hangs the GPU: returns false when the phase did not complete in time)
__device__ __forceinline__ bool mbar_wait(uint32_t mbar_saddr, uint32_t parity, int max_tries = 1 << 22)
{
    for (int i = 0; i < max_tries; ++i) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(mbar_saddr), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

} // namespace umma
} // namespace pmctf
