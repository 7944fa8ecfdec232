// pmctf_umma_test.cu -- executes a host-provided list of tcgen05 kind::i8 MMAs on operands staged in shared
// memory and returns the raw TMEM accumulators.  It is the unit test of the descriptor / layout conventions
// in pmctf_umma.cuh (tests/test_gpu_umma.py emulates the same op list in numpy) and a small timing probe.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pmctf_b200.h"
#include "pmctf_umma.cuh"

namespace pmctf {

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const int8_t *__restrict__ A, int a_bytes, const int8_t *__restrict__ B,
                                                               int b_bytes, const pmctf_umma_op_t *__restrict__ ops, int n_ops,
                                                               int n_blocks, int block_stride_bytes, int out_cols,
                                                               int *__restrict__ out, int repeat, long long *__restrict__ cycles,
                                                               int *__restrict__ err)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint64_t s_ad[128], s_bd[128];
    __shared__ uint32_t s_dcol[128], s_idesc[128], s_acc[128];
    uint8_t *sA = smem_raw;
    uint8_t *sB = smem_raw + ((a_bytes + 127) / 128) * 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < a_bytes / 16; i += blockDim.x)
        reinterpret_cast<int4 *>(sA)[i] = __ldg(reinterpret_cast<const int4 *>(A) + i);
    for (int i = tid; i < b_bytes / 16; i += blockDim.x)
        reinterpret_cast<int4 *>(sB)[i] = __ldg(reinterpret_cast<const int4 *>(B) + i);
    const uint32_t bar = umma::smem_u32(&mbar);
    if (tid == 0) {
        umma::mbar_init(bar, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, 128);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t a0 = umma::smem_u32(sA), b0 = umma::smem_u32(sB);
    // descriptors are built once by the whole CTA; the issuing thread only adds the block offset
    if (tid < n_ops) {
        const pmctf_umma_op_t op = ops[tid];
        s_ad[tid] = umma::smem_desc(a0 + op.a_off, op.a_lbo, op.a_sbo);
        s_bd[tid] = umma::smem_desc(b0 + op.b_off, op.b_lbo, op.b_sbo);
        s_dcol[tid] = tbase + op.d_col;
        s_idesc[tid] = umma::idesc_s8(op.n, op.a_unsigned != 0);
        s_acc[tid] = op.accumulate;
    }
    __syncthreads();

    uint32_t phase = 0;
    bool ok = true;
    for (int blk = 0; blk < n_blocks && ok; ++blk) {
        // accumulators start at zero (op lists may initialise only part of the columns with their first MMA)
        for (int c0 = 0; c0 < out_cols; c0 += 16) umma::tmem_zero16(tbase + ((uint32_t)(warp * 32) << 16) + c0);
        umma::tmem_st_wait();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        long long t0 = 0;
        if (tid == 0) {
            t0 = clock64();
            const uint64_t blk_off = (uint64_t)((blk * block_stride_bytes) >> 4);
            for (int r = 0; r < repeat; ++r) {
#pragma unroll 6
                for (int i = 0; i < n_ops; ++i)
                    umma::mma_s8(s_dcol[i], s_ad[i] + blk_off, s_bd[i], s_idesc[i], (r > 0) ? 1u : s_acc[i]);
            }
            umma::commit(bar);
        }
        ok = umma::mbar_wait(bar, phase);
        phase ^= 1;
        if (tid == 0 && cycles) cycles[blk] = clock64() - t0;
        if (!ok) {
            if (tid == 0) atomicExch(err, 1 + blk);
            break;
        }
        umma::fence_after_sync();
        for (int c0 = 0; c0 < out_cols; c0 += 16) {
            uint32_t v[16];
            umma::tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
            umma::tmem_ld_wait();
            int *o = out + ((long long)blk * 128 + warp * 32 + lane) * out_cols + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = (int)v[j];
        }
        umma::fence_before_sync();
        __syncthreads(); // every warp has drained the accumulators before the next block overwrites them
        umma::fence_after_sync();
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 128);
}

} // namespace pmctf

extern "C" int pmctf_umma_selftest(const signed char *A, int a_bytes, const signed char *B, int b_bytes, const pmctf_umma_op_t *ops,
                                   int n_ops, int n_blocks, int block_stride_bytes, int out_cols, int *out, int repeat,
                                   long long *cycles, int *err, void *stream)
{
    if (!A || !B || !ops || !out || !err || n_ops <= 0 || n_ops > 128 || n_blocks <= 0 || repeat <= 0) return PMCTF_EINVAL;
    if ((a_bytes & 15) || (b_bytes & 15) || out_cols <= 0 || out_cols > 128 || (out_cols & 15)) return PMCTF_ESHAPE;
    const int smem = ((a_bytes + 127) / 128) * 128 + b_bytes;
    if (smem > 200 * 1024) return PMCTF_ESHAPE;
    cudaError_t e = cudaFuncSetAttribute(pmctf::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    pmctf::umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(reinterpret_cast<const int8_t *>(A), a_bytes,
                                                                        reinterpret_cast<const int8_t *>(B), b_bytes, ops, n_ops, n_blocks,
                                                                        block_stride_bytes, out_cols, out, repeat, cycles, err);
    return (int)cudaGetLastError();
}
