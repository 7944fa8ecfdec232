// pmctf_pp.cu -- PostProcess (pMCTF/layers/postprocessing.py:20-44): the de-quantisation filter pWave++ applies to every
// reconstructed plane right behind the synthesis transform (pWave.py:299-300,347,455,526).  SURVEY.md section 8f row 2.
//
//      t = conv1(x)                       1 -> 64, 3x3
//      for 6 ResBlocks: t = conv_b(lrelu_0.2(conv_a(t))) + t          64 -> 64, 3x3 each
//      y = x + conv3(conv2(t) + conv1(x))                            64 -> 64, then 64 -> 1
//
// 13 dense 64 -> 64 convolutions = 958 kFLOP per pixel.  They run as implicit GEMMs on the 5th-generation tensor cores:
// tcgen05.mma kind::f16 with bf16 operands (activations and weights rounded to bf16 once, RN) and fp32 accumulators in TMEM;
// the residual stream, the biases and the skip connections stay fp32.  Same operand trick as the lifting kernel
// (pmctf_lift_tc.cu): the input tile lives in shared memory as a linear pixel array (pitch 32) of 16-byte records, one plane per
// group of 8 channels, K-major / un-swizzled, so a filter tap is only a different descriptor start address; one MMA
// (M = 128 pixels, N = 64, K = 16) consumes two planes (leading byte offset = plane size).  36 MMAs per 128-pixel block.
//
// Feature maps live in HBM in CHUNK-PLANAR layouts chosen so that one thread per pixel is fully coalesced on both sides of the
// tensor core: bf16 operands as [N][C/8][H][W][8] (a pixel's 8 channels = one 16-byte operand record; a warp's 32 pixels of a
// tile row = 512 contiguous bytes), fp32 residual stream as [N][C/4][H][W][4].
//
// CTA (1 per SM, persistent over tiles of 16 x 30 outputs = 4 blocks): warps 0-3 epilogue (TMEM lane quarters), warps 4-7 load
// the next input tile (16-byte loads of operand records, zero padding), warp 8 issues the MMAs.  Two input buffers
// and two sets of four accumulators (all 512 TMEM columns), so the loads of tile t+1 and the epilogue of tile t-1 overlap the MMAs
// of tile t.  All 9 x 64 x 64 weights stay resident in shared memory (72 KB, one TMA bulk copy per CTA).
//
// Numerics: not bit-exact by construction (the tensor core's fp32 summation order is unspecified).  The CPU oracle restates the
// reference in plain fp32 (oracle/pmctf_oracle.c: orc_postprocess); the tests bound the difference by the north-star
// tolerance for frames (1e-3 on the [0,1] pixel scale) and check single layers against a bf16-operand / fp32-accumulate
// emulation to 1e-4.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "pmctf_b200.h"
#include "pmctf_umma.cuh"

namespace pmctf {
namespace pp {

constexpr int C = 64;                                // input channels of the tensor-core layers
constexpr int TH = 16, TW = 30, P = 32;              // output tile; pixel pitch of the staged input (TW + 2)
constexpr int IN_R = TH + 2;
constexpr int NBLK = (TH * P) / 128;                 // 4 blocks of 128 output pixel records (2 junk columns per row)
constexpr int NPIX = 585;                            // >= IN_R * P + 2 (over-read of the last block), == 1 (mod 8): planes start in
                                                     // different bank groups, the staging stores are conflict-free
constexpr int PLANE = NPIX * 16;                     // bytes of one 8-channel plane
constexpr int INBUF = (C / 8) * PLANE;               // 74 880 B
constexpr int WMAX = 36 * 64 * 32;                   // 73 728 B: [tap][k-step][chunk][co][8 ci] bf16
constexpr int SM_W = 0;
constexpr int SM_IN = SM_W + WMAX;
constexpr int SM_BAR = SM_IN + 2 * INBUF;
constexpr int SM_BIAS = SM_BAR + 128;                // 64 floats
constexpr int SMEM_BYTES = SM_BIAS + 256;
static_assert(TH * P == NBLK * 128 && IN_R * P + 2 + 2 * P <= NPIX + 2 * P, "tile geometry");
static_assert((NBLK - 1) * 128 + 127 + 2 * P + 2 < NPIX, "operand reads stay inside the plane");
static_assert(SM_IN % 128 == 0 && INBUF % 16 == 0 && SM_BAR % 8 == 0 && SMEM_BYTES <= 227 * 1024, "shared memory");

constexpr int EPI_WARPS = 4, LOAD_WARPS = 4;
constexpr int MMA_WARP = EPI_WARPS + LOAD_WARPS;
constexpr int NT = 32 * (MMA_WARP + 1);

// 32-bit instruction descriptor of kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}

struct ConvD {
    const __nv_bfloat16 *in;      // bf16 [N][8][H][W][8], 64 channels
    const uint8_t *wimg;          // packed operand image (pp_pack_kernel)
    const float *bias;            // [co]
    const float *res;             // fp32 [N][16][H][W][4], added after the bias (skip connection), or null
    float *out_f32;               // fp32 [N][16][H][W][4], or null
    __nv_bfloat16 *out_bf16;      // bf16 [N][8][H][W][8], or null
    const float *x_plane;         // last layer (co == 1): y = (x * in_mul + conv) * out_mul on [N,1,H,W] planes
    float *y_plane;
    float slope, in_mul, out_mul; // LeakyReLU slope (1 = identity)
    int n, h, w, co;              // co: 64, or 1 (padded to N = 16)
};

template <int CO_PAD>
__global__ void __launch_bounds__(NT, 1) pp_conv64_kernel(const __grid_constant__ ConvD a, int *__restrict__ err)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);   // in_full[2], in_empty[2], acc_full[2], acc_empty[2], weights
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_BAR + 96);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int H = a.h, W = a.w;
    const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
    const int n_tiles = tiles_x * tiles_y * a.n;
    constexpr int WBYTES = 36 * CO_PAD * 32;
    constexpr int ACC_COLS = NBLK * 64;              // one accumulator set (blocks 64 columns apart also when CO_PAD < 64)

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(umma::smem_u32(bars + i), 32 * LOAD_WARPS);      // in_full
            umma::mbar_init(umma::smem_u32(bars + 2 + i), 1);                // in_empty  (tcgen05.commit)
            umma::mbar_init(umma::smem_u32(bars + 4 + i), 1);                // acc_full  (tcgen05.commit)
            umma::mbar_init(umma::smem_u32(bars + 6 + i), 32 * EPI_WARPS);   // acc_empty
        }
        umma::mbar_init(umma::smem_u32(bars + 8), 1);
        umma::fence_mbar_init();
        umma::mbar_expect_tx(umma::smem_u32(bars + 8), WBYTES);
        umma::bulk_g2s(umma::smem_u32(smem + SM_W), a.wimg, WBYTES, umma::smem_u32(bars + 8));
    }
    if (warp == MMA_WARP) umma::tmem_alloc(tmem_slot, 512);
    if (tid < 64) reinterpret_cast<float *>(smem + SM_BIAS)[tid] = tid < a.co ? a.bias[tid] : 0.0f;
    // the over-read records behind the tile (they only feed junk rows of the last block) hold zeros
    for (int i = tid; i < 2 * (C / 8) * (NPIX - IN_R * P); i += NT) {
        const int buf = i / ((C / 8) * (NPIX - IN_R * P)), j = i - buf * ((C / 8) * (NPIX - IN_R * P));
        const int pl = j / (NPIX - IN_R * P), px = IN_R * P + j - pl * (NPIX - IN_R * P);
        *reinterpret_cast<uint4 *>(smem + SM_IN + buf * INBUF + pl * PLANE + px * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t in_full = umma::smem_u32(bars), in_empty = umma::smem_u32(bars + 2), acc_full = umma::smem_u32(bars + 4),
                   acc_empty = umma::smem_u32(bars + 6);
    bool ok = true;

    if (warp >= EPI_WARPS && warp < MMA_WARP) {
        // ---- loaders: NHWC bf16 -> planes of 16-byte pixel records, zero padding outside the image ----------------------
        const int lt = tid - 32 * EPI_WARPS;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            if (it >= 2) {   // the MMAs that read this buffer two tiles ago have completed
                ok = umma::mbar_wait(in_empty + 8 * buf, (uint32_t)((it >> 1) - 1) & 1u);
                if (!ok) break;
            }
            const int n = tile / (tiles_x * tiles_y), trem = tile - n * (tiles_x * tiles_y);
            const int ty = trem / tiles_x, y0 = ty * TH, x0 = (trem - ty * tiles_x) * TW;
            uint8_t *dst = smem + SM_IN + buf * INBUF;
            const long long plane_px = (long long)H * W;
            const uint4 *src = reinterpret_cast<const uint4 *>(a.in) + (long long)n * (C / 8) * plane_px;   // one uint4 = one operand record
            constexpr int ITEMS = IN_R * P * (C / 8), PER = ITEMS / (32 * LOAD_WARPS);
            static_assert(ITEMS % (32 * LOAD_WARPS) == 0 && (IN_R * P) % 32 == 0, "a loader warp = 32 consecutive pixels of one chunk");
            uint4 v[PER];
#pragma unroll
            for (int k = 0; k < PER; ++k) {   // all loads of a thread in flight together; a warp = 32 consecutive pixels of one row and chunk
                const int i = lt + k * 32 * LOAD_WARPS, ch = i / (IN_R * P), px = i - ch * (IN_R * P);
                const int r = px >> 5, c = px & 31, gy = y0 - 1 + r, gx = x0 - 1 + c;
                v[k] = make_uint4(0u, 0u, 0u, 0u);
                if (gy >= 0 && gy < H && gx >= 0 && gx < W) v[k] = __ldg(src + ch * plane_px + (long long)gy * W + gx);
            }
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int i = lt + k * 32 * LOAD_WARPS, ch = i / (IN_R * P), px = i - ch * (IN_R * P);
                *reinterpret_cast<uint4 *>(dst + ch * PLANE + px * 16) = v[k];
            }
            umma::fence_proxy_async();
            mbar_arrive(in_full + 8 * buf);
        }
    } else if (warp == MMA_WARP) {
        // ---- MMA issue -----------------------------------------------------------------------------------------------
        ok = __shfl_sync(0xffffffffu, (int)umma::mbar_wait(umma::smem_u32(bars + 8), 0u), 0) != 0;   // weights have landed
        int it = 0;
        for (int tile = blockIdx.x; ok && tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            ok = __shfl_sync(0xffffffffu, (int)umma::mbar_wait(in_full + 8 * buf, (uint32_t)(it >> 1) & 1u), 0) != 0;
            if (!ok) break;
            if (it >= 2) {   // the epilogue has drained this accumulator set
                ok = __shfl_sync(0xffffffffu, (int)umma::mbar_wait(acc_empty + 8 * buf, (uint32_t)((it >> 1) - 1) & 1u), 0) != 0;
                if (!ok) break;
            }
            umma::fence_after_sync();
            if (umma::elect_one()) {
                // descriptors: the start address sits in the low 14 bits (16-byte units) and never carries out of them (shared memory
                // < 256 KB), so a tap / k-step / block is ONE 64-bit add of a constant to a base descriptor
                const uint64_t a0 = umma::smem_desc(umma::smem_u32(smem + SM_IN + buf * INBUF), PLANE, 128);
                const uint64_t b0 = umma::smem_desc(umma::smem_u32(smem + SM_W), CO_PAD * 16, 128);
#pragma unroll 1
                for (int blk = 0; blk < NBLK; ++blk) {
                    const uint32_t d = tbase + buf * ACC_COLS + blk * 64;
                    const uint64_t ab = a0 + (uint64_t)(blk * 128);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int off = (tap / 3) * P + (tap % 3);
#pragma unroll
                        for (int ks = 0; ks < C / 16; ++ks)
                            mma_bf16(d, ab + (uint64_t)((2 * ks) * (PLANE / 16) + off), b0 + (uint64_t)((tap * (C / 16) + ks) * (CO_PAD * 2)),
                                     idesc_bf16(CO_PAD), (tap | ks) ? 1u : 0u);
                    }
                }
                umma::commit(in_empty + 8 * buf);    // input buffer free once these MMAs have completed ...
                umma::commit(acc_full + 8 * buf);    // ... and the accumulators are ready
            }
            __syncwarp();
        }
    } else {
        // ---- epilogue: TMEM -> bias / skip / LeakyReLU -> global ------------------------------------------------------------
        const int quarter = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            ok = __all_sync(0xffffffffu, (int)umma::mbar_wait(acc_full + 8 * buf, (uint32_t)(it >> 1) & 1u)) != 0;
            if (!ok) break;
            umma::fence_after_sync();
            const int n = tile / (tiles_x * tiles_y), trem = tile - n * (tiles_x * tiles_y);
            const int ty = trem / tiles_x, y0 = ty * TH, x0 = (trem - ty * tiles_x) * TW;
#pragma unroll 1
            for (int blk = 0; blk < NBLK; ++blk) {
                const int m = blk * 128 + quarter * 32 + lane;
                const int r = m >> 5, c = m & 31, gy = y0 + r, gx = x0 + c;
                const bool valid = c < TW && gy < H && gx < W;
                const long long plane_px = (long long)H * W, pix = (long long)gy * W + gx;
                const float *sbias = reinterpret_cast<const float *>(smem + SM_BIAS);
                const uint32_t taddr = tbase + ((uint32_t)(quarter * 32) << 16) + buf * ACC_COLS + blk * 64;
                if (CO_PAD == 64) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {   // 16 channels per round
                        uint32_t o[16];
                        umma::tmem_ld16(taddr + 16 * q, o);
                        float4 rr[4];
                        const bool use_res = a.res != nullptr && valid;
                        if (use_res) {   // lanes = consecutive pixels: 512 contiguous bytes per load instruction
                            const float4 *rp = reinterpret_cast<const float4 *>(a.res) + ((long long)n * 16 + 4 * q) * plane_px + pix;
#pragma unroll
                            for (int j = 0; j < 4; ++j) rr[j] = __ldg(rp + j * plane_px);
                        }
                        umma::tmem_ld_wait();
                        if (valid) {
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 bb = *reinterpret_cast<const float4 *>(sbias + 16 * q + 4 * j);
                                v[4 * j] = __uint_as_float(o[4 * j]) + bb.x; v[4 * j + 1] = __uint_as_float(o[4 * j + 1]) + bb.y;
                                v[4 * j + 2] = __uint_as_float(o[4 * j + 2]) + bb.z; v[4 * j + 3] = __uint_as_float(o[4 * j + 3]) + bb.w;
                            }
                            if (use_res) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    v[4 * j] += rr[j].x; v[4 * j + 1] += rr[j].y; v[4 * j + 2] += rr[j].z; v[4 * j + 3] += rr[j].w;
                                }
                            }
                            if (a.slope != 1.0f) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) v[j] = v[j] >= 0.0f ? v[j] : v[j] * a.slope;
                            }
                            if (a.out_f32) {
                                float4 *op = reinterpret_cast<float4 *>(a.out_f32) + ((long long)n * 16 + 4 * q) * plane_px + pix;
#pragma unroll
                                for (int j = 0; j < 4; ++j) op[j * plane_px] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                            }
                            if (a.out_bf16) {
                                uint32_t pk[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                                    pk[j] = *reinterpret_cast<const uint32_t *>(&b2);
                                }
                                uint4 *op = reinterpret_cast<uint4 *>(a.out_bf16) + ((long long)n * 8 + 2 * q) * plane_px + pix;
                                op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                op[plane_px] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                            }
                        }
                    }
                } else {   // last layer: one real output channel (column 0), y = (x * in_mul + conv + bias) * out_mul
                    uint32_t o[16];
                    umma::tmem_ld16(taddr, o);
                    umma::tmem_ld_wait();
                    if (valid) {
                        const float t = __uint_as_float(o[0]) + sbias[0];
                        a.y_plane[(long long)n * plane_px + pix] = (__ldg(a.x_plane + (long long)n * plane_px + pix) * a.in_mul + t) * a.out_mul;
                    }
                }
            }
            umma::fence_before_sync();
            mbar_arrive(acc_empty + 8 * buf);
        }
    }
    if (__syncthreads_or(!ok)) {
        if (tid == 0 && err) {
            *reinterpret_cast<volatile int *>(err) = 1;
            __threadfence_system();
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) umma::tmem_dealloc(tbase, 512);
}

// conv1 of PostProcess: 1 -> 64 channels on the CUDA cores (9 MACs per output), x * in_mul as the input; writes the fp32 NHWC
// feature map (the skip operand of postprocessing.py:40) and its bf16 copy (the operand of the first ResBlock)
__global__ void __launch_bounds__(256) pp_conv_in_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                                                         float in_mul, float *__restrict__ out_f32, __nv_bfloat16 *__restrict__ out_bf16,
                                                         int N, int H, int W)
{
    __shared__ float sw[64 * 9], sb[64];
    for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 64) sb[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    const long long plane_px = (long long)H * W, all_px = (long long)N * plane_px, total = all_px * 4;   // one thread = one pixel x 16 channels
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i / all_px);                 // consecutive threads = consecutive pixels (coalesced both ways)
        const long long apix = i - q * all_px;
        const long long n = apix / plane_px, pix = apix - n * plane_px;
        const int gx = (int)(pix % W), gy = (int)(pix / W);
        const float *xp = x + n * plane_px;
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int yy = gy + k / 3 - 1, xx = gx + k % 3 - 1;
            v[k] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(xp + (long long)yy * W + xx) * in_mul : 0.0f;
        }
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float acc = sb[16 * q + j];
#pragma unroll
            for (int k = 0; k < 9; ++k) acc = fmaf(sw[(16 * q + j) * 9 + k], v[k], acc);
            o[j] = acc;
        }
        float4 *of = reinterpret_cast<float4 *>(out_f32) + (n * 16 + 4 * q) * plane_px + pix;
#pragma unroll
        for (int j = 0; j < 4; ++j) of[j * plane_px] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
            pk[j] = *reinterpret_cast<const uint32_t *>(&b2);
        }
        uint4 *op = reinterpret_cast<uint4 *>(out_bf16) + (n * 8 + 2 * q) * plane_px + pix;
        op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        op[plane_px] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
}

// OIHW fp32 [co][64][3][3] -> bf16 operand image [tap][k-step][chunk][co_pad rows][8 ci] (rows beyond co are zero)
__global__ void pp_pack_kernel(const float *__restrict__ w, int co, int co_pad, __nv_bfloat16 *__restrict__ img)
{
    const int total = 36 * co_pad * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = i & 7, row = (i >> 3) % co_pad, chunk = (i / (8 * co_pad)) & 1, ks = (i / (16 * co_pad)) & 3, tap = i / (64 * co_pad);
        const int ci = ks * 16 + chunk * 8 + e;
        const float v = row < co ? w[((long long)row * 64 + ci) * 9 + tap] : 0.0f;
        img[i] = __float2bfloat16_rn(v);
    }
}

// fp32 -> bf16 element-wise (RN); layout conversions are the caller's business
__global__ void __launch_bounds__(256) pp_to_bf16_kernel(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out, long long n4)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(in) + i);
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        reinterpret_cast<uint2 *>(out)[i] = make_uint2(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b));
    }
}

} // namespace pp

int tc_watchdog(volatile int **host, int **dev);   // pmctf_kernels.cu: the per-device watchdog word (mapped pinned memory)
// pmctf_ctx.cu: the CTA-pair kernel (cta_group::2, TMA tensor loads) instantiated for 64 channels -- same feature-map layouts
int pair_conv64(const void *in_bf16, const void *packed_w, const float *bias, const float *res, float slope, float *out_f32, void *out_bf16, int N,
                int H, int W, void *stream);
int pair_pack64(const float *w, void *packed, void *stream);
static bool use_pair()
{   // PMCTF_PP_SINGLE_CTA=1 keeps the round-2a single-CTA kernel (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PMCTF_PP_SINGLE_CTA");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}
void count_launch();                               // pmctf_kernels.cu

static int launch_conv64(const pp::ConvD &d, cudaStream_t st)
{
    static int sms_of[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
    if (dev < 0 || dev >= 64) return PMCTF_EINVAL;
    if (!sms_of[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pp::pp_conv64_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, pp::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pp::pp_conv64_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, pp::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return (int)cudaGetLastError();
        sms_of[dev] = sms;
    }
    volatile int *herr = nullptr;
    int *derr = nullptr;
    int e = tc_watchdog(&herr, &derr);
    if (e) return e;
    if (herr[0] != 0) return PMCTF_ETIMEOUT;
    const long long tiles = (long long)((d.w + pp::TW - 1) / pp::TW) * ((d.h + pp::TH - 1) / pp::TH) * d.n;
    if (tiles <= 0 || tiles > 0x7fffffffLL) return PMCTF_ESHAPE;
    const unsigned grid = (unsigned)(tiles < sms_of[dev] ? tiles : sms_of[dev]);
    if (d.co == 64)
        pp::pp_conv64_kernel<64><<<grid, pp::NT, pp::SMEM_BYTES, st>>>(d, derr);
    else
        pp::pp_conv64_kernel<16><<<grid, pp::NT, pp::SMEM_BYTES, st>>>(d, derr);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace pmctf

using namespace pmctf;

extern "C" {

long long pmctf_pp_packed_bytes(int co) { return co == 64 ? 36LL * 64 * 32 : (co >= 1 && co <= 16 ? 36LL * 16 * 32 : 0); }

int pmctf_pp_pack_conv(const float *w, int co, void *packed, void *stream)
{
    if (!w || !packed || !(co == 64 || (co >= 1 && co <= 16))) return PMCTF_EINVAL;
    if (((uintptr_t)packed & 15) != 0) return PMCTF_EINVAL;
    if (co == 64 && use_pair()) return pair_pack64(w, packed, stream);
    const int co_pad = co == 64 ? 64 : 16;
    pp::pp_pack_kernel<<<(36 * co_pad * 16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, co, co_pad, (__nv_bfloat16 *)packed);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_pp_conv_in(const float *x, const float *w, const float *b, float in_mul, float *out_f32, void *out_bf16, int N, int H, int W,
                     void *stream)
{
    if (!x || !w || !b || !out_f32 || !out_bf16 || N <= 0 || H <= 0 || W <= 0) return PMCTF_EINVAL;
    long long blocks = ((long long)N * H * W * 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pp::pp_conv_in_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, w, b, in_mul, out_f32, (__nv_bfloat16 *)out_bf16, N, H, W);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_pp_to_bf16(const float *in, void *out, long long n, void *stream)
{
    if (n == 0) return 0;
    if (!in || !out || n < 0 || (n & 3) || (((uintptr_t)in | (uintptr_t)out) & 7)) return PMCTF_EINVAL;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pp::pp_to_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16 *)out, n / 4);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_pp_conv64(const void *in_bf16, const void *packed_w, const float *bias, int co, const float *residual, float lrelu_slope,
                    float *out_f32, void *out_bf16, const float *x_plane, float in_mul, float out_mul, float *y_plane, int N, int H,
                    int W, void *stream)
{
    if (!in_bf16 || !packed_w || !bias || N <= 0 || H <= 0 || W <= 0) return PMCTF_EINVAL;
    if (co == 64) {
        if (!out_f32 && !out_bf16) return PMCTF_EINVAL;
    } else if (co == 1) {
        if (!x_plane || !y_plane) return PMCTF_EINVAL;
    } else {
        return PMCTF_EINVAL;
    }
    if ((((uintptr_t)in_bf16 | (uintptr_t)packed_w | (uintptr_t)out_f32 | (uintptr_t)out_bf16 | (uintptr_t)residual) & 15) != 0) return PMCTF_EINVAL;
    if (co == 64 && use_pair()) return pair_conv64(in_bf16, packed_w, bias, residual, lrelu_slope, out_f32, out_bf16, N, H, W, stream);
    pp::ConvD d;
    d.in = (const __nv_bfloat16 *)in_bf16; d.wimg = (const uint8_t *)packed_w; d.bias = bias; d.res = residual;
    d.out_f32 = out_f32; d.out_bf16 = (__nv_bfloat16 *)out_bf16; d.x_plane = x_plane; d.y_plane = y_plane;
    d.slope = lrelu_slope; d.in_mul = in_mul; d.out_mul = out_mul; d.n = N; d.h = H; d.w = W; d.co = co;
    return launch_conv64(d, (cudaStream_t)stream);
}

long long pmctf_postprocess_workspace(int H, int W) { return (long long)H * W * 64 * (3 * 4 + 2 * 2); }

int pmctf_postprocess(const float *x, const pmctf_postprocess_t *p, float in_mul, float out_mul, float *y, int N, int H, int W,
                      void *workspace, long long workspace_bytes, void *stream)
{
    if (!x || !p || !y || !workspace || N <= 0 || H <= 0 || W <= 0) return PMCTF_EINVAL;
    if (workspace_bytes < pmctf_postprocess_workspace(H, W) || ((uintptr_t)workspace & 255)) return PMCTF_EWORKSPACE;
    if (!p->conv1_w || !p->conv1_b || !p->conv2_w || !p->conv2_b || !p->conv3_w || !p->conv3_b) return PMCTF_EINVAL;
    for (int i = 0; i < 12; ++i)
        if (!p->res_w[i] || !p->res_b[i]) return PMCTF_EINVAL;
    const long long px = (long long)H * W;
    uint8_t *ws = (uint8_t *)workspace;
    float *c1 = (float *)ws, *ra = c1 + px * 64, *rb = ra + px * 64;     // fp32: conv1 features (skip), residual stream ping / pong
    void *ba = (void *)(rb + px * 64), *bb = (void *)((uint8_t *)ba + px * 128);   // bf16 operand buffers
    for (int n = 0; n < N; ++n) {   // plane by plane: the workspace holds one plane's feature maps
        const float *xn = x + n * px;
        float *yn = y + n * px;
        int e = pmctf_pp_conv_in(xn, p->conv1_w, p->conv1_b, in_mul, c1, ba, 1, H, W, stream);   // t = conv1(x)
        if (e) return e;
        const float *cur_f = c1;
        void *cur_b = ba, *tmp_b = bb;
        for (int r = 0; r < 6; ++r) {   // ResBlock: t = conv_b(lrelu(conv_a(t))) + t
            e = pmctf_pp_conv64(cur_b, p->res_w[2 * r], p->res_b[2 * r], 64, nullptr, 0.2f, nullptr, tmp_b, nullptr, 1, 1, nullptr, 1, H, W, stream);
            if (e) return e;
            float *nxt_f = (cur_f == ra) ? rb : ra;
            e = pmctf_pp_conv64(tmp_b, p->res_w[2 * r + 1], p->res_b[2 * r + 1], 64, cur_f, 1.0f, nxt_f, cur_b, nullptr, 1, 1, nullptr, 1, H, W, stream);
            if (e) return e;
            cur_f = nxt_f;
        }
        // tmp = conv2(t) + conv1(x);  y = x + conv3(tmp)
        e = pmctf_pp_conv64(cur_b, p->conv2_w, p->conv2_b, 64, c1, 1.0f, nullptr, tmp_b, nullptr, 1, 1, nullptr, 1, H, W, stream);
        if (e) return e;
        e = pmctf_pp_conv64(tmp_b, p->conv3_w, p->conv3_b, 1, nullptr, 1.0f, nullptr, nullptr, xn, in_mul, out_mul, yn, 1, H, W, stream);
        if (e) return e;
    }
    return 0;
}

} // extern "C"
