// pmctf_pairconv.cu -- a k x k convolution (k = 1, 3, 7) between channel-chunked bf16 feature maps as a tcgen05 CTA-pair
// implicit GEMM with run-time channel counts, and on top of it SpyNet motion estimation:
// pMCTF/layers/video/video_net.py:74-121 (MEBasic: 7x7 convolutions 8 -> 32 -> 64 -> 32 -> 16 -> 2 with ReLU; ME_Spynet: six
// pyramid levels, each warping the second image by the upsampled flow of the level below, video_net.py:113-119), called once
// per coded frame pair at pMCTF_L.py:260,460.  640 kFLOP per luma pixel -- more than three times the whole lifting path.
//
// Same machine as csrc/pmctf_ctx.cu (read that header first): the two CTAs of a cluster issue one MMA with M = 256
// (tcgen05.mma.cta_group::2, bf16 operands, fp32 accumulators in TMEM); each CTA stages its own 4-row pixel tile by ONE TMA
// tensor load (box {8 ch, 32 px, 4 + k - 1 rows, C_in / 8 planes}, zero fill = padding) and holds half of the output channels'
// weights, resident for the launch (at most 49 taps x 4 k-steps x 16 rows x 32 B = 100 KB for the 64 -> 32 layer).  A filter tap is a
// descriptor start address (row pitch 32 records), so the 7 x 7 layers issue 49 x C_in / 16 MMAs per tile pair straight from the
// staged tile: no im2col, no per-tap copies.  Here everything the network-specific kernel has as compile-time constants
// (channel counts, taps, tile width 32 - (k - 1)) is a run-time field of the descriptor.
//
// Motion estimation runs in the ENCODER only (the decoder receives coded vectors), so bf16 operands cost no encoder / decoder
// drift; the tests bound the flow error against the fp32 oracle.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pmctf_b200.h"
#include "pmctf_umma.cuh"

namespace pmctf {
namespace pair {

constexpr int TH = 4, P = 32;                         // output rows per CTA tile; pixel pitch of the staged input
constexpr int N_IN = 2, MAX_DEPTH = 8, N_ACC = 4, ACC_STRIDE = 128;   // N_IN full-size tiles of room = up to MAX_DEPTH smaller ones
constexpr int MAX_W = 100352;                         // bytes of resident weights per CTA
constexpr int MAX_IN = 40960;                         // bytes of one staged tile (8 planes x 10 rows x 32 records x 16 B)
constexpr int SM_W = 0;
constexpr int SM_IN = SM_W + MAX_W;
constexpr int SM_BAR = SM_IN + N_IN * MAX_IN;
constexpr int SM_BIAS = SM_BAR + 256;                 // barriers: in_full[8], in_empty[8], acc_full[4], acc_empty[4], weights, peer weights
constexpr int SMEM_BYTES = SM_BIAS + 128 * 4;
static_assert(SM_IN % 128 == 0 && MAX_IN % 128 == 0 && SM_BAR % 8 == 0 && SMEM_BYTES <= 227 * 1024, "shared memory");
constexpr int EPI_SETS = 2, EPI_WARPS = 4 * EPI_SETS;
constexpr int PRODUCER_WARP = 0, MMA_WARP = 1, EPI_WARP0 = 2;
constexpr int NT = 32 * (EPI_WARP0 + EPI_WARPS);

__device__ __forceinline__ uint32_t idesc_bf16_m256(uint32_t n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t mbar_saddr, uint32_t parity, int max_tries = 1 << 22)
{
    for (int i = 0; i < max_tries; ++i) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(mbar_saddr), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void commit_pair(uint32_t mbar_saddr)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(mbar_saddr),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst_saddr, const CUtensorMap *map, uint32_t leader_mbar_cluster_addr, int c0,
                                                 int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            dst_saddr),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// The MMAs of one tile pair, issued by ONE thread.  That thread executes a dependent instruction every ~4 cycles, so the descriptor
// arithmetic per MMA must stay at a handful of instructions or the issue loop, not the tensor pipe, paces the layer (measured: 132
// cycles per MMA with run-time loops over taps and k-steps, against 16 cycles of math).  With the filter size and the k-step count as
// template parameters every A-descriptor offset is an immediate; the B descriptor advances by one run-time stride per MMA.
template <int KS, int KSTEPS>
__device__ __forceinline__ void issue_tile(uint32_t d, uint64_t a0, uint64_t bd, uint64_t b_step, uint32_t idesc)
{
    constexpr int PLANE16 = (TH + KS - 1) * P;   // plane size in 16-byte units
#pragma unroll
    for (int ky = 0; ky < KS; ++ky)
#pragma unroll
        for (int kx = 0; kx < KS; ++kx)
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                mma_bf16_pair(d, a0 + (uint64_t)(ky * P + kx + 2 * ks * PLANE16), bd, idesc, (ky | kx | ks) ? 1u : 0u);
                bd += b_step;
            }
}

struct ConvD {
    const uint8_t *wimg;          // [rank][tap][k-step][chunk][cout_pad / 2 rows][8 ci] bf16
    const float *bias;            // [cout]
    __nv_bfloat16 *out_bf16;      // [N][cout_pad / 8][H][W][8] (channels beyond cout are written as zeros), or null
    float *out_nchw;              // fp32 [N][cout][H][W], or null
    const float *add_nchw;        // fp32 [N][cout][H][W] added to the NCHW output (SpyNet: flow_up + conv5, video_net.py:115-119), or null
    float slope;                  // LeakyReLU slope applied before the stores (0 = ReLU, 1 = identity)
    int n, h, w;
    int ks, planes, cout, cout_pad;   // filter size (1, 3, 7), input planes of 8 channels (even), real / padded output channels
    int depth;                        // input ring depth: as many staged tiles as fit the room of N_IN full-size ones (2 .. 8)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1)
    pair_conv_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvD a, int *__restrict__ err)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_BAR + 248);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int H = a.h, W = a.w, pad = a.ks >> 1, TW = P - 2 * pad, in_r = TH + 2 * pad;
    const int ksteps = a.planes >> 1, taps = a.ks * a.ks, nhalf = a.cout_pad >> 1;
    const int plane_bytes = in_r * P * 16, inbuf = a.planes * plane_bytes, wslab = 2 * nhalf * 16, wbytes = taps * ksteps * wslab;
    const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
    const int n_tiles = tiles_x * tiles_y * a.n;
    const int n_pairs = (n_tiles + 1) >> 1;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    const uint32_t in_full = umma::smem_u32(bars), in_empty = umma::smem_u32(bars + 8), acc_full = umma::smem_u32(bars + 16),
                   acc_empty = umma::smem_u32(bars + 20), w_bar = umma::smem_u32(bars + 24), wpeer_bar = umma::smem_u32(bars + 25);
    const int depth = a.depth;
    if (tid == 0) {
        for (int i = 0; i < MAX_DEPTH; ++i) {
            umma::mbar_init(in_full + 8 * i, 1);
            umma::mbar_init(in_empty + 8 * i, 1);
        }
        for (int i = 0; i < N_ACC; ++i) {
            umma::mbar_init(acc_full + 8 * i, 1);
            umma::mbar_init(acc_empty + 8 * i, 2 * 128);
        }
        umma::mbar_init(w_bar, 1);
        umma::mbar_init(wpeer_bar, 1);
        umma::fence_mbar_init();
        umma::mbar_expect_tx(w_bar, (uint32_t)wbytes);
        umma::bulk_g2s(umma::smem_u32(smem + SM_W), a.wimg + (size_t)rank * wbytes, (uint32_t)wbytes, w_bar);
    }
    if (warp == MMA_WARP) tmem_alloc_pair(tmem_slot, 512);
    if (tid < 128) reinterpret_cast<float *>(smem + SM_BIAS)[tid] = tid < a.cout ? a.bias[tid] : 0.0f;
    umma::fence_before_sync();
    __syncthreads();
    cluster_sync();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    bool ok = true;

    if (warp == PRODUCER_WARP) {
        if (lane == 0) {
            int it = 0;
            for (int tp = cluster_id; tp < n_pairs; tp += n_clusters, ++it) {
                const int buf = it % depth, use = it / depth;
                if (use > 0) {   // the MMAs that read this buffer `depth` tiles ago have completed (multicast commit)
                    ok = umma::mbar_wait(in_empty + 8 * buf, (uint32_t)(use - 1) & 1u);
                    if (!ok) break;
                }
                int tile = 2 * tp + (int)rank;
                if (tile >= n_tiles) tile = n_tiles - 1;
                const int n = tile / (tiles_x * tiles_y), trem = tile - n * (tiles_x * tiles_y);
                const int ty = trem / tiles_x, y0 = ty * TH, x0 = (trem - ty * tiles_x) * TW;
                if (leader) umma::mbar_expect_tx(in_full + 8 * buf, 2u * (uint32_t)inbuf);
                tma_load_4d_pair(umma::smem_u32(smem + SM_IN + buf * inbuf), &tmap, mapa(in_full + 8 * buf, 0), 0, x0 - pad, y0 - pad,
                                 n * a.planes);
            }
        }
        ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
    } else if (warp == MMA_WARP) {
        ok = __shfl_sync(0xffffffffu, (int)umma::mbar_wait(w_bar, 0u), 0) != 0;
        if (!leader) {
            if (ok && lane == 0) mbar_arrive_cluster(mapa(wpeer_bar, 0));
        } else {
            if (ok) ok = __shfl_sync(0xffffffffu, (int)mbar_wait_cluster(wpeer_bar, 0u), 0) != 0;
            const uint32_t idesc = idesc_bf16_m256((uint32_t)a.cout_pad);
            int it = 0;
            for (int tp = cluster_id; ok && tp < n_pairs; tp += n_clusters, ++it) {
                const int buf = it % depth, abuf = it & (N_ACC - 1);
                ok = __shfl_sync(0xffffffffu, (int)mbar_wait_cluster(in_full + 8 * buf, (uint32_t)(it / depth) & 1u), 0) != 0;
                if (!ok) break;
                if (it >= N_ACC) {
                    ok = __shfl_sync(0xffffffffu, (int)mbar_wait_cluster(acc_empty + 8 * abuf, (uint32_t)((it >> 2) - 1) & 1u), 0) != 0;
                    if (!ok) break;
                }
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    const uint64_t a0 = umma::smem_desc(umma::smem_u32(smem + SM_IN + buf * inbuf), (uint32_t)plane_bytes, 128);
                    const uint64_t b0 = umma::smem_desc(umma::smem_u32(smem + SM_W), (uint32_t)(nhalf * 16), 128);
                    const uint32_t d = tbase + abuf * ACC_STRIDE;
                    const uint64_t a_kstep = (uint64_t)(2 * (plane_bytes >> 4)), b_step = (uint64_t)(wslab >> 4);
                    const int shape = a.ks * 8 + ksteps;
                    if (shape == 7 * 8 + 1) issue_tile<7, 1>(d, a0, b0, b_step, idesc);
                    else if (shape == 7 * 8 + 2) issue_tile<7, 2>(d, a0, b0, b_step, idesc);
                    else if (shape == 7 * 8 + 4) issue_tile<7, 4>(d, a0, b0, b_step, idesc);
                    else if (shape == 3 * 8 + 1) issue_tile<3, 1>(d, a0, b0, b_step, idesc);
                    else if (shape == 3 * 8 + 2) issue_tile<3, 2>(d, a0, b0, b_step, idesc);
                    else if (shape == 3 * 8 + 4) issue_tile<3, 4>(d, a0, b0, b_step, idesc);
                    else {   // any other (k, channels) combination: run-time loops
                        uint64_t bd = b0;
                        uint32_t acc = 0u;
                        for (int ky = 0; ky < a.ks; ++ky)
                            for (int kx = 0; kx < a.ks; ++kx) {
                                uint64_t ad = a0 + (uint64_t)(ky * P + kx);
                                for (int ks = 0; ks < ksteps; ++ks) {
                                    mma_bf16_pair(d, ad, bd, idesc, acc);
                                    acc = 1u;
                                    ad += a_kstep;
                                    bd += b_step;
                                }
                            }
                    }
                    commit_pair(in_empty + 8 * buf);
                    commit_pair(acc_full + 8 * abuf);
                }
                __syncwarp();
            }
        }
    } else {
        const int ew = warp - EPI_WARP0, set = ew >> 2, quarter = warp & 3;
        const float *sbias = reinterpret_cast<const float *>(smem + SM_BIAS);
        const long long plane_px = (long long)H * W;
        const uint32_t acc_empty_leader = mapa(acc_empty, 0);
        const int rounds = a.cout_pad >> 4;
        int it = 0;
        for (int tp = cluster_id; tp < n_pairs; tp += n_clusters, ++it) {
            if ((it & 1) != set) continue;
            const int abuf = it & (N_ACC - 1);
            int tile = 2 * tp + (int)rank;
            const bool tile_valid = tile < n_tiles;
            if (!tile_valid) tile = n_tiles - 1;
            const int n = tile / (tiles_x * tiles_y), trem = tile - n * (tiles_x * tiles_y);
            const int ty = trem / tiles_x, y0 = ty * TH, x0 = (trem - ty * tiles_x) * TW;
            const int m = quarter * 32 + lane, r = m >> 5, c = m & 31, gy = y0 + r, gx = x0 + c;
            const bool valid = tile_valid && c < TW && gy < H && gx < W;
            const long long pix = (long long)gy * W + gx;
            ok = __all_sync(0xffffffffu, (int)umma::mbar_wait(acc_full + 8 * abuf, (uint32_t)(it >> 2) & 1u)) != 0;
            if (!ok) break;
            umma::fence_after_sync();
            const uint32_t taddr = tbase + ((uint32_t)(quarter * 32) << 16) + abuf * ACC_STRIDE;
#pragma unroll 1
            for (int q = 0; q < rounds; ++q) {
                uint32_t o[16];
                umma::tmem_ld16(taddr + 16 * q, o);
                umma::tmem_ld_wait();
                if (valid) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int ch = 16 * q + j;
                        float t = __uint_as_float(o[j]) + sbias[ch];
                        if (a.add_nchw && ch < a.cout) t += __ldg(a.add_nchw + ((long long)n * a.cout + ch) * plane_px + pix);
                        t = t >= 0.0f ? t : t * a.slope;
                        v[j] = ch < a.cout ? t : 0.0f;
                    }
                    if (a.out_nchw) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (16 * q + j < a.cout) a.out_nchw[((long long)n * a.cout + 16 * q + j) * plane_px + pix] = v[j];
                    }
                    if (a.out_bf16) {
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                            pk[j] = *reinterpret_cast<const uint32_t *>(&b2);
                        }
                        uint4 *op = reinterpret_cast<uint4 *>(a.out_bf16) + ((long long)n * (a.cout_pad >> 3) + 2 * q) * plane_px + pix;
                        op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        op[plane_px] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
            }
            umma::fence_before_sync();
            mbar_arrive_cluster_relaxed(acc_empty_leader + 8 * abuf);
        }
    }
    if (__syncthreads_or(!ok)) {
        if (tid == 0 && err) {
            *reinterpret_cast<volatile int *>(err) = 1;
            __threadfence_system();
        }
    }
    umma::fence_before_sync();
    cluster_sync();
    if (warp == MMA_WARP) tmem_dealloc_pair(tbase, 512);
}

// OIHW fp32 [cout][cin][k][k] -> per-rank bf16 operand images [rank][tap][k-step][chunk][cout_pad / 2 rows][8 ci], zero padded
__global__ void pair_pack_kernel(const float *__restrict__ w, int cout, int cin, int taps, int cout_pad, int planes, __nv_bfloat16 *__restrict__ img)
{
    const int nhalf = cout_pad >> 1, ksteps = planes >> 1;
    const int total = 2 * taps * ksteps * 2 * nhalf * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int t = i;
        const int e = t & 7; t >>= 3;
        const int row = t % nhalf; t /= nhalf;
        const int chunk = t & 1; t >>= 1;
        const int ks = t % ksteps; t /= ksteps;
        const int tap = t % taps; t /= taps;
        const int rank = t;
        const int co = rank * nhalf + row, ci = ks * 16 + chunk * 8 + e;
        img[i] = __float2bfloat16_rn((co < cout && ci < cin) ? w[((long long)co * cin + ci) * taps + tap] : 0.0f);
    }
}

// One SpyNet level's network input (video_net.py:113-119): flow_up = 2 * bilinear_x2(flow) (align_corners=False), the second
// image warped by it (bilinear, border clamp, the flow_warp of video_net.py:32-55), and the bf16 operand records
// [im1 (3), warp(im2) (3), flow_up (2), 8 zeros] of the first convolution.  flow == nullptr: the coarsest level (zero flow).
__global__ void __launch_bounds__(256) spynet_prep_kernel(const float *__restrict__ im1, const float *__restrict__ im2, const float *__restrict__ flow,
                                                          float *__restrict__ flow_up, __nv_bfloat16 *__restrict__ rec, int N, int H, int W)
{
    const long long plane_px = (long long)H * W, total = (long long)N * plane_px;
    const int h2 = H >> 1, w2 = W >> 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / plane_px, pix = i - n * plane_px;
        const int x = (int)(pix % W), y = (int)(pix / W);
        float fx = 0.0f, fy = 0.0f;
        if (flow) {   // F.interpolate(scale 2, bilinear, align_corners=False): source coordinate (dst + 0.5) / 2 - 0.5, clamped at 0
            float sx = fmaxf((x + 0.5f) * 0.5f - 0.5f, 0.0f), sy = fmaxf((y + 0.5f) * 0.5f - 0.5f, 0.0f);
            const int x0 = (int)sx, y0 = (int)sy, x1 = min(x0 + 1, w2 - 1), y1 = min(y0 + 1, h2 - 1);
            const float ax = sx - x0, ay = sy - y0;
            const float *f0 = flow + (n * 2) * (long long)h2 * w2, *f1 = f0 + (long long)h2 * w2;
            fx = 2.0f * ((1 - ay) * ((1 - ax) * __ldg(f0 + y0 * w2 + x0) + ax * __ldg(f0 + y0 * w2 + x1)) +
                         ay * ((1 - ax) * __ldg(f0 + y1 * w2 + x0) + ax * __ldg(f0 + y1 * w2 + x1)));
            fy = 2.0f * ((1 - ay) * ((1 - ax) * __ldg(f1 + y0 * w2 + x0) + ax * __ldg(f1 + y0 * w2 + x1)) +
                         ay * ((1 - ax) * __ldg(f1 + y1 * w2 + x0) + ax * __ldg(f1 + y1 * w2 + x1)));
        }
        flow_up[(n * 2) * plane_px + pix] = fx;
        flow_up[(n * 2 + 1) * plane_px + pix] = fy;
        // backward warp with border clamp (grid_sample, align_corners=True: the normalised grid maps back to pixel x + fx)
        const float px = fminf(fmaxf(x + fx, 0.0f), (float)(W - 1)), py = fminf(fmaxf(y + fy, 0.0f), (float)(H - 1));
        const int qx0 = (int)px, qy0 = (int)py, qx1 = min(qx0 + 1, W - 1), qy1 = min(qy0 + 1, H - 1);
        const float bx = px - qx0, by = py - qy0;
        float v[16];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float *p1 = im1 + (n * 3 + c) * plane_px, *p2 = im2 + (n * 3 + c) * plane_px;
            v[c] = __ldg(p1 + pix);
            v[3 + c] = (1 - by) * ((1 - bx) * __ldg(p2 + (long long)qy0 * W + qx0) + bx * __ldg(p2 + (long long)qy0 * W + qx1)) +
                       by * ((1 - bx) * __ldg(p2 + (long long)qy1 * W + qx0) + bx * __ldg(p2 + (long long)qy1 * W + qx1));
        }
        v[6] = fx;
        v[7] = fy;
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            pk[j] = *reinterpret_cast<const uint32_t *>(&b2);
        }
        uint4 *op = reinterpret_cast<uint4 *>(rec) + (n * 2) * plane_px + pix;
        op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        op[plane_px] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// F.avg_pool2d(x, 2, 2) on [N*C, H, W] planes (the image pyramid, video_net.py:104-106)
__global__ void __launch_bounds__(256) avgpool2_kernel(const float *__restrict__ in, float *__restrict__ out, long long planes, int H, int W)
{
    const int h2 = H >> 1, w2 = W >> 1;
    const long long total = planes * h2 * w2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w2), y = (int)((i / w2) % h2);
        const long long p = i / ((long long)w2 * h2);
        const float *s = in + p * H * W + (long long)(2 * y) * W + 2 * x;
        out[i] = ((__ldg(s) + __ldg(s + 1)) + (__ldg(s + W) + __ldg(s + W + 1))) * 0.25f;
    }
}

} // namespace pair

int tc_watchdog(volatile int **host, int **dev);   // pmctf_kernels.cu
void count_launch();

typedef CUresult (*encode_tiled_fn2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn2 tensor_map_encoder2()
{
    static encode_tiled_fn2 fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn2)p;
    }
    return fn;
}

static unsigned grid_for(long long items, int per_block)
{
    long long blocks = (items + per_block - 1) / per_block;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    return (unsigned)blocks;
}

static bool pair_shape_ok(int ks, int cin_pad, int cout_pad)
{
    if (!(ks == 1 || ks == 3 || ks == 7) || cin_pad < 16 || cin_pad > 64 || (cin_pad & 15) || cout_pad < 16 || cout_pad > 128 || (cout_pad & 15))
        return false;
    const long long wbytes = (long long)ks * ks * (cin_pad / 16) * (cout_pad / 2) * 32;
    const long long inbuf = (long long)(cin_pad / 8) * (pair::TH + ks - 1) * pair::P * 16;
    return wbytes <= pair::MAX_W && inbuf <= pair::MAX_IN;
}

} // namespace pmctf

using namespace pmctf;

extern "C" {

long long pmctf_pair_packed_bytes(int ks, int cin_pad, int cout_pad)
{
    return pair_shape_ok(ks, cin_pad, cout_pad) ? 2LL * ks * ks * (cin_pad / 16) * (cout_pad / 2) * 32 : 0;
}

int pmctf_pair_pack_conv(const float *w, int cout, int cin, int ks, int cin_pad, int cout_pad, void *packed, void *stream)
{
    if (!w || !packed || cout < 1 || cin < 1 || cout > cout_pad || cin > cin_pad || !pair_shape_ok(ks, cin_pad, cout_pad) || ((uintptr_t)packed & 15))
        return PMCTF_EINVAL;
    const int total = 2 * ks * ks * (cin_pad / 16) * 2 * (cout_pad / 2) * 8;
    pair::pair_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, cout, cin, ks * ks, cout_pad, cin_pad / 8, (__nv_bfloat16 *)packed);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_pair_conv(const void *in_bf16, const void *packed_w, const float *bias, int ks, int cin_pad, int cout, int cout_pad, float slope,
                    void *out_bf16, float *out_nchw, const float *add_nchw, int N, int H, int W, void *stream)
{
    if (!in_bf16 || !packed_w || !bias || N <= 0 || H <= 0 || W <= 0 || cout < 1 || cout > cout_pad || !pair_shape_ok(ks, cin_pad, cout_pad) ||
        (!out_bf16 && !out_nchw))
        return PMCTF_EINVAL;
    if ((((uintptr_t)in_bf16 | (uintptr_t)packed_w | (uintptr_t)out_bf16) & 15) != 0 || (add_nchw && !out_nchw)) return PMCTF_EINVAL;
    const int planes = cin_pad / 8;
    if ((long long)N * planes > 0x7fffffffLL || W > (1 << 20) || H > (1 << 20)) return PMCTF_ESHAPE;
    static int configured_for[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
    if (dev < 0 || dev >= 64) return PMCTF_EINVAL;
    if (!configured_for[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pair::pair_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return (int)cudaGetLastError();
        configured_for[dev] = sms;
    }
    encode_tiled_fn2 enc = tensor_map_encoder2();
    if (!enc) return PMCTF_EINVAL;
    CUtensorMap map;
    const cuuint64_t gdim[4] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * planes};
    const cuuint64_t gstr[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    const cuuint32_t box[4] = {8, pair::P, (cuuint32_t)(pair::TH + ks - 1), (cuuint32_t)planes};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(in_bf16), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return PMCTF_ESHAPE;
    volatile int *herr = nullptr;
    int *derr = nullptr;
    int e = tc_watchdog(&herr, &derr);
    if (e) return e;
    if (herr[0] != 0) return PMCTF_ETIMEOUT;
    const int TW = pair::P - (ks - 1);
    const long long tiles = (long long)((W + TW - 1) / TW) * ((H + pair::TH - 1) / pair::TH) * N;
    if (tiles <= 0 || tiles > 0x7fffffffLL) return PMCTF_ESHAPE;
    const long long pairs = (tiles + 1) / 2, max_clusters = configured_for[dev] / 2;
    const unsigned grid = 2u * (unsigned)(pairs < max_clusters ? pairs : max_clusters);
    pair::ConvD d;
    d.wimg = (const uint8_t *)packed_w; d.bias = bias; d.out_bf16 = (__nv_bfloat16 *)out_bf16; d.out_nchw = out_nchw; d.add_nchw = add_nchw;
    d.slope = slope; d.n = N; d.h = H; d.w = W; d.ks = ks; d.planes = planes; d.cout = cout; d.cout_pad = cout_pad;
    const int inbuf = planes * (pair::TH + ks - 1) * pair::P * 16;
    d.depth = pair::N_IN * pair::MAX_IN / inbuf;   // small input tiles (few channels) get a deeper ring: the TMA latency, not the MMAs, paces them
    if (d.depth > pair::MAX_DEPTH) d.depth = pair::MAX_DEPTH;
    pair::pair_conv_kernel<<<grid, pair::NT, pair::SMEM_BYTES, (cudaStream_t)stream>>>(map, d, derr);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_spynet_prep(const float *im1, const float *im2, const float *flow, float *flow_up, void *rec_bf16, int N, int H, int W, void *stream)
{
    if (!im1 || !im2 || !flow_up || !rec_bf16 || N <= 0 || H <= 0 || W <= 0 || ((uintptr_t)rec_bf16 & 15)) return PMCTF_EINVAL;
    if (flow && ((H | W) & 1)) return PMCTF_ESHAPE;
    pair::spynet_prep_kernel<<<grid_for((long long)N * H * W, 256), 256, 0, (cudaStream_t)stream>>>(im1, im2, flow, flow_up, (__nv_bfloat16 *)rec_bf16, N,
                                                                                                  H, W);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_avgpool2(const float *in, float *out, long long planes, int H, int W, void *stream)
{
    if (!in || !out || planes <= 0 || H < 2 || W < 2 || ((H | W) & 1)) return PMCTF_EINVAL;
    pair::avgpool2_kernel<<<grid_for(planes * (H / 2) * (W / 2), 256), 256, 0, (cudaStream_t)stream>>>(in, out, planes, H, W);
    count_launch();
    return (int)cudaGetLastError();
}

} // extern "C"
