// pmctf_ctx.cu -- the four-step entropy-parameter network of pWave++:
// pMCTF/layers/context_fusion_4step.py:9-249 (ContextResidual, ContextFusionFourStep), its DepthConvBlock head
// (pMCTF/layers/video/layers.py:113-172) and the masked quantiser `process_with_mask` (:115-125) that owns the final
// symbol round((s - mu) * mask).  Called once per coded subband at pWave.py:259-290 (forward), :405-421 (compress),
// :500-512 (decompress); 22 dense 112 -> 112 3x3 convolutions + one 1x1 per subband = 4.97 MFLOP per coefficient, 80 % of
// all the codec's FLOPs.
//
// The 112 -> 112 layers run as implicit GEMMs on the 5th-generation tensor cores in CTA-PAIR mode (tcgen05.mma
// cta_group::2, M = 256, N = 112, K = 16, bf16 operands, fp32 accumulators in TMEM):
//   * a pair of CTAs (cluster of 2 = one TPC) works on two pixel tiles at once; each CTA stages ITS tile (128 pixel records
//     per 8-channel plane, the A operand rows of its half of M) and holds HALF of the weights (56 of the 112 output channels,
//     the B operand rows of its half of N): all 9 x 7 weight slabs of a layer stay resident in shared memory (113 KB per
//     CTA) -- with one CTA per tile they would not fit (226 KB) and would have to be re-streamed per tile;
//   * per-SM operand traffic per MMA is 128 x 32 B (A) + 56 x 32 B (B) = 5.9 KB = 46 cycles at 128 B / cycle against 56 cycles
//     of math: the layer is bounded by the tensor pipe, not by the shared-memory operand fetch (a single-CTA N = 64 split
//     would need 48 cycles of fetch for 32 of math);
//   * input tiles arrive by TMA tensor loads (cp.async.bulk.tensor.4d, one instruction per tile and CTA, zero fill outside the
//     image = the convolution's zero padding) straight into the K-major un-swizzled operand layout: feature maps live in HBM
//     as [N][C/8][H][W][8] bf16, so a box {8 ch, 32 px, 6 rows, 14 planes} IS the array of 16-byte pixel records the
//     descriptors of pmctf_lift_tc.cu / pmctf_pp.cu walk (a filter tap = a different descriptor start address);
//   * warp roles per CTA: 1 TMA producer, 1 MMA issuer (leader CTA only issues; completion is multicast to both CTAs with
//     tcgen05.commit.cta_group::2), 8 epilogue warps in two sets that alternate tiles (TMEM -> bias / skip / LeakyReLU ->
//     fp32 and bf16 feature maps, coalesced: lanes = consecutive pixels); 2 input buffers, 4 accumulator buffers.
// Everything else of the module (1 -> 112 input convolutions, the depthwise head, the 112 -> 2 projections, the masked
// quantiser with the int16 symbol / table-index staging for the rANS coder) is small CUDA-core kernels in this file.
//
// Numerics: bf16 operands (RN), fp32 accumulation in the tensor core's order, fp32 biases / skips / residual stream.  Not
// bit-exact against an fp32 evaluation by construction; tests/test_gpu_ctx.py reports the count of final symbols that differ
// from the fp32 oracle (oracle/ctx_oracle.py, pinned to the reference's module) next to the parameter errors.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pmctf_b200.h"
#include "pmctf_umma.cuh"

namespace pmctf {
namespace ctx {

constexpr int C = 112;                               // channels of every tensor-core layer (context_fusion_4step.py / pWave.py:70)
constexpr int KS = C / 16, CH = C / 8, CQ = C / 4;   // k-steps per tap, 8-channel operand planes, float4 groups
constexpr int TH = 4, TW = 30, P = 32;               // output tile per CTA; pixel pitch of the staged input (TW + 2)
constexpr int IN_R = TH + 2;
constexpr int PLANE = IN_R * P * 16;                 // bytes of one 8-channel plane of a staged tile
constexpr int INBUF = CH * PLANE;                    // 43 008 B = one TMA box
constexpr int NHALF = C / 2;                         // B rows (output channels) held by one CTA of the pair
constexpr int WSLAB = 2 * NHALF * 16;                // bytes of one (tap, k-step) weight slab per CTA: [chunk][56 rows][8 ci]
constexpr int WHALF_MAX = 9 * KS * WSLAB;            // 112 896 B
constexpr int N_IN = 2, N_ACC = 4, ACC_STRIDE = 128; // input / accumulator ring depths, TMEM columns per accumulator
constexpr int SM_W = 0;
constexpr int SM_IN = SM_W + WHALF_MAX;
constexpr int SM_BAR = SM_IN + N_IN * INBUF;
constexpr int SM_BIAS = SM_BAR + 128;
constexpr int SM_HEAD = SM_BIAS + C * 4;             // 2 x 112 floats of the fused projection
constexpr int SMEM_BYTES = SM_HEAD + 2 * C * 4;
static_assert(TH * P == 128, "one 128-row MMA block per CTA and tile");
static_assert(SM_IN % 128 == 0 && INBUF % 128 == 0 && SM_BAR % 8 == 0 && SMEM_BYTES <= 227 * 1024, "shared memory");
static_assert(N_ACC * ACC_STRIDE <= 512 && C <= ACC_STRIDE, "TMEM columns");

constexpr int EPI_SETS = 2, EPI_WARPS = 4 * EPI_SETS;
constexpr int PRODUCER_WARP = 0, MMA_WARP = 1, EPI_WARP0 = 2;
constexpr int NT = 32 * (EPI_WARP0 + EPI_WARPS);

// The same geometry for any channel count that is a multiple of 16: 112 (the entropy-parameter network, the constants above) and
// 64 (PostProcess, csrc/pmctf_pp.cu, whose 64 -> 64 layers run through this kernel as well).
template <int CC>
struct Geo {
    static constexpr int C = CC, KS = CC / 16, CH = CC / 8, CQ = CC / 4;
    static constexpr int PLANE = IN_R * P * 16, INBUF = CH * PLANE, NHALF = CC / 2, WSLAB = 2 * NHALF * 16, WHALF_MAX = 9 * KS * WSLAB;
    static constexpr int SM_W = 0, SM_IN = SM_W + WHALF_MAX, SM_BAR = SM_IN + N_IN * INBUF, SM_BIAS = SM_BAR + 128, SM_HEAD = SM_BIAS + CC * 4;
    static constexpr int SMEM_BYTES = SM_HEAD + 2 * CC * 4;
    static_assert(SM_IN % 128 == 0 && INBUF % 128 == 0 && SM_BAR % 8 == 0 && SMEM_BYTES <= 227 * 1024 && CC % 16 == 0 && CC <= ACC_STRIDE, "geometry");
};

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 256 (CTA pair), N = n
__host__ __device__ constexpr uint32_t idesc_bf16_m256(uint32_t n)
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
// address of the same shared-memory location in CTA `rank` of the cluster (shared::cluster window)
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
// "this thread has finished reading its accumulator rows" (the tcgen05.wait::ld before it has completed the reads): no memory
// ordering is carried, so the arrive must not wait for the thread's outstanding global stores (a release at cluster scope would)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// bounded wait that synchronises with arrivals from the peer CTA
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
// all MMAs issued so far by this thread arrive (once complete) on the barrier at this offset in BOTH CTAs of the pair
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
// TMA tensor load of this CTA's box into ITS shared memory; the bytes are counted on the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst_saddr, const CUtensorMap *map, uint32_t leader_mbar_cluster_addr, int c0,
                                                 int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            dst_saddr),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

struct ConvD {
    const uint8_t *wimg;          // packed operand image: [rank][tap][k-step][chunk][56 rows][8 ci] bf16
    const float *bias;            // [112]
    const float *res, *res2;      // fp32 [N][28][H][W][4] added after the bias (skip connections), or null
    float *out_f32;               // fp32 [N][28][H][W][4], or null
    __nv_bfloat16 *out_bf16;      // bf16 [N][14][H][W][8], or null
    const float *head_w, *head_b; // optional fused 112 -> 2 projection of the layer's output (y_spatial_prior_k_out.2): [2][112], [2]
    float *head_scales, *head_means;   // [N,1,H,W] each
    float slope;                  // LeakyReLU slope applied last (1 = identity)
    int n, h, w, taps;            // taps: 9 (3x3, padding 1) or 1 (1x1)
};

template <int CC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1)
    ctx_conv_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvD a, int *__restrict__ err)
{
    using G = Geo<CC>;   // the names below shadow the namespace-level constants of the 112-channel case
    constexpr int C = G::C, KS = G::KS, CH = G::CH, CQ = G::CQ, PLANE = G::PLANE, INBUF = G::INBUF, NHALF = G::NHALF, WSLAB = G::WSLAB;
    constexpr int SM_W = G::SM_W, SM_IN = G::SM_IN, SM_BAR = G::SM_BAR, SM_BIAS = G::SM_BIAS, SM_HEAD = G::SM_HEAD;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    // barriers: in_full[2] (leader's are used), in_empty[2], acc_full[4], acc_empty[4] (leader's are used), weights, peer-weights
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_BAR + 120);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int H = a.h, W = a.w;
    const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
    const int n_tiles = tiles_x * tiles_y * a.n;
    const int n_pairs = (n_tiles + 1) >> 1;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int wbytes = a.taps * KS * WSLAB;

    const uint32_t in_full = umma::smem_u32(bars), in_empty = umma::smem_u32(bars + 2), acc_full = umma::smem_u32(bars + 4),
                   acc_empty = umma::smem_u32(bars + 8), w_bar = umma::smem_u32(bars + 12), wpeer_bar = umma::smem_u32(bars + 13);
    if (tid == 0) {
        for (int i = 0; i < N_IN; ++i) {
            umma::mbar_init(in_full + 8 * i, 1);                  // leader producer's arrive.expect_tx (+ the bytes of both boxes)
            umma::mbar_init(in_empty + 8 * i, 1);                 // tcgen05.commit (multicast)
        }
        for (int i = 0; i < N_ACC; ++i) {
            umma::mbar_init(acc_full + 8 * i, 1);                 // tcgen05.commit (multicast)
            umma::mbar_init(acc_empty + 8 * i, 2 * 128);          // one epilogue set of each CTA
        }
        umma::mbar_init(w_bar, 1);
        umma::mbar_init(wpeer_bar, 1);
        umma::fence_mbar_init();
        umma::mbar_expect_tx(w_bar, (uint32_t)wbytes);            // this CTA's half of the layer's weights
        umma::bulk_g2s(umma::smem_u32(smem + SM_W), a.wimg + (size_t)rank * wbytes, (uint32_t)wbytes, w_bar);
    }
    if (warp == MMA_WARP) tmem_alloc_pair(tmem_slot, 512);
    if (tid < C) reinterpret_cast<float *>(smem + SM_BIAS)[tid] = a.bias[tid];
    if (a.head_w && tid < 2 * C) reinterpret_cast<float *>(smem + SM_HEAD)[tid] = a.head_w[tid];
    umma::fence_before_sync();
    __syncthreads();
    cluster_sync();   // both CTAs' barriers are initialised before anything signals across the pair
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    bool ok = true;

    if (warp == PRODUCER_WARP) {
        // ---- TMA producer (both CTAs): one box per tile into this CTA's buffer, counted on the leader's barrier --------------
        if (lane == 0) {
            int it = 0;
            for (int tp = cluster_id; tp < n_pairs; tp += n_clusters, ++it) {
                const int buf = it & 1;
                if (it >= N_IN) {   // the MMAs that read this buffer have completed (multicast commit)
                    ok = umma::mbar_wait(in_empty + 8 * buf, (uint32_t)((it >> 1) - 1) & 1u);
                    if (!ok) break;
                }
                int tile = 2 * tp + (int)rank;
                if (tile >= n_tiles) tile = n_tiles - 1;   // odd tile count: the peer repeats the last tile, its stores are skipped
                const int n = tile / (tiles_x * tiles_y), trem = tile - n * (tiles_x * tiles_y);
                const int ty = trem / tiles_x, y0 = ty * TH, x0 = (trem - ty * tiles_x) * TW;
                if (leader) umma::mbar_expect_tx(in_full + 8 * buf, 2u * INBUF);
                const int pad = a.taps == 9 ? 1 : 0;
                tma_load_4d_pair(umma::smem_u32(smem + SM_IN + buf * INBUF), &tmap, mapa(in_full + 8 * buf, 0), 0, x0 - pad, y0 - pad, n * CH);
            }
        }
        ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
    } else if (warp == MMA_WARP) {
        // ---- MMA issue (leader CTA) ----------------------------------------------------------------------------------------
        ok = __shfl_sync(0xffffffffu, (int)umma::mbar_wait(w_bar, 0u), 0) != 0;   // this CTA's weights have landed
        if (!leader) {
            if (ok && lane == 0) mbar_arrive_cluster(mapa(wpeer_bar, 0));          // tell the leader
        } else {
            if (ok) ok = __shfl_sync(0xffffffffu, (int)mbar_wait_cluster(wpeer_bar, 0u), 0) != 0;
            int it = 0;
            for (int tp = cluster_id; ok && tp < n_pairs; tp += n_clusters, ++it) {
                const int buf = it & 1, abuf = it & (N_ACC - 1);
                ok = __shfl_sync(0xffffffffu, (int)mbar_wait_cluster(in_full + 8 * buf, (uint32_t)(it >> 1) & 1u), 0) != 0;
                if (!ok) break;
                if (it >= N_ACC) {   // both CTAs' epilogues have drained this accumulator
                    ok = __shfl_sync(0xffffffffu, (int)mbar_wait_cluster(acc_empty + 8 * abuf, (uint32_t)((it >> 2) - 1) & 1u), 0) != 0;
                    if (!ok) break;
                }
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    const uint64_t a0 = umma::smem_desc(umma::smem_u32(smem + SM_IN + buf * INBUF), PLANE, 128);
                    const uint64_t b0 = umma::smem_desc(umma::smem_u32(smem + SM_W), NHALF * 16, 128);
                    const uint32_t d = tbase + abuf * ACC_STRIDE;
                    if (a.taps == 9) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const int off = (tap / 3) * P + (tap % 3);
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks)
                                mma_bf16_pair(d, a0 + (uint64_t)((2 * ks) * (PLANE / 16) + off), b0 + (uint64_t)((tap * KS + ks) * (WSLAB / 16)),
                                              idesc_bf16_m256(C), (tap | ks) ? 1u : 0u);
                        }
                    } else {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks)
                            mma_bf16_pair(d, a0 + (uint64_t)((2 * ks) * (PLANE / 16)), b0 + (uint64_t)(ks * (WSLAB / 16)), idesc_bf16_m256(C),
                                          ks ? 1u : 0u);
                    }
                    commit_pair(in_empty + 8 * buf);     // both CTAs' input buffers are free once these MMAs have completed ...
                    commit_pair(acc_full + 8 * abuf);    // ... and both CTAs' accumulators are ready
                }
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue (both CTAs): TMEM -> bias / skips / LeakyReLU -> global; two warp sets alternate tiles ---------------
        const int ew = warp - EPI_WARP0, set = ew >> 2, quarter = warp & 3;   // a warp may only touch TMEM lanes 32 * (warp % 4) ...
        const float *sbias = reinterpret_cast<const float *>(smem + SM_BIAS);
        const float *shead = reinterpret_cast<const float *>(smem + SM_HEAD);
        const long long plane_px = (long long)H * W;
        const uint32_t acc_empty_leader = mapa(acc_empty, 0);
        const int pad_shift = a.taps == 9 ? 0 : 0;
        (void)pad_shift;
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
            const bool use_res = a.res != nullptr && valid, use_res2 = a.res2 != nullptr && valid;
            const float4 *rp = reinterpret_cast<const float4 *>(a.res) + (long long)n * CQ * plane_px + pix;
            const float4 *rp2 = reinterpret_cast<const float4 *>(a.res2) + (long long)n * CQ * plane_px + pix;
            // the skip operands of this tile start their way from HBM to L2 while the MMAs of the tile are still running ...
            if (use_res)
                for (int j = 4; j < CQ; ++j) prefetch_l2(rp + j * plane_px);
            if (use_res2)
                for (int j = 4; j < CQ; ++j) prefetch_l2(rp2 + j * plane_px);
            float4 rr[4], rr2[4];
            // ... and those of the first 16 channels are requested before the accumulators are waited for
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                rr[j] = use_res ? __ldg(rp + j * plane_px) : make_float4(0.f, 0.f, 0.f, 0.f);
                rr2[j] = use_res2 ? __ldg(rp2 + j * plane_px) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            ok = __all_sync(0xffffffffu, (int)umma::mbar_wait(acc_full + 8 * abuf, (uint32_t)(it >> 2) & 1u)) != 0;
            if (!ok) break;
            umma::fence_after_sync();
            const uint32_t taddr = tbase + ((uint32_t)(quarter * 32) << 16) + abuf * ACC_STRIDE;
            float hs = 0.0f, hm = 0.0f;      // fused projection: one fma chain per output over the channels in ascending order
            if (a.head_w) {
                hs = __ldg(a.head_b);
                hm = __ldg(a.head_b + 1);
            }
#pragma unroll 1
            for (int q = 0; q < KS; ++q) {   // 16 channels per round
                uint32_t o[16];
                umma::tmem_ld16(taddr + 16 * q, o);
                float4 nr[4], nr2[4];
                if (q + 1 < KS) {   // next round's skip operands in flight during this round's arithmetic and stores
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        nr[j] = use_res ? __ldg(rp + (4 * (q + 1) + j) * plane_px) : make_float4(0.f, 0.f, 0.f, 0.f);
                        nr2[j] = use_res2 ? __ldg(rp2 + (4 * (q + 1) + j) * plane_px) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                umma::tmem_ld_wait();
                if (valid) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bb = *reinterpret_cast<const float4 *>(sbias + 16 * q + 4 * j);
                        v[4 * j] = ((__uint_as_float(o[4 * j]) + bb.x) + rr[j].x) + rr2[j].x;
                        v[4 * j + 1] = ((__uint_as_float(o[4 * j + 1]) + bb.y) + rr[j].y) + rr2[j].y;
                        v[4 * j + 2] = ((__uint_as_float(o[4 * j + 2]) + bb.z) + rr[j].z) + rr2[j].z;
                        v[4 * j + 3] = ((__uint_as_float(o[4 * j + 3]) + bb.w) + rr[j].w) + rr2[j].w;
                    }
                    if (a.slope != 1.0f) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = v[j] >= 0.0f ? v[j] : v[j] * a.slope;
                    }
                    if (a.head_w) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            hs = fmaf(shead[16 * q + j], v[j], hs);
                            hm = fmaf(shead[C + 16 * q + j], v[j], hm);
                        }
                    }
                    if (a.out_f32) {
                        float4 *op = reinterpret_cast<float4 *>(a.out_f32) + ((long long)n * CQ + 4 * q) * plane_px + pix;
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
                        uint4 *op = reinterpret_cast<uint4 *>(a.out_bf16) + ((long long)n * CH + 2 * q) * plane_px + pix;
                        op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        op[plane_px] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    rr[j] = nr[j];
                    rr2[j] = nr2[j];
                }
            }
            if (a.head_w && valid) {
                a.head_scales[(long long)n * plane_px + pix] = hs;
                a.head_means[(long long)n * plane_px + pix] = hm;
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
    cluster_sync();   // the peer may still be reading accumulators / receiving multicast arrivals
    if (warp == MMA_WARP) tmem_dealloc_pair(tbase, 512);
}

// OIHW fp32 [112][112][k][k] -> per-rank bf16 operand images [rank][tap][k-step][chunk][56 rows][8 ci]
template <int CC>
__global__ void ctx_pack_kernel(const float *__restrict__ w, int taps, __nv_bfloat16 *__restrict__ img)
{
    constexpr int C = CC, KS = CC / 16, NHALF = CC / 2;
    const int total = 2 * taps * KS * 2 * NHALF * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int t = i;
        const int e = t & 7; t >>= 3;
        const int row = t % NHALF; t /= NHALF;
        const int chunk = t & 1; t >>= 1;
        const int ks = t % KS; t /= KS;
        const int tap = t % taps; t /= taps;
        const int rank = t;
        const int co = rank * NHALF + row, ci = ks * 16 + chunk * 8 + e;
        img[i] = __float2bfloat16_rn(w[((long long)co * C + ci) * taps + tap]);
    }
}

// 1 or 2 single-channel planes -> 112 channels, 3x3, zero padding, on the CUDA cores (conv1_context :47, y_spatial_prior_k.0 :62-86):
// writes the fp32 feature map and its bf16 operand copy.  One thread = one pixel x 16 output channels.
template <int CIN>
__global__ void __launch_bounds__(256) ctx_conv_in_kernel(const float *__restrict__ x0, const float *__restrict__ x1,
                                                          const float *__restrict__ w, const float *__restrict__ b,
                                                          float *__restrict__ out_f32, __nv_bfloat16 *__restrict__ out_bf16, int N, int H, int W)
{
    __shared__ float sw[C * CIN * 9], sb[C];
    for (int i = threadIdx.x; i < C * CIN * 9; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < C) sb[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    const long long plane_px = (long long)H * W, all_px = (long long)N * plane_px, total = all_px * KS;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i / all_px);
        const long long apix = i - q * all_px;
        const long long n = apix / plane_px, pix = apix - n * plane_px;
        const int gx = (int)(pix % W), gy = (int)(pix / W);
        float v[CIN][9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int yy = gy + k / 3 - 1, xx = gx + k % 3 - 1;
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
            v[0][k] = in ? __ldg(x0 + n * plane_px + (long long)yy * W + xx) : 0.0f;
            if (CIN == 2) v[CIN - 1][k] = in ? __ldg(x1 + n * plane_px + (long long)yy * W + xx) : 0.0f;
        }
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float acc = sb[16 * q + j];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                for (int k = 0; k < 9; ++k) acc = fmaf(sw[((16 * q + j) * CIN + ci) * 9 + k], v[ci][k], acc);
            o[j] = acc;
        }
        float4 *of = reinterpret_cast<float4 *>(out_f32) + (n * CQ + 4 * q) * plane_px + pix;
#pragma unroll
        for (int j = 0; j < 4; ++j) of[j * plane_px] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
            pk[j] = *reinterpret_cast<const uint32_t *>(&b2);
        }
        uint4 *op = reinterpret_cast<uint4 *>(out_bf16) + (n * CH + 2 * q) * plane_px + pix;
        op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        op[plane_px] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
}

// lower_level_subband (:49-52): nearest-neighbour x2 upsampling of the coarser level's subband, then a 1 -> 1 3x3 convolution
__global__ void __launch_bounds__(256) ctx_lower_kernel(const float *__restrict__ prev, const float *__restrict__ w, const float *__restrict__ b,
                                                        float *__restrict__ out, int N, int h, int wd)
{
    const int H = 2 * h, W = 2 * wd;
    const long long total = (long long)N * H * W;
    float k9[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) k9[k] = __ldg(w + k);
    const float bias = __ldg(b);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const long long n = i / ((long long)W * H);
        float acc = bias;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
            const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(prev + (n * h + (yy >> 1)) * wd + (xx >> 1)) : 0.0f;
            acc = fmaf(k9[k], v, acc);
        }
        out[i] = acc;
    }
}

struct DcbD {   // the tail of DepthConvBlock(112, 2) (layers/video/layers.py:113-172) behind its first 1x1 convolution
    const float *dw_w, *dw_b;        // depth_conv  [112][1][3][3], [112]
    const float *pw_w, *pw_b;        // conv2       [2][112], [2]
    const float *ad_w, *ad_b;        // adaptor     [2][112], [2]
    const float *f1_w, *f1_b;        // ConvFFN.conv.0  [8][2], [8]
    const float *f2_w, *f2_b;        // ConvFFN.conv.2  [2][8], [2]
};

// t1 = LeakyReLU_0.01(conv1(ctx)) (tensor-core layer) -> depthwise 3x3 -> 1x1 to 2 channels, + adaptor(ctx), then the FFN
// (identity + LeakyReLU_0.1(W2 LeakyReLU_0.1(W1 u + b1) + b2)).  One block = a 32 x 4 pixel tile, one thread per pixel; the channels
// walk in float4 groups, each group's 34 x 6 halo tile staged in shared memory (double buffered) so that a pixel's nine taps are
// shared-memory reads and every t1 value is fetched from global memory 1.6 instead of 9 times.
constexpr int DT_W = 32, DT_H = 4, DT_PW = DT_W + 2, DT_PH = DT_H + 2;
__global__ void __launch_bounds__(DT_W * DT_H) ctx_dcb_tail_kernel(const float *__restrict__ t1, const float *__restrict__ cx, const DcbD p,
                                                                  float *__restrict__ scales, float *__restrict__ means, int N, int H, int W)
{
    __shared__ float s_dw[C * 9], s_dwb[C], s_pw[2 * C], s_ad[2 * C];
    __shared__ float4 s_t[2][DT_PH * DT_PW];
    const int tid = threadIdx.x;
    for (int i = tid; i < C * 9; i += blockDim.x) s_dw[i] = p.dw_w[i];
    for (int i = tid; i < C; i += blockDim.x) s_dwb[i] = p.dw_b[i];
    for (int i = tid; i < 2 * C; i += blockDim.x) {
        s_pw[i] = p.pw_w[i];
        s_ad[i] = p.ad_w[i];
    }
    const int tiles_x = (W + DT_W - 1) / DT_W, tiles_y = (H + DT_H - 1) / DT_H;
    const long long plane_px = (long long)H * W, n_tiles = (long long)tiles_x * tiles_y * N;
    const int lx = tid & (DT_W - 1), ly = tid >> 5;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long n = tile / ((long long)tiles_x * tiles_y);
        const int trem = (int)(tile - n * tiles_x * tiles_y), ty = trem / tiles_x, y0 = ty * DT_H, x0 = (trem - ty * tiles_x) * DT_W;
        const int gx = x0 + lx, gy = y0 + ly;
        const bool valid = gx < W && gy < H;
        const float4 *tp = reinterpret_cast<const float4 *>(t1) + n * CQ * plane_px;
        const float4 *cp = reinterpret_cast<const float4 *>(cx) + n * CQ * plane_px + (long long)gy * W + gx;
        auto stage = [&](int g, int buf) {
            for (int i = tid; i < DT_PH * DT_PW; i += DT_W * DT_H) {
                const int r = i / DT_PW, c = i - r * DT_PW, yy = y0 - 1 + r, xx = x0 - 1 + c;
                s_t[buf][i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(tp + g * plane_px + (long long)yy * W + xx)
                                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        float u0 = 0.0f, u1 = 0.0f, a0 = 0.0f, a1 = 0.0f;
        __syncthreads();          // weights staged / previous tile's buffers free
        stage(0, 0);
        for (int g = 0; g < CQ; ++g) {
            __syncthreads();      // group g is in s_t[g & 1]; everybody has left group g - 1's buffer
            if (g + 1 < CQ) stage(g + 1, (g + 1) & 1);
            const float4 *st = s_t[g & 1];
            float d[4] = {s_dwb[4 * g], s_dwb[4 * g + 1], s_dwb[4 * g + 2], s_dwb[4 * g + 3]};
#pragma unroll
            for (int k = 0; k < 9; ++k) {   // zero padding is in the staged tile: the chain order (ky, kx) matches the per-tap loop
                const float4 v = st[(ly + k / 3) * DT_PW + lx + k % 3];
                const int yy = gy + k / 3 - 1, xx = gx + k % 3 - 1;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;   // keep the arithmetic identical to skipping the tap
                d[0] = fmaf(s_dw[(4 * g) * 9 + k], v.x, d[0]);
                d[1] = fmaf(s_dw[(4 * g + 1) * 9 + k], v.y, d[1]);
                d[2] = fmaf(s_dw[(4 * g + 2) * 9 + k], v.z, d[2]);
                d[3] = fmaf(s_dw[(4 * g + 3) * 9 + k], v.w, d[3]);
            }
            const float4 cv = valid ? __ldg(cp + g * plane_px) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float c4[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                u0 = fmaf(s_pw[4 * g + j], d[j], u0);
                u1 = fmaf(s_pw[C + 4 * g + j], d[j], u1);
                a0 = fmaf(s_ad[4 * g + j], c4[j], a0);
                a1 = fmaf(s_ad[C + 4 * g + j], c4[j], a1);
            }
        }
        if (!valid) continue;
        const float y0v = (u0 + __ldg(p.pw_b)) + (a0 + __ldg(p.ad_b)), y1v = (u1 + __ldg(p.pw_b + 1)) + (a1 + __ldg(p.ad_b + 1));
        float hdn[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = fmaf(__ldg(p.f1_w + 2 * j + 1), y1v, fmaf(__ldg(p.f1_w + 2 * j), y0v, __ldg(p.f1_b + j)));
            hdn[j] = t >= 0.0f ? t : t * 0.1f;
        }
        float z0 = __ldg(p.f2_b), z1 = __ldg(p.f2_b + 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            z0 = fmaf(__ldg(p.f2_w + j), hdn[j], z0);
            z1 = fmaf(__ldg(p.f2_w + 8 + j), hdn[j], z1);
        }
        z0 = z0 >= 0.0f ? z0 : z0 * 0.1f;
        z1 = z1 >= 0.0f ? z1 : z1 * 0.1f;
        const long long o = n * plane_px + (long long)gy * W + gx;
        scales[o] = y0v + z0;     // chunk(2, dim=1): channel 0 = scales, channel 1 = means (:172)
        means[o] = y1v + z1;
    }
}

// the 112 -> 2 projection that ends y_spatial_prior_k_out (:66-70): 1x1 convolution on the fp32 feature map
__global__ void __launch_bounds__(256) ctx_head_kernel(const float *__restrict__ f, const float *__restrict__ w, const float *__restrict__ b,
                                                       float *__restrict__ scales, float *__restrict__ means, int N, int H, int W)
{
    __shared__ float sw[2 * C];
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const long long plane_px = (long long)H * W, total = (long long)N * plane_px;
    const float b0 = __ldg(b), b1 = __ldg(b + 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / plane_px, pix = i - n * plane_px;
        const float4 *fp = reinterpret_cast<const float4 *>(f) + n * CQ * plane_px + pix;
        float s = b0, m = b1;
#pragma unroll 4
        for (int g = 0; g < CQ; ++g) {
            const float4 v = __ldg(fp + g * plane_px);
            s = fmaf(sw[4 * g + 3], v.w, fmaf(sw[4 * g + 2], v.z, fmaf(sw[4 * g + 1], v.y, fmaf(sw[4 * g], v.x, s))));
            m = fmaf(sw[C + 4 * g + 3], v.w, fmaf(sw[C + 4 * g + 2], v.z, fmaf(sw[C + 4 * g + 1], v.y, fmaf(sw[C + 4 * g], v.x, m))));
        }
        scales[i] = s;
        means[i] = m;
    }
}

struct StepD {
    const float *x;          // the subband (encoder), or null (decoder)
    const short *dec_sym;    // decoded symbols of this step as int16 [N*H*W] (decoder), or null
    const float *scales, *means;
    float *x_hat, *x_q, *s_hat, *x_res;     // accumulated over the four steps (x_res / x_q may be null)
    short *sym16, *idx16;                   // this step's full planes for the entropy coder (zero / index 0 off the mask), or null
    float log_min, log_step, top;
    int step, lossy, first;
    int N, H, W;
};

// process_with_mask (:115-125) for step k on the positions of mask k (2x2 pattern: k = 2 * (y & 1) + (x & 1)), writing the
// running sums x_hat_so_far / x_q / s_hat / x_res position-wise (the masks partition the plane, so the reference's sums of masked
// planes are selections), plus the int16 symbols and scale-table indexes GaussianEncoder.encode would derive from the step's
// masked planes (entropy_models.py:37-40,266-275).  Decoder form (:209-247): x_hat = (symbol + mean) on the mask.
__global__ void __launch_bounds__(256) ctx_mask_step_kernel(const StepD d)
{
    const long long plane_px = (long long)d.H * d.W, total = (long long)d.N * plane_px;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i % plane_px;
        const int x = (int)(pix % d.W), y = (int)(pix / d.W);
        const bool on = (2 * (y & 1) + (x & 1)) == d.step;
        short s16 = 0, i16 = 0;
        if (on) {
            const float sc = __ldg(d.scales + i);
            float q = 0.0f;
            if (d.x || d.dec_sym) {
                float mean = __ldg(d.means + i);
                if (!d.lossy) mean = rintf(mean);
                float res;
                if (d.x) {
                    res = __ldg(d.x + i) - mean;
                    q = rintf(res);
                } else {
                    q = (float)__ldg(d.dec_sym + i);
                    res = q;
                }
                d.x_hat[i] = q + mean;
                if (d.x_q) d.x_q[i] = q;
                if (d.x_res) d.x_res[i] = res;
                d.s_hat[i] = sc;
            }
            if (d.idx16) {
                float v = (logf(fmaxf(sc, 1e-5f)) - d.log_min) / d.log_step;
                v = fminf(fmaxf(v, 0.0f), d.top);
                i16 = (short)(int)v;
                s16 = (short)(int)fminf(fmaxf(q, -30000.0f), 30000.0f);
            }
        } else if (d.first && (d.x || d.dec_sym)) {   // step 0 initialises the running planes
            d.x_hat[i] = 0.0f;
            if (d.x_q) d.x_q[i] = 0.0f;
            if (d.x_res) d.x_res[i] = 0.0f;
            d.s_hat[i] = 0.0f;
        }
        if (d.idx16) {
            if (!on) {   // a zero scale maps to max(0, 1e-5) -> table index 0 in the reference
                float v = (logf(1e-5f) - d.log_min) / d.log_step;
                v = fminf(fmaxf(v, 0.0f), d.top);
                i16 = (short)(int)v;
            }
            d.idx16[i] = i16;
            if (d.sym16) d.sym16[i] = s16;
        }
    }
}

// ---- element-wise glue of the coder around the networks, one pass each instead of 8 / 25 ATen launches per subband ---------------
// ConvLSTM cell behind its two convolutions (long_context.py:16-34): a = conv_in(x) + conv_hidden(h); all three gates are sigmoid(a),
// the candidate is tanh(a): c' = s c + s tanh(a), h' = s tanh(c')
__global__ void __launch_bounds__(256) lstm_gates_kernel(const float *__restrict__ a_in, const float *__restrict__ a_hid, const float *__restrict__ c,
                                                         float *__restrict__ h_out, float *__restrict__ c_out, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = __ldg(a_in + i) + __ldg(a_hid + i);
        const float s = 1.0f / (1.0f + expf(-a));
        const float cn = s * __ldg(c + i) + s * tanhf(a);
        c_out[i] = cn;
        h_out[i] = s * tanhf(cn);
    }
}
// rate estimate of a band (gaussian_model.py:37-55 with torch.distributions.Laplace.cdf): P = cdf(y + .5) - cdf(y - .5),
// cdf(v) = .5 - .5 sign(v) expm1(-|v| / b), b = clamp(sigma, 1e-5, 1e10); bits = max(-log2(P + 1e-5), 0)
__global__ void __launch_bounds__(256) laplace_bits_kernel(const float *__restrict__ y, const float *__restrict__ sigma, float *__restrict__ bits,
                                                           long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float b = fminf(fmaxf(__ldg(sigma + i), 1e-5f), 1e10f), v = __ldg(y + i);
        const float hi = v + 0.5f, lo = v - 0.5f;
        const float shi = (hi > 0.0f) - (hi < 0.0f), slo = (lo > 0.0f) - (lo < 0.0f);
        const float chi = 0.5f - 0.5f * shi * expm1f(-fabsf(hi) / b), clo = 0.5f - 0.5f * slo * expm1f(-fabsf(lo) / b);
        const float t = -1.0f * logf((chi - clo) + 1e-5f) / 0.6931471805599453f;
        bits[i] = fmaxf(t, 0.0f);
    }
}

} // namespace ctx

int tc_watchdog(volatile int **host, int **dev);   // pmctf_kernels.cu
void count_launch();

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn tensor_map_encoder()
{   // the driver entry point is fetched through the runtime: the library does not link libcuda
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    }
    return fn;
}

static unsigned small_grid(long long items, int per_block)
{
    long long blocks = (items + per_block - 1) / per_block;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    return (unsigned)blocks;
}

template <int CC>
static int launch_pair_conv(const void *in_bf16, const void *packed_w, int taps, const float *bias, const float *res, const float *res2,
                          float lrelu_slope, float *out_f32, void *out_bf16, const float *head_w, const float *head_b, float *head_scales,
                          float *head_means, int N, int H, int W, void *stream)
{
    if (!in_bf16 || !packed_w || !bias || N <= 0 || H <= 0 || W <= 0 || !(taps == 9 || taps == 1) || (!out_f32 && !out_bf16 && !head_w))
        return PMCTF_EINVAL;
    if (head_w && (!head_b || !head_scales || !head_means)) return PMCTF_EINVAL;
    if ((((uintptr_t)in_bf16 | (uintptr_t)packed_w | (uintptr_t)out_f32 | (uintptr_t)out_bf16 | (uintptr_t)res | (uintptr_t)res2) & 15) != 0)
        return PMCTF_EINVAL;
    using G = ctx::Geo<CC>;
    if ((long long)N * G::CH > 0x7fffffffLL || W > (1 << 20) || H > (1 << 20)) return PMCTF_ESHAPE;
    static int configured_for[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
    if (dev < 0 || dev >= 64) return PMCTF_EINVAL;
    if (!configured_for[dev]) {
        cudaError_t e = cudaFuncSetAttribute(ctx::ctx_conv_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return (int)cudaGetLastError();
        configured_for[dev] = sms;
    }
    encode_tiled_fn enc = tensor_map_encoder();
    if (!enc) return PMCTF_EINVAL;
    CUtensorMap map;
    const cuuint64_t gdim[4] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * G::CH};
    const cuuint64_t gstr[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    const cuuint32_t box[4] = {8, ctx::P, ctx::IN_R, G::CH};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(in_bf16), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return PMCTF_ESHAPE;
    volatile int *herr = nullptr;
    int *derr = nullptr;
    int e = tc_watchdog(&herr, &derr);
    if (e) return e;
    if (herr[0] != 0) return PMCTF_ETIMEOUT;
    const long long tiles = (long long)((W + ctx::TW - 1) / ctx::TW) * ((H + ctx::TH - 1) / ctx::TH) * N;
    if (tiles <= 0 || tiles > 0x7fffffffLL) return PMCTF_ESHAPE;
    const long long pairs = (tiles + 1) / 2, max_clusters = configured_for[dev] / 2;
    const unsigned grid = 2u * (unsigned)(pairs < max_clusters ? pairs : max_clusters);
    ctx::ConvD d;
    d.wimg = (const uint8_t *)packed_w; d.bias = bias; d.res = res; d.res2 = res2; d.out_f32 = out_f32; d.out_bf16 = (__nv_bfloat16 *)out_bf16;
    d.head_w = head_w; d.head_b = head_b; d.head_scales = head_scales; d.head_means = head_means;
    d.slope = lrelu_slope; d.n = N; d.h = H; d.w = W; d.taps = taps;
    ctx::ctx_conv_kernel<CC><<<grid, ctx::NT, G::SMEM_BYTES, (cudaStream_t)stream>>>(map, d, derr);
    count_launch();
    return (int)cudaGetLastError();
}

// PostProcess (csrc/pmctf_pp.cu) runs its 64 -> 64 layers through the same kernel
int pair_conv64(const void *in_bf16, const void *packed_w, const float *bias, const float *res, float slope, float *out_f32, void *out_bf16, int N,
                int H, int W, void *stream)
{
    return launch_pair_conv<64>(in_bf16, packed_w, 9, bias, res, nullptr, slope, out_f32, out_bf16, nullptr, nullptr, nullptr, nullptr, N, H, W, stream);
}
int pair_pack64(const float *w, void *packed, void *stream)
{
    const int total = 2 * 9 * 4 * 2 * 32 * 8;
    ctx::ctx_pack_kernel<64><<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, 9, (__nv_bfloat16 *)packed);
    count_launch();
    return (int)cudaGetLastError();
}
} // namespace pmctf

using namespace pmctf;

extern "C" {

long long pmctf_ctx_packed_bytes(int taps) { return (taps == 9 || taps == 1) ? 2LL * taps * ctx::KS * ctx::WSLAB : 0; }

int pmctf_ctx_pack_conv(const float *w, int taps, void *packed, void *stream)
{
    if (!w || !packed || !(taps == 9 || taps == 1) || ((uintptr_t)packed & 15)) return PMCTF_EINVAL;
    const int total = 2 * taps * ctx::KS * 2 * ctx::NHALF * 8;
    ctx::ctx_pack_kernel<112><<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, taps, (__nv_bfloat16 *)packed);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_ctx_conv_in(const float *x0, const float *x1, const float *w, const float *b, float *out_f32, void *out_bf16, int N, int H, int W,
                      void *stream)
{
    if (!x0 || !w || !b || !out_f32 || !out_bf16 || N <= 0 || H <= 0 || W <= 0) return PMCTF_EINVAL;
    if ((((uintptr_t)out_f32 | (uintptr_t)out_bf16) & 15) != 0) return PMCTF_EINVAL;
    const unsigned grid = small_grid((long long)N * H * W * ctx::KS, 256);
    if (x1)
        ctx::ctx_conv_in_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(x0, x1, w, b, out_f32, (__nv_bfloat16 *)out_bf16, N, H, W);
    else
        ctx::ctx_conv_in_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x0, nullptr, w, b, out_f32, (__nv_bfloat16 *)out_bf16, N, H, W);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_ctx_conv112(const void *in_bf16, const void *packed_w, int taps, const float *bias, const float *res, const float *res2,
                      float lrelu_slope, float *out_f32, void *out_bf16, int N, int H, int W, void *stream)
{
    return launch_pair_conv<112>(in_bf16, packed_w, taps, bias, res, res2, lrelu_slope, out_f32, out_bf16, nullptr, nullptr, nullptr, nullptr, N, H, W,
                          stream);
}

int pmctf_ctx_conv112_head(const void *in_bf16, const void *packed_w, int taps, const float *bias, const float *res, const float *res2,
                           float lrelu_slope, const float *head_w, const float *head_b, float *scales, float *means, int N, int H, int W,
                           void *stream)
{
    if (!head_w) return PMCTF_EINVAL;
    return launch_pair_conv<112>(in_bf16, packed_w, taps, bias, res, res2, lrelu_slope, nullptr, nullptr, head_w, head_b, scales, means, N, H, W, stream);
}

int pmctf_ctx_lower_subband(const float *prev, const float *w, const float *b, float *out, int N, int h, int w_, void *stream)
{
    if (!prev || !w || !b || !out || N <= 0 || h <= 0 || w_ <= 0) return PMCTF_EINVAL;
    ctx::ctx_lower_kernel<<<small_grid((long long)N * h * w_ * 4, 256), 256, 0, (cudaStream_t)stream>>>(prev, w, b, out, N, h, w_);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_ctx_dcb_tail(const float *t1, const float *ctx_f32, const pmctf_ctx_dcb_t *p, float *scales, float *means, int N, int H, int W,
                       void *stream)
{
    if (!t1 || !ctx_f32 || !p || !scales || !means || N <= 0 || H <= 0 || W <= 0) return PMCTF_EINVAL;
    if (!p->dw_w || !p->dw_b || !p->pw_w || !p->pw_b || !p->ad_w || !p->ad_b || !p->f1_w || !p->f1_b || !p->f2_w || !p->f2_b) return PMCTF_EINVAL;
    ctx::DcbD d;
    d.dw_w = p->dw_w; d.dw_b = p->dw_b; d.pw_w = p->pw_w; d.pw_b = p->pw_b; d.ad_w = p->ad_w; d.ad_b = p->ad_b;
    d.f1_w = p->f1_w; d.f1_b = p->f1_b; d.f2_w = p->f2_w; d.f2_b = p->f2_b;
    const long long dtiles = (long long)((W + ctx::DT_W - 1) / ctx::DT_W) * ((H + ctx::DT_H - 1) / ctx::DT_H) * N;
    ctx::ctx_dcb_tail_kernel<<<(unsigned)(dtiles < 148 * 8 ? dtiles : 148 * 8), ctx::DT_W * ctx::DT_H, 0, (cudaStream_t)stream>>>(t1, ctx_f32, d, scales,
                                                                                                                        means, N, H, W);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_ctx_head(const float *feat, const float *w, const float *b, float *scales, float *means, int N, int H, int W, void *stream)
{
    if (!feat || !w || !b || !scales || !means || N <= 0 || H <= 0 || W <= 0) return PMCTF_EINVAL;
    ctx::ctx_head_kernel<<<small_grid((long long)N * H * W, 256), 256, 0, (cudaStream_t)stream>>>(feat, w, b, scales, means, N, H, W);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_lstm_gates(const float *a_in, const float *a_hid, const float *c, float *h_out, float *c_out, long long n, void *stream)
{
    if (n == 0) return 0;
    if (!a_in || !a_hid || !c || !h_out || !c_out || n < 0) return PMCTF_EINVAL;
    ctx::lstm_gates_kernel<<<small_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(a_in, a_hid, c, h_out, c_out, n);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_laplace_bits(const float *y, const float *sigma, float *bits, long long n, void *stream)
{
    if (n == 0) return 0;
    if (!y || !sigma || !bits || n < 0) return PMCTF_EINVAL;
    ctx::laplace_bits_kernel<<<small_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(y, sigma, bits, n);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_ctx_mask_step(const pmctf_ctx_step_t *s, void *stream)
{
    if (!s || (s->x && s->dec_sym) || !s->scales || !s->means || s->step < 0 || s->step > 3 || s->N <= 0 || s->H <= 0 || s->W <= 0)
        return PMCTF_EINVAL;
    const bool index_only = !s->x && !s->dec_sym;   // decoder, before the step's symbols exist: the table indexes alone
    if (index_only ? (!s->idx16 || s->sym16) : (!s->x_hat || !s->s_hat)) return PMCTF_EINVAL;
    if (s->sym16 && !s->idx16) return PMCTF_EINVAL;
    if (s->idx16 && (!(s->log_scale_step > 0.0f) || s->scale_levels < 1 || s->scale_levels > 32767)) return PMCTF_EINVAL;
    ctx::StepD d;
    d.x = s->x; d.dec_sym = s->dec_sym; d.scales = s->scales; d.means = s->means; d.x_hat = s->x_hat; d.x_q = s->x_q; d.s_hat = s->s_hat;
    d.x_res = s->x_res; d.sym16 = s->sym16; d.idx16 = s->idx16; d.log_min = s->log_scale_min; d.log_step = s->log_scale_step;
    d.top = (float)(s->scale_levels - 1); d.step = s->step; d.lossy = s->lossy; d.first = s->step == 0; d.N = s->N; d.H = s->H; d.W = s->W;
    ctx::ctx_mask_step_kernel<<<small_grid((long long)s->N * s->H * s->W, 256), 256, 0, (cudaStream_t)stream>>>(d);
    count_launch();
    return (int)cudaGetLastError();
}

} // extern "C"
