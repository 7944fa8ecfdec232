// pmctf_kernels.cu -- sm_100a kernels for the MCTF + pWave++ lifting hot path.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false ...
// (-fmad=false is part of the arithmetic contract: fma happens only where fmaf() is written.)
//
// Kernels
//   lift_step_kernel<SRC>   fused {plane | flow-warp | 3-tap skip} -> PredictUpdate CNN (4 conv
//                           layers chained in shared memory, halo recompute) -> lifting
//                           accumulate.  One launch == one lifting step == one HBM pass.
//   flow_warp_kernel        stand-alone bilinear backward warp (flow_warp of the reference)
//   chroma_mv_down_kernel   2x2 mean / 2 of the luma motion field
//   quantize / dequantize   per-subband elementwise
//   pack_pu_kernel          OIHW -> kernel weight layout
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "pmctf_b200.h"
#include "pmctf_common.cuh"

namespace pmctf {

// number of kernels this library has launched (process-wide; bench.py reports the delta over its timed region)
static std::atomic<unsigned long long> g_launches{0};
#define PMCTF_LAUNCHED() (pmctf::g_launches.fetch_add(1, std::memory_order_relaxed), (int)cudaGetLastError())

// ------------------------------------------------------------------------------------------
// tile geometry (logical coordinates: what the reference's conv2d sees)
constexpr int TH = 32, TW = 32;
constexpr int S_ROWS = TH + 8, S_COLS = TW + 8, S_P = 41;    // source tile s, origin (-4,-4)
constexpr int T_ROWS = TH + 10, T_P = 41;                    // raw src tile for SKIP3, origin (-5,-4)
constexpr int A1_ROWS = TH + 6, A1_COLS = TW + 6, A1_P = 40; // tanh(conv1), origin (-3,-3)
constexpr int A2_ROWS = TH + 4, A2_COLS = TW + 4, A2_P = 40; // tanh(conv2), origin (-2,-2)
constexpr int A3_ROWS = TH + 2, A3_COLS = TW + 2, A3_P = 36; // conv1 + conv3, origin (-1,-1)
constexpr int O_P = 33;                                      // PU output tile

// packed PredictUpdate weights (floats)
constexpr int W1_OFF = 0;      // [9][16]          (k, co)
constexpr int B1_OFF = 144;    // [16]
constexpr int W2_OFF = 160;    // [16 ci][9][16 co]
constexpr int B2_OFF = 2464;
constexpr int W3_OFF = 2480;
constexpr int B3_OFF = 4784;
constexpr int W4_OFF = 4800;   // [16 ci][12] (9 used)
constexpr int B4_OFF = 4992;
constexpr int WPACK = 5000;        // fp32 part of the packed block
constexpr int Q_OFF = 5000;        // int8 tensor-core operand images of conv2 / conv3 (2 x 10240 bytes)
constexpr int QBYTES = 10240;      // 5 tap pairs x 1536 B, then the 80-row (zero padded) image of tap pair 0 at +7680
constexpr int SC_OFF = 10120;      // 2^-(22+Sw) of conv2, conv3
static_assert(PMCTF_PU_PACKED_FLOATS == 10128 && SC_OFF + 8 == PMCTF_PU_PACKED_FLOATS, "header and kernel disagree on the packed size");

// shared memory carve-up (floats)
constexpr int SM_W = 0;
constexpr int SM_S = SM_W + WPACK;                 // 5000
constexpr int SM_T = SM_S + S_ROWS * S_P;          // + 1640
constexpr int SM_A1 = SM_T + 1724;                 // T tile 42*41 = 1722, padded for 16 B alignment
constexpr int SM_A2 = SM_A1 + 16 * A1_ROWS * A1_P; // + 24320
constexpr int SM_END = SM_A2 + 16 * A2_ROWS * A2_P;
constexpr int SM_TANH_B = SM_END * 4;              // byte offset of the tanh table
constexpr int SMEM_BYTES = SM_TANH_B + TANH_SMEM_BYTES;
static_assert(SM_TANH_B % 16 == 0 && (SM_A1 * 4) % 16 == 0 && (SM_A2 * 4) % 16 == 0, "activation planes must be 16 B aligned");
static_assert(16 * A3_ROWS * A3_P <= 16 * A1_ROWS * A1_P, "a3 aliases a1");
static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit the 227 KB per-CTA limit");

constexpr int NT = 672; // 21 warps, <= 96 registers each, one CTA per SM

// ------------------------------------------------------------------------------------------
// 16 -> 16 channel 3x3 layer on shared-memory planes.  Each work item is CO_T output channels
// x 4 consecutive pixels; every accumulator is one sequential fma chain in (ci, ky, kx) order.
template <int CO_T, int IN_P, int OUT_ROWS, int OUT_STRIPS, typename Epi>
__device__ __forceinline__ void conv16_layer(const float *__restrict__ in, const int in_plane,
                                             const float *__restrict__ wp, const float *__restrict__ bias, Epi epi)
{
    constexpr int NCG = 16 / CO_T;
    constexpr int PER_CG = OUT_ROWS * OUT_STRIPS;
    constexpr int ITEMS = NCG * PER_CG;
    for (int it = threadIdx.x; it < ITEMS; it += NT) {
        const int cg = it / PER_CG;
        const int rem = it - cg * PER_CG;
        const int r = rem / OUT_STRIPS;
        const int st = rem - r * OUT_STRIPS;
        // accumulators as (co, co+1) pairs: one packed fp32x2 FMA (FFMA2) updates two output channels with the same input
        // value -- bit-identical to two fmaf() chains, half the issue slots
        float2 acc2[CO_T / 2][4];
#pragma unroll
        for (int c2 = 0; c2 < CO_T / 2; ++c2) {
            const float2 b = make_float2(bias[cg * CO_T + 2 * c2], bias[cg * CO_T + 2 * c2 + 1]);
#pragma unroll
            for (int p = 0; p < 4; ++p) acc2[c2][p] = b;
        }
        const float *ip = in + r * IN_P + st * 4;
        const float *wc = wp + cg * CO_T;
#pragma unroll 1
        for (int ci = 0; ci < 16; ++ci) {
            float patch[3][6];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float4 a = *reinterpret_cast<const float4 *>(ip + ky * IN_P);
                const float2 b = *reinterpret_cast<const float2 *>(ip + ky * IN_P + 4);
                patch[ky][0] = a.x; patch[ky][1] = a.y; patch[ky][2] = a.z; patch[ky][3] = a.w;
                patch[ky][4] = b.x; patch[ky][5] = b.y;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                float2 wv[CO_T / 2];
#pragma unroll
                for (int c4 = 0; c4 < CO_T / 4; ++c4) {
                    const float4 t = *reinterpret_cast<const float4 *>(wc + k * 16 + c4 * 4);
                    wv[c4 * 2] = make_float2(t.x, t.y);
                    wv[c4 * 2 + 1] = make_float2(t.z, t.w);
                }
                const int ky = k / 3, kx = k - ky * 3;
#pragma unroll
                for (int c2 = 0; c2 < CO_T / 2; ++c2)
#pragma unroll
                    for (int p = 0; p < 4; ++p) acc2[c2][p] = ffma2(wv[c2], patch[ky][kx + p], acc2[c2][p]);
            }
            ip += in_plane;
            wc += 144;
        }
        float acc[CO_T][4];
#pragma unroll
        for (int c2 = 0; c2 < CO_T / 2; ++c2)
#pragma unroll
            for (int p = 0; p < 4; ++p) { acc[2 * c2][p] = acc2[c2][p].x; acc[2 * c2 + 1][p] = acc2[c2][p].y; }
        epi(cg, r, st, acc);
    }
}

template <int SRC>
__global__ void __launch_bounds__(NT, 1) lift_step_kernel(const __grid_constant__ StepD a)
{
    extern __shared__ __align__(16) float smem[];
    float *sw = smem + SM_W;
    float *ss = smem + SM_S;
    float *stile = smem + SM_T;
    float *a1 = smem + SM_A1;
    float *a2 = smem + SM_A2;
    float *a3 = a1;   // a1 is dead once conv2 has been computed
    float *so = a2;   // a2 is dead once conv3 has been computed
    const float *ttab = reinterpret_cast<const float *>(reinterpret_cast<const unsigned char *>(smem) + SM_TANH_B);

    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const int H = a.h, W = a.w;

    // ---- phase A: weights + source tile ------------------------------------------------------
    {
        const float4 *g = reinterpret_cast<const float4 *>(a.pu_packed);
        float4 *d = reinterpret_cast<float4 *>(sw);
        for (int i = tid; i < WPACK / 4; i += NT) d[i] = __ldg(g + i);
        load_tanh_table(reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(smem) + SM_TANH_B), tid, NT);
    }
    const bool xfast_src = a.src.cs <= a.src.rs;
    if (SRC == PMCTF_SRC_PLANE || SRC == PMCTF_SRC_SKIP3) {
        // raw source tile (with the reference's divisions applied), zero outside the image;
        // threads run along the axis with the smaller stride so global loads coalesce.
        constexpr int ROWS = (SRC == PMCTF_SRC_SKIP3) ? T_ROWS : S_ROWS;
        constexpr int ROFF = (SRC == PMCTF_SRC_SKIP3) ? 5 : 4;
        float *dst = (SRC == PMCTF_SRC_SKIP3) ? stile : ss;
        const float *sp = a.src.p + plane_off(a.src, n);
        const bool dodiv = (a.src_div1 != 1.0f) || (a.src_div2 != 1.0f);
        for (int i = tid; i < ROWS * S_COLS; i += NT) {
            int r, c;
            if (xfast_src) { r = i / S_COLS; c = i - r * S_COLS; }
            else { c = i / ROWS; r = i - c * ROWS; }
            const int gy = y0 - ROFF + r, gx = x0 - 4 + c;
            float v = 0.0f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                v = __ldg(sp + (long long)gy * a.src.rs + (long long)gx * a.src.cs);
                if (dodiv) v = (v / a.src_div1) / a.src_div2;
            }
            dst[r * S_P + c] = v;
        }
    } else {
        // s = warp(src, sign * mv): pMCTF_L.py:301,307 -> video_net.py:32-55
        const float *sp = a.src.p + plane_off(a.src, n);
        for (int i = tid; i < S_ROWS * S_COLS; i += NT) {
            const int r = i / S_COLS, c = i - r * S_COLS;
            const int gy = y0 - 4 + r, gx = x0 - 4 + c;
            float v = 0.0f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                float fx, fy;
                load_mv(a.mv, a.mv_share, a.mv_down, a.mv_h, a.mv_w, n, gy, gx, a.mv_sign, fx, fy);
                v = warp_sample(sp, a.src.rs, a.src.cs, H, W, __ldg(a.lin_x + gx), __ldg(a.lin_y + gy), fx, fy, a.sx, a.sy);
                if (a.round_src) v = rintf(v);
            }
            ss[r * S_P + c] = v;
        }
    }
    __syncthreads();
    if (SRC == PMCTF_SRC_SKIP3) {
        // skip = conv(3,1)(ReflectionPad2d((0,0,1,1))(src)) + bias: lifting_1d.py:91,105-106
        for (int i = tid; i < S_ROWS * S_COLS; i += NT) {
            const int r = i / S_COLS, c = i - r * S_COLS;
            const int gy = y0 - 4 + r, gx = x0 - 4 + c;
            float v = 0.0f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const int ym = (gy == 0) ? 1 : gy - 1;
                const int yp = (gy == H - 1) ? H - 2 : gy + 1;
                const int base = y0 - 5;
                v = a.tap_bias;
                v = fmaf(a.tap0, stile[(ym - base) * T_P + c], v);
                v = fmaf(a.tap1, stile[(gy - base) * T_P + c], v);
                v = fmaf(a.tap2, stile[(yp - base) * T_P + c], v);
            }
            ss[r * S_P + c] = v;
        }
        __syncthreads();
    }

    // ---- phase C: conv1 (1 -> 16) + tanh -> a1 -------------------------------------------------
    {
        const float in_mul = a.in_mul;
        constexpr int PER = A1_ROWS * A1_COLS;
        for (int it = tid; it < 4 * PER; it += NT) {
            const int cq = it / PER;
            const int rem = it - cq * PER;
            const int r = rem / A1_COLS, c = rem - r * A1_COLS;
            const int gy = y0 - 3 + r, gx = x0 - 3 + c;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const float4 b = *reinterpret_cast<const float4 *>(sw + B1_OFF + cq * 4);
                float acc0 = b.x, acc1 = b.y, acc2 = b.z, acc3 = b.w;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const int ky = k / 3, kx = k - ky * 3;
                    const float v = ss[(r + ky) * S_P + c + kx] * in_mul;
                    const float4 wv = *reinterpret_cast<const float4 *>(sw + W1_OFF + k * 16 + cq * 4);
                    acc0 = fmaf(wv.x, v, acc0); acc1 = fmaf(wv.y, v, acc1);
                    acc2 = fmaf(wv.z, v, acc2); acc3 = fmaf(wv.w, v, acc3);
                }
                o = make_float4(tanh_det(acc0, ttab), tanh_det(acc1, ttab), tanh_det(acc2, ttab), tanh_det(acc3, ttab));
            }
            float *d = a1 + (cq * 4) * (A1_ROWS * A1_P) + r * A1_P + c;
            d[0] = o.x; d[A1_ROWS * A1_P] = o.y; d[2 * A1_ROWS * A1_P] = o.z; d[3 * A1_ROWS * A1_P] = o.w;
        }
    }
    __syncthreads();

    // ---- phase D: conv2 (16 -> 16) + tanh -> a2 -------------------------------------------------
    conv16_layer<8, A1_P, A2_ROWS, A2_COLS / 4>(a1, A1_ROWS * A1_P, sw + W2_OFF, sw + B2_OFF,
        [&](int cg, int r, int st, float (&acc)[8][4]) {
            const int gy = y0 - 2 + r;
            const bool rowin = gy >= 0 && gy < H;
#pragma unroll
            for (int co = 0; co < 8; ++co) {
                float v[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int gx = x0 - 2 + st * 4 + p;
                    v[p] = (rowin && gx >= 0 && gx < W) ? tanh_det(acc[co][p], ttab) : 0.0f;
                }
                *reinterpret_cast<float4 *>(a2 + (cg * 8 + co) * (A2_ROWS * A2_P) + r * A2_P + st * 4) =
                    make_float4(v[0], v[1], v[2], v[3]);
            }
        });
    __syncthreads();

    // ---- phase E: conv3 (16 -> 16) + conv1 residual -> a3 ---------------------------------------
    conv16_layer<8, A2_P, A3_ROWS, (A3_COLS + 3) / 4>(a2, A2_ROWS * A2_P, sw + W3_OFF, sw + B3_OFF,
        [&](int cg, int r, int st, float (&acc)[8][4]) {
            const int gy = y0 - 1 + r;
            const bool rowin = gy >= 0 && gy < H;
            const float in_mul = a.in_mul;
            float v[8][4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int c = st * 4 + p;
                const int gx = x0 - 1 + c;
                const bool in = rowin && gx >= 0 && gx < W && c < A3_COLS;
                // conv1 at this position, recomputed with the identical fma chain (lifting_1d.py:45)
                float c1[8];
#pragma unroll
                for (int co = 0; co < 8; ++co) c1[co] = sw[B1_OFF + cg * 8 + co];
                if (in) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const int ky = k / 3, kx = k - ky * 3;
                        const float sv = ss[(r + 2 + ky) * S_P + c + 2 + kx] * in_mul;
#pragma unroll
                        for (int co = 0; co < 8; ++co) c1[co] = fmaf(sw[W1_OFF + k * 16 + cg * 8 + co], sv, c1[co]);
                    }
                }
#pragma unroll
                for (int co = 0; co < 8; ++co) v[co][p] = in ? (c1[co] + acc[co][p]) : 0.0f;
            }
#pragma unroll
            for (int co = 0; co < 8; ++co)
                *reinterpret_cast<float4 *>(a3 + (cg * 8 + co) * (A3_ROWS * A3_P) + r * A3_P + st * 4) =
                    make_float4(v[co][0], v[co][1], v[co][2], v[co][3]);
        });
    __syncthreads();

    // ---- phase F: conv4 (16 -> 1) -> so ----------------------------------------------------------
    for (int it = tid; it < TH * (TW / 2); it += NT) {
        const int r = it / (TW / 2), st = it - r * (TW / 2);
        float acc0 = sw[B4_OFF], acc1 = acc0;
        const float *ip = a3 + r * A3_P + st * 2;
        const float *wc = sw + W4_OFF;
#pragma unroll 4
        for (int ci = 0; ci < 16; ++ci) {
            const float4 w0 = *reinterpret_cast<const float4 *>(wc);
            const float4 w1 = *reinterpret_cast<const float4 *>(wc + 4);
            const float w8 = wc[8];
            const float wk[9] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w8};
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float2 p0 = *reinterpret_cast<const float2 *>(ip + ky * A3_P);
                const float2 p1 = *reinterpret_cast<const float2 *>(ip + ky * A3_P + 2);
                const float pv[4] = {p0.x, p0.y, p1.x, p1.y};
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    acc0 = fmaf(wk[ky * 3 + kx], pv[kx], acc0);
                    acc1 = fmaf(wk[ky * 3 + kx], pv[kx + 1], acc1);
                }
            }
            ip += A3_ROWS * A3_P;
            wc += 12;
        }
        so[r * O_P + st * 2] = acc0;
        so[r * O_P + st * 2 + 1] = acc1;
    }
    __syncthreads();

    // ---- phase G: lifting arithmetic + stores, threads along the output's contiguous axis ---------
    {
        const bool xfast = a.out.cs <= a.out.rs;
        const long long o_off = plane_off(a.out, n);
        const long long b_off = a.base.p ? plane_off(a.base, n) : 0;
        const long long p_off = a.pred.p ? plane_off(a.pred, n) : 0;
        const long long x_off = a.aux.p ? plane_off(a.aux, n) : 0;
        const float bd1 = (n >= a.div_group_n) ? a.base_div1_g1 : a.base_div1;
        const bool bdiv = (bd1 != 1.0f) || (a.base_div2 != 1.0f);
        for (int i = tid; i < TH * TW; i += NT) {
            int r, c;
            if (xfast) { r = i / TW; c = i - r * TW; }
            else { c = i / TH; r = i - c * TH; }
            const int gy = y0 + r, gx = x0 + c;
            if (gy >= H || gx >= W) continue;
            const float pu = so[r * O_P + c];
            float res;
            if (a.mode == PMCTF_MODE_PU) {
                res = pu;
            } else {
                const float s = ss[(r + 4) * S_P + c + 4];
                const float t = pu * a.post_mul;
                float tmp = s + t * 0.1f;
                if (a.round_tmp) tmp = rintf(tmp);
                const float rr = tmp * a.out_mul;
                if (a.pred.p) a.pred.p[p_off + (long long)gy * a.pred.rs + (long long)gx * a.pred.cs] = rr;
                if (a.mode == PMCTF_MODE_FILTER) {
                    res = rr;
                } else {
                    float b = __ldg(a.base.p + b_off + (long long)gy * a.base.rs + (long long)gx * a.base.cs);
                    if (bdiv) b = (b / bd1) / a.base_div2;
                    res = (a.sign > 0.0f) ? b + rr : b - rr;
                    res = res * a.final_mul;
                }
            }
            a.out.p[o_off + (long long)gy * a.out.rs + (long long)gx * a.out.cs] = res;
            if (a.aux.p) {
                const float raw = (SRC == PMCTF_SRC_SKIP3) ? stile[(r + 5) * T_P + c + 4] : ss[(r + 4) * S_P + c + 4];
                a.aux.p[x_off + (long long)gy * a.aux.rs + (long long)gx * a.aux.cs] = raw * a.aux_mul;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// flow_warp (video_net.py:32-55): one thread per output pixel, all C channels (any W / alignment)
__global__ void __launch_bounds__(256) flow_warp_kernel(const float *__restrict__ im, const float *__restrict__ flow,
                                                        const float *__restrict__ lin_x, const float *__restrict__ lin_y,
                                                        float *__restrict__ out, int N, int C, int H, int W, int flowN,
                                                        float sign, float sx, float sy, int round_out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z;
    if (x >= W) return;
    float fx, fy;
    load_mv(flow, N / flowN, 0, H, W, n, y, x, sign, fx, fy);
    const float lx = __ldg(lin_x + x), ly = __ldg(lin_y + y);
    for (int c = 0; c < C; ++c) {
        const float *p = im + ((long long)n * C + c) * H * W;
        float v = warp_sample(p, W, 1, H, W, lx, ly, fx, fy, sx, sy);
        if (round_out) v = rintf(v);
        out[((long long)n * C + c) * H * W + (long long)y * W + x] = v;
    }
}

// the same with four pixels per thread, one in each of four consecutive rows: every load and store of a warp stays coalesced
// (consecutive lanes = consecutive x) while the 8 motion-vector loads and then the 16 gathers of a thread are in flight together
// (the one-pixel form is bound by memory latency).  The kernel is bound by its instruction stream, not by HBM, so everything
// that is not the reference's coordinate arithmetic is kept off it: 32-bit offsets inside a plane (H * W < 2^31, checked by the
// launcher), one 64-bit base per plane, the sample position and the four weights computed once per pixel for all C channels,
// the plane of the motion field resolved without an integer division when every image has its own field.
#ifndef PMCTF_WARP_ROWS
#define PMCTF_WARP_ROWS 4
#endif
#ifndef PMCTF_WARP_MINB
#define PMCTF_WARP_MINB 12   // 40 registers: 48 warps per SM (measured: 8 ... 12 blocks equal within 1 %, 16 blocks spill and lose 6 %)
#endif
constexpr int WR = PMCTF_WARP_ROWS;   // rows per thread
__global__ void __launch_bounds__(128, PMCTF_WARP_MINB) flow_warp4_kernel(const float *__restrict__ im, const float *__restrict__ flow,
                                                         const float *__restrict__ lin_x, const float *__restrict__ lin_y,
                                                         float *__restrict__ out, int N, int C, int H, int W, int share,
                                                         float sign, float sx, float sy, int round_out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * WR;
    const int n = blockIdx.z;
    if (x >= W) return;
    const int plane = H * W;
    const int f = (share == 1) ? n : n / share;      // `share` consecutive images use one field
    const float *fb = flow + (size_t)f * 2 * (size_t)plane;
    const int o0 = y0 * W + x;
    const float lx = __ldg(lin_x + x);
    const float xmax = (float)(W - 1), ymax = (float)(H - 1);
    float fx[WR], fy[WR], ly[WR];
#pragma unroll
    for (int j = 0; j < WR; ++j) {
        fx[j] = fy[j] = ly[j] = 0.0f;
        if (y0 + j < H) {
            fx[j] = sign * __ldg(fb + o0 + j * W);
            fy[j] = sign * __ldg(fb + plane + o0 + j * W);
            ly[j] = __ldg(lin_y + y0 + j);
        }
    }
    // video_net.py:42-50 + ATen grid_sampler_2d (align_corners=True, border), op for op as warp_sample()
    int idx[WR];
    float nw[WR], ne[WR], sw[WR], se[WR];
    bool x1ok[WR], y1ok[WR];
#pragma unroll
    for (int j = 0; j < WR; ++j) {
        const float gx = lx + fx[j] / sx;
        const float gy = ly[j] + fy[j] / sy;
        float ix = (gx + 1.0f) * sx;
        float iy = (gy + 1.0f) * sy;
        ix = fminf(fmaxf(ix, 0.0f), xmax);
        iy = fminf(fmaxf(iy, 0.0f), ymax);
        const float xf = floorf(ix), yf = floorf(iy);
        const float w = ix - xf, e = 1.0f - w, nn = iy - yf, so = 1.0f - nn;
        nw[j] = so * e; ne[j] = so * w; sw[j] = nn * e; se[j] = nn * w;
        const int x0i = (int)xf, y0i = (int)yf;
        x1ok[j] = x0i + 1 <= W - 1;
        y1ok[j] = y0i + 1 <= H - 1;
        idx[j] = y0i * W + x0i;
    }
    const bool rnd = round_out != 0;
    for (int c = 0; c < C; ++c) {
        const size_t pb = (size_t)(n * C + c) * (size_t)plane;
        const float *ip = im + pb;
        float v[WR];
#pragma unroll
        for (int j = 0; j < WR; ++j) {
            v[j] = 0.0f;
            if (y0 + j < H) {
                const float *p = ip + idx[j];
                const float *q = p + W;
                const float vnw = __ldg(p);
                const float vne = x1ok[j] ? __ldg(p + 1) : 0.0f;
                const float vsw = y1ok[j] ? __ldg(q) : 0.0f;
                const float vse = (x1ok[j] && y1ok[j]) ? __ldg(q + 1) : 0.0f;
                float acc = vnw * nw[j];
                acc = fmaf(vne, ne[j], acc);
                acc = fmaf(vsw, sw[j], acc);
                acc = fmaf(vse, se[j], acc);
                v[j] = rnd ? rintf(acc) : acc;
            }
        }
        float *op = out + pb + o0;
#pragma unroll
        for (int j = 0; j < WR; ++j)
            if (y0 + j < H) op[j * W] = v[j];
    }
}

__global__ void __launch_bounds__(256) chroma_mv_down_kernel(const float *__restrict__ mv, float *__restrict__ out,
                                                             int planes, int H, int W)
{
    const int w2 = W / 2, h2 = H / 2;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, p = blockIdx.z;
    if (x >= w2 || p >= planes) return;
    const float *a = mv + (long long)p * H * W + (long long)(2 * y) * W + 2 * x;
    const float2 r0 = __ldg(reinterpret_cast<const float2 *>(a));
    const float2 r1 = __ldg(reinterpret_cast<const float2 *>(a + W));
    const float v = ((r0.x * 0.25f + r0.y * 0.25f) + r1.x * 0.25f) + r1.y * 0.25f;
    out[(long long)p * h2 * w2 + (long long)y * w2 + x] = v / 2.0f;
}

__device__ __forceinline__ float quant1(float v, float q, float clip, int lossy, int do_round)
{
    v = lossy ? v * q : v;
    v = fminf(fmaxf(v, -clip), clip);
    return do_round ? rintf(v) : v;
}

// the element-wise kernels take a 128-bit path when the pointers are 16-byte aligned (vec = 1): n4 = n / 4 quads, then the tail
__global__ void __launch_bounds__(256) quantize_kernel(const float *__restrict__ s, float q, float clip, int lossy,
                                                       int do_round, float *__restrict__ out, long long n, int vec)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec ? n >> 2 : 0;
    for (long long i = t; i < n4; i += nt) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(s) + i);
        reinterpret_cast<float4 *>(out)[i] = make_float4(quant1(v.x, q, clip, lossy, do_round), quant1(v.y, q, clip, lossy, do_round),
                                                         quant1(v.z, q, clip, lossy, do_round), quant1(v.w, q, clip, lossy, do_round));
    }
    for (long long i = 4 * n4 + t; i < n; i += nt) out[i] = quant1(s[i], q, clip, lossy, do_round);
}

__global__ void __launch_bounds__(256) dequantize_kernel(const float *__restrict__ s, float q, int lossy,
                                                         float *__restrict__ out, long long n, int vec)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec ? n >> 2 : 0;
    for (long long i = t; i < n4; i += nt) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(s) + i);
        reinterpret_cast<float4 *>(out)[i] = lossy ? make_float4(v.x / q, v.y / q, v.z / q, v.w / q) : v;
    }
    for (long long i = 4 * n4 + t; i < n; i += nt) out[i] = lossy ? s[i] / q : s[i];
}

// quantise with per-plane integer statistics of the symbols (sum |sym|, #nonzero): exact u64
// arithmetic, so the rate statistics gathered across GPUs do not depend on reduction order.
__global__ void __launch_bounds__(256) quantize_stats_kernel(const float *__restrict__ s, float q, float clip, int lossy,
                                                             float *__restrict__ out, long long plane_elems,
                                                             unsigned long long *__restrict__ stats, int vec)
{
    const int plane = blockIdx.y;
    const float *sp = s + (long long)plane * plane_elems;
    float *op = out + (long long)plane * plane_elems;
    unsigned long long sum = 0;
    unsigned int nnz = 0;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec ? plane_elems >> 2 : 0;
    for (long long i = t; i < n4; i += nt) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(sp) + i);
        const float4 r = make_float4(quant1(v.x, q, clip, lossy, 1), quant1(v.y, q, clip, lossy, 1), quant1(v.z, q, clip, lossy, 1),
                                     quant1(v.w, q, clip, lossy, 1));
        reinterpret_cast<float4 *>(op)[i] = r;
        const unsigned int a0 = (unsigned int)fabsf(r.x), a1 = (unsigned int)fabsf(r.y), a2 = (unsigned int)fabsf(r.z), a3 = (unsigned int)fabsf(r.w);
        sum += a0 + a1 + a2 + a3;
        nnz += (a0 != 0) + (a1 != 0) + (a2 != 0) + (a3 != 0);
    }
    for (long long i = 4 * n4 + t; i < plane_elems; i += nt) {
        const float v = quant1(sp[i], q, clip, lossy, 1);
        op[i] = v;
        const unsigned int iv = (unsigned int)fabsf(v);
        sum += iv;
        nnz += iv != 0;
    }
    unsigned long long s64 = sum, n64 = nnz;
    for (int o = 16; o > 0; o >>= 1) {
        s64 += __shfl_down_sync(0xffffffffu, s64, o);
        n64 += __shfl_down_sync(0xffffffffu, n64, o);
    }
    __shared__ unsigned long long red[2][8];
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s64; red[1][threadIdx.x >> 5] = n64; }
    __syncthreads();
    if (threadIdx.x == 0) {   // one pair of atomics per block
        for (int w = 1; w < 8; ++w) { s64 += red[0][w]; n64 += red[1][w]; }
        if (s64 | n64) {
            atomicAdd(stats + 2 * plane, s64);
            atomicAdd(stats + 2 * plane + 1, n64);
        }
    }
}

// The coder's per-band step for a batch of planes that do NOT share a quantisation step (the H frames of all temporal levels of
// a GOP coded as one batch: hp_q_scale differs per level, pMCTF_L.py:343-347): sym = rint(clamp(s * q[p], +-clip)); the fp32
// output is the symbol or, with `dequant`, already sym / q[p] (dequantize_subbands, pWave.py:191-202 -- the same single
// division the fused inverse-lifting loads would do); optional int16 copy of the symbols at sym16[p * stride + i] (what the
// reference hands to the entropy coder, entropy_models.py:37-40) and the exact per-plane statistics.
__global__ void __launch_bounds__(256) quantize_code_kernel(const float *__restrict__ s, const float *__restrict__ q_tab, float clip,
                                                            int lossy, int dequant, float *__restrict__ out, short *__restrict__ sym16,
                                                            long long sym16_stride, long long plane_elems,
                                                            unsigned long long *__restrict__ stats, int vec)
{
    const int plane = blockIdx.y;
    const float q = __ldg(q_tab + plane);
    const float *sp = s + (long long)plane * plane_elems;
    float *op = out + (long long)plane * plane_elems;
    short *yp = sym16 ? sym16 + (long long)plane * sym16_stride : nullptr;
    const bool div = dequant && lossy;
    unsigned long long sum = 0;
    unsigned int nnz = 0;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec ? plane_elems >> 2 : 0;
    for (long long i = t; i < n4; i += nt) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(sp) + i);
        const float4 r = make_float4(quant1(v.x, q, clip, lossy, 1), quant1(v.y, q, clip, lossy, 1), quant1(v.z, q, clip, lossy, 1),
                                     quant1(v.w, q, clip, lossy, 1));
        reinterpret_cast<float4 *>(op)[i] = div ? make_float4(r.x / q, r.y / q, r.z / q, r.w / q) : r;
        if (yp) reinterpret_cast<short4 *>(yp)[i] = make_short4((short)(int)r.x, (short)(int)r.y, (short)(int)r.z, (short)(int)r.w);
        const unsigned int a0 = (unsigned int)fabsf(r.x), a1 = (unsigned int)fabsf(r.y), a2 = (unsigned int)fabsf(r.z), a3 = (unsigned int)fabsf(r.w);
        sum += a0 + a1 + a2 + a3;
        nnz += (a0 != 0) + (a1 != 0) + (a2 != 0) + (a3 != 0);
    }
    for (long long i = 4 * n4 + t; i < plane_elems; i += nt) {
        const float v = quant1(sp[i], q, clip, lossy, 1);
        op[i] = div ? v / q : v;
        if (yp) yp[i] = (short)(int)v;
        const unsigned int iv = (unsigned int)fabsf(v);
        sum += iv;
        nnz += iv != 0;
    }
    unsigned long long s64 = sum, n64 = nnz;
    for (int o = 16; o > 0; o >>= 1) {
        s64 += __shfl_down_sync(0xffffffffu, s64, o);
        n64 += __shfl_down_sync(0xffffffffu, n64, o);
    }
    __shared__ unsigned long long red[2][8];
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s64; red[1][threadIdx.x >> 5] = n64; }
    __syncthreads();
    if (threadIdx.x == 0 && stats) {   // one pair of atomics per block
        for (int w = 1; w < 8; ++w) { s64 += red[0][w]; n64 += red[1][w]; }
        if (s64 | n64) {
            atomicAdd(stats + 2 * plane, s64);
            atomicAdd(stats + 2 * plane + 1, n64);
        }
    }
}

// 8-bit samples -> fp32 planes, zero-padded bottom/right (np_image_to_tensor + F.pad, test_pMCTF_flex.py:151-192)
constexpr int UNPACK_ROWS = 4;   // rows per thread: four loads in flight before the four 16-byte stores, a quarter of the blocks
__global__ void __launch_bounds__(256) unpack_u8_kernel(const unsigned char *__restrict__ src, float *__restrict__ dst,
                                                        int h0, int w0, int hp, int wp)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y0 = blockIdx.y * UNPACK_ROWS;
    const long long n = blockIdx.z;
    if (x4 >= wp) return;
    const bool vec = x4 + 3 < w0 && (w0 & 3) == 0;
    float4 v[UNPACK_ROWS];
#pragma unroll
    for (int j = 0; j < UNPACK_ROWS; ++j) {
        const int y = y0 + j;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y < h0) {
            const unsigned char *sp = src + (n * h0 + y) * (long long)w0;
            if (vec) {
                const uchar4 u = __ldg(reinterpret_cast<const uchar4 *>(sp + x4));
                v[j] = make_float4(u.x, u.y, u.z, u.w);
            } else {
                if (x4 + 0 < w0) v[j].x = sp[x4 + 0];
                if (x4 + 1 < w0) v[j].y = sp[x4 + 1];
                if (x4 + 2 < w0) v[j].z = sp[x4 + 2];
                if (x4 + 3 < w0) v[j].w = sp[x4 + 3];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < UNPACK_ROWS; ++j)
        if (y0 + j < hp) *reinterpret_cast<float4 *>(dst + (n * hp + y0 + j) * (long long)wp + x4) = v[j];
}
__global__ void __launch_bounds__(256) frame_sse_kernel(const float *__restrict__ rec, const unsigned char *__restrict__ orig,
                                                        int h0, int w0, int hp, int wp, unsigned long long *__restrict__ sse, int vec)
{
    // a block walks rows blockIdx.x, blockIdx.x + gridDim.x, ... of plane blockIdx.y; one atomic per block
    const long long n = blockIdx.y;
    unsigned long long acc = 0;
    for (int y = blockIdx.x; y < h0; y += gridDim.x) {
        const float *rp = rec + (n * hp + y) * (long long)wp;
        const unsigned char *op = orig + (n * h0 + y) * (long long)w0;
        unsigned int row = 0;   // <= 255^2 * (w0 / threads + 4) per thread
        const int w4 = vec ? w0 >> 2 : 0;
        auto sq4 = [](const float4 r, const uchar4 o) -> unsigned int {
            const int d0 = (int)rintf(fminf(fmaxf(r.x, 0.0f), 255.0f)) - (int)o.x, d1 = (int)rintf(fminf(fmaxf(r.y, 0.0f), 255.0f)) - (int)o.y;
            const int d2 = (int)rintf(fminf(fmaxf(r.z, 0.0f), 255.0f)) - (int)o.z, d3 = (int)rintf(fminf(fmaxf(r.w, 0.0f), 255.0f)) - (int)o.w;
            return (unsigned int)(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
        };
        int i = threadIdx.x;
        for (; i + (int)blockDim.x < w4; i += 2 * blockDim.x) {   // two independent pairs of loads in flight (integer sums: any order)
            const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rp) + i), r1 = __ldg(reinterpret_cast<const float4 *>(rp) + i + blockDim.x);
            const uchar4 o0 = __ldg(reinterpret_cast<const uchar4 *>(op) + i), o1 = __ldg(reinterpret_cast<const uchar4 *>(op) + i + blockDim.x);
            row += sq4(r0, o0) + sq4(r1, o1);
        }
        for (; i < w4; i += blockDim.x)
            row += sq4(__ldg(reinterpret_cast<const float4 *>(rp) + i), __ldg(reinterpret_cast<const uchar4 *>(op) + i));
        for (int x = 4 * w4 + threadIdx.x; x < w0; x += blockDim.x) {
            const int d = (int)rintf(fminf(fmaxf(rp[x], 0.0f), 255.0f)) - (int)op[x];
            row += (unsigned int)(d * d);
        }
        acc += row;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    __shared__ unsigned long long red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) acc += red[w];
        if (acc) atomicAdd(sse + n, acc);
    }
}


__global__ void __launch_bounds__(512) pack_pu_kernel(const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                                                      const float *b3, const float *w4, const float *b4, float *packed)
{
    for (int i = threadIdx.x; i < WPACK; i += blockDim.x) {
        float v = 0.0f;
        if (i < B1_OFF) { const int k = i / 16, co = i % 16; v = w1[co * 9 + k]; }
        else if (i < W2_OFF) v = b1[i - B1_OFF];
        else if (i < B2_OFF) { const int j = i - W2_OFF; const int ci = j / 144, k = (j % 144) / 16, co = j % 16; v = w2[(co * 16 + ci) * 9 + k]; }
        else if (i < W3_OFF) v = b2[i - B2_OFF];
        else if (i < B3_OFF) { const int j = i - W3_OFF; const int ci = j / 144, k = (j % 144) / 16, co = j % 16; v = w3[(co * 16 + ci) * 9 + k]; }
        else if (i < W4_OFF) v = b3[i - B3_OFF];
        else if (i < B4_OFF) { const int j = i - W4_OFF; const int ci = j / 12, k = j % 12; v = (k < 9) ? w4[ci * 9 + k] : 0.0f; }
        else if (i == B4_OFF) v = b4[0];
        packed[i] = v;
    }
    // tensor-core operands of conv2 / conv3: W = rint(w * 2^Sw), Sw = 22 - e with max|w| = m * 2^e, m in [0.5, 1);
    // three signed-byte digits W = e0*2^16 + e1*2^8 + e2; image layout (tests/umma_ref.py:pack_weights):
    //   tap pairs tp = 0..3 and tap 8 alone (tp = 4, second chunk zero): byte [tp*1536 + chunk*768 + (digit*16 + co)*16 + ci];
    //   tap 8 against the activation digit pair (d1 | d2): byte [7680 + chunk*1024 + (row block*16 + co)*16 + ci], weight digit j
    //   in row block j of chunk 0 and row block j + 1 of chunk 1 (the other row blocks zero)
    __shared__ float red[512];
    __shared__ int s_sw[2];
    int8_t *img = reinterpret_cast<int8_t *>(packed + Q_OFF);
    for (int layer = 0; layer < 2; ++layer) {
        const float *w = layer == 0 ? w2 : w3;
        float m = 0.0f;
        for (int i = threadIdx.x; i < 2304; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
        red[threadIdx.x] = m;
        __syncthreads();
        for (int o = 256; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            int e = 0;
            if (red[0] > 0.0f) frexpf(red[0], &e);
            s_sw[layer] = 22 - e;
            packed[SC_OFF + layer] = ldexpf(1.0f, -(22 + (22 - e)));
        }
        __syncthreads();
        const float up = ldexpf(1.0f, s_sw[layer]);
        for (int i = threadIdx.x; i < 5 * 2 * 16 * 16; i += blockDim.x) {
            const int ci = i & 15, co = (i >> 4) & 15, chunk = (i >> 8) & 1, tp = i >> 9;
            // tap pairs: (0,0)(0,1) | (1,0)(1,1) | (2,0)(2,1) | (0,2)(1,2) | (2,2) -
            int ky, kx;
            if (tp < 3) { ky = tp; kx = chunk; }
            else if (tp == 3) { ky = chunk; kx = 2; }
            else { ky = 2; kx = 2; }
            int V = 0;
            if (!(tp == 4 && chunk == 1)) V = __float2int_rn(w[(co * 16 + ci) * 9 + ky * 3 + kx] * up);
            const int d2 = (int)(signed char)(V & 0xFF);
            const int V1 = (V - d2) >> 8;
            const int d1 = (int)(signed char)(V1 & 0xFF);
            const int d0 = (V1 - d1) >> 8;
            int8_t *o = img + layer * QBYTES + tp * 1536 + chunk * 768 + co * 16 + ci;
            o[0] = (int8_t)d0; o[256] = (int8_t)d1; o[512] = (int8_t)d2;
            if (tp == 4 && chunk == 0) {
                int8_t *z = img + layer * QBYTES + 7680 + co * 16 + ci;
                z[0] = (int8_t)d0; z[256] = (int8_t)d1; z[512] = (int8_t)d2; z[768] = 0;
                z[1024] = 0; z[1024 + 256] = (int8_t)d0; z[1024 + 512] = (int8_t)d1; z[1024 + 768] = (int8_t)d2;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
static PlaneD to_dev(const pmctf_plane_t &p)
{
    PlaneD d;
    d.p = p.p; d.gs = p.gs; d.bs = p.bs; d.rs = p.rs; d.cs = p.cs; d.group_n = p.group_n;
    return d;
}

int launch_step_tc(const StepD &d, int src_kind, int *err_flag, cudaStream_t st); // pmctf_lift_tc.cu
int register_packed_weights(const float *packed, cudaStream_t st);                 // pmctf_lift_tc.cu
int release_packed_weights(const float *packed);                                   // pmctf_lift_tc.cu

// process-wide DEFAULT of the convolution arithmetic; descriptors may name a mode per call (conv_mode fields)
static std::atomic<int> g_conv_mode{PMCTF_CONV_TENSOR};

// Per device: the error word a tensor-core kernel sets when one of its bounded mbarrier waits gave up (+ the phase stamps of the
// timing build).  It lives in MAPPED PINNED host memory: the kernel writes it through the device alias, the host reads it
// without any copy or synchronisation -- every launch checks it first, so a timed-out kernel makes all later calls on that
// device fail with PMCTF_ETIMEOUT until pmctf_tc_clear_error().
static volatile int *g_tc_err_host[64] = {nullptr};
static int *g_tc_err_dev[64] = {nullptr};
static std::atomic<int> g_tc_err_state[64];   // 0 none, 1 being created, 2 ready

static int tc_err_buffer(bool create, volatile int **host, int **devp)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
    if (dev < 0 || dev >= 64) return PMCTF_EINVAL;
    if (g_tc_err_state[dev].load(std::memory_order_acquire) != 2) {
        if (!create) { *host = nullptr; *devp = nullptr; return 0; }
        int expect = 0;
        if (g_tc_err_state[dev].compare_exchange_strong(expect, 1)) {
            void *h = nullptr, *d = nullptr;   // [0] error flag, [2..33] phase stamps of the timing build
            cudaError_t e = cudaHostAlloc(&h, 2048, cudaHostAllocMapped | cudaHostAllocPortable);
            if (e == cudaSuccess) e = cudaHostGetDevicePointer(&d, h, 0);
            if (e != cudaSuccess) { g_tc_err_state[dev].store(0); return (int)e; }
            for (int i = 0; i < 512; ++i) ((int *)h)[i] = 0;
            g_tc_err_host[dev] = (volatile int *)h;
            g_tc_err_dev[dev] = (int *)d;
            g_tc_err_state[dev].store(2, std::memory_order_release);
        } else {
            while (g_tc_err_state[dev].load(std::memory_order_acquire) == 1) { }
            if (g_tc_err_state[dev].load(std::memory_order_acquire) != 2) return PMCTF_EINVAL;
        }
    }
    *host = g_tc_err_host[dev];
    *devp = g_tc_err_dev[dev];
    return 0;
}

// used by the other translation units that launch tensor-core kernels (pmctf_pp.cu)
int tc_watchdog(volatile int **host, int **dev) { return tc_err_buffer(true, host, dev); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static int resolve_conv_mode(int requested)
{
    if (requested == PMCTF_CONV_DEFAULT) return g_conv_mode.load(std::memory_order_relaxed);
    return requested;
}

static int launch_step(const StepD &d, int src_kind, int conv_mode, cudaStream_t st)
{
    conv_mode = resolve_conv_mode(conv_mode);
    if (conv_mode != PMCTF_CONV_TENSOR && conv_mode != PMCTF_CONV_FFMA) return PMCTF_EINVAL;
    if (conv_mode == PMCTF_CONV_TENSOR) {
        volatile int *herr = nullptr;
        int *derr = nullptr;
        int e = tc_err_buffer(true, &herr, &derr);
        if (e) return e;
        if (herr[0] != 0) return PMCTF_ETIMEOUT;   // an earlier tensor-core launch on this device gave up: its outputs are incomplete
        e = launch_step_tc(d, src_kind, derr, st);
        if (e == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
        return e;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e;
        e = cudaFuncSetAttribute(lift_step_kernel<PMCTF_SRC_PLANE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(lift_step_kernel<PMCTF_SRC_WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(lift_step_kernel<PMCTF_SRC_SKIP3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid((d.w + TW - 1) / TW, (d.h + TH - 1) / TH, d.n);
    if (grid.y > 65535 || grid.z > 65535) return PMCTF_ESHAPE;
    switch (src_kind) {
    case PMCTF_SRC_PLANE: lift_step_kernel<PMCTF_SRC_PLANE><<<grid, NT, SMEM_BYTES, st>>>(d); break;
    case PMCTF_SRC_WARP: lift_step_kernel<PMCTF_SRC_WARP><<<grid, NT, SMEM_BYTES, st>>>(d); break;
    case PMCTF_SRC_SKIP3: lift_step_kernel<PMCTF_SRC_SKIP3><<<grid, NT, SMEM_BYTES, st>>>(d); break;
    default: return PMCTF_EINVAL;
    }
    return PMCTF_LAUNCHED();
}

// validated conversion of the public step descriptor
static int run_step(const pmctf_step_t &s, cudaStream_t st, int div_group_n = 0x7fffffff, float base_div1_g1 = 1.0f)
{
    if (s.conv_mode != PMCTF_CONV_DEFAULT && s.conv_mode != PMCTF_CONV_FFMA && s.conv_mode != PMCTF_CONV_TENSOR) return PMCTF_EINVAL;
    if (s.n <= 0 || s.h <= 0 || s.w <= 0 || !s.pu_packed || !s.out.p || !s.src.p) return PMCTF_EINVAL;
    if (s.mode < 0 || s.mode > 2) return PMCTF_EINVAL;
    if (s.mode == PMCTF_MODE_ACCUM && !s.base.p) return PMCTF_EINVAL;
    if (s.src_kind == PMCTF_SRC_SKIP3 && s.h < 2) return PMCTF_ESHAPE; // reflection needs two rows
    if (s.src_kind == PMCTF_SRC_WARP) {
        if (!s.mv || !s.lin_x || !s.lin_y) return PMCTF_EINVAL;
        if (s.mv_n < 1 || s.n % s.mv_n) return PMCTF_ESHAPE;
        if (s.h < 2 || s.w < 2) return PMCTF_ESHAPE;
        if (s.aux.p) return PMCTF_EINVAL;
    }
    StepD d;
    d.n = s.n; d.div_group_n = div_group_n; d.h = s.h; d.w = s.w; d.mode = s.mode;
    d.src = to_dev(s.src); d.src_div1 = s.src_div1; d.src_div2 = s.src_div2;
    d.mv = s.mv; d.mv_share = s.mv_n > 0 ? s.n / s.mv_n : 1; d.mv_down = s.mv_down;
    d.mv_h = s.mv_down ? 2 * s.h : s.h; d.mv_w = s.mv_down ? 2 * s.w : s.w;
    d.mv_sign = s.mv_sign;
    d.sx = (float)(((double)s.w - 1.0) / 2.0); d.sy = (float)(((double)s.h - 1.0) / 2.0);
    d.lin_x = s.lin_x; d.lin_y = s.lin_y; d.round_src = s.round_src;
    d.tap0 = s.tap0; d.tap1 = s.tap1; d.tap2 = s.tap2; d.tap_bias = s.tap_bias;
    d.pu_packed = s.pu_packed; d.in_mul = s.in_mul; d.post_mul = s.post_mul; d.out_mul = s.out_mul;
    d.round_tmp = s.round_tmp;
    d.base = to_dev(s.base); d.base_div1 = s.base_div1; d.base_div2 = s.base_div2;
    d.base_div1_g1 = base_div1_g1;
    d.sign = s.sign; d.final_mul = s.final_mul;
    d.out = to_dev(s.out); d.pred = to_dev(s.pred); d.aux = to_dev(s.aux);
    d.aux_mul = s.aux_mul;
    return launch_step(d, s.src_kind, s.conv_mode, st);
}

static pmctf_plane_t dense(const float *p, int H, int W)
{
    pmctf_plane_t pl = {};
    pl.p = const_cast<float *>(p); pl.bs = (long long)H * W; pl.rs = W; pl.cs = 1;
    return pl;
}

// two groups of `group_n` dense planes each, the second group starting at p1
static pmctf_plane_t two_groups(pmctf_plane_t pl, const float *p1, int group_n)
{
    pl.gs = p1 - pl.p;
    pl.group_n = group_n;
    return pl;
}

static pmctf_step_t blank_step(int n, int h, int w)
{
    pmctf_step_t s = {};
    s.n = n; s.h = h; s.w = w;
    s.src_div1 = s.src_div2 = s.base_div1 = s.base_div2 = 1.0f;
    s.in_mul = s.post_mul = s.out_mul = s.final_mul = s.aux_mul = 1.0f;
    s.sign = 1.0f; s.mv_sign = 1.0f;
    return s;
}

} // namespace pmctf

using namespace pmctf;

// ==========================================================================================
extern "C" {

int pmctf_abi_version(void) { return PMCTF_ABI_VERSION; }

unsigned long long pmctf_launch_count(void) { return pmctf::g_launches.load(std::memory_order_relaxed); }

int pmctf_set_conv_mode(int mode)
{
    if (mode != PMCTF_CONV_FFMA && mode != PMCTF_CONV_TENSOR) return PMCTF_EINVAL;
    pmctf::g_conv_mode.store(mode, std::memory_order_relaxed);
    return 0;
}

int pmctf_get_conv_mode(void) { return pmctf::g_conv_mode.load(std::memory_order_relaxed); }

int pmctf_tc_debug_times(long long *out16)
{
    if (!out16) return PMCTF_EINVAL;
    for (int i = 0; i < 16; ++i) out16[i] = 0;
    volatile int *herr = nullptr;
    int *derr = nullptr;
    if (pmctf::tc_err_buffer(false, &herr, &derr) || !herr) return 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return (int)cudaGetLastError();
    volatile long long *t = reinterpret_cast<volatile long long *>(herr + 2);
    for (int i = 0; i < 16; ++i) { out16[i] = t[i]; t[i] = 0; }
    return 0;
}

int pmctf_tc_error_flag(void)
{
    volatile int *herr = nullptr;
    int *derr = nullptr;
    if (pmctf::tc_err_buffer(false, &herr, &derr)) return -1;
    return herr ? herr[0] : 0;
}

int pmctf_tc_clear_error(void)
{
    volatile int *herr = nullptr;
    int *derr = nullptr;
    if (pmctf::tc_err_buffer(false, &herr, &derr)) return PMCTF_EINVAL;
    if (herr) herr[0] = 0;
    return 0;
}

int pmctf_tc_inject_timeout(void)
{
    volatile int *herr = nullptr;
    int *derr = nullptr;
    const int e = pmctf::tc_err_buffer(true, &herr, &derr);
    if (e) return e;
    herr[0] = 1;
    return 0;
}

int pmctf_release_pu_weights(const float *packed)
{
    if (!packed) return PMCTF_EINVAL;
    return pmctf::release_packed_weights(packed);
}

const char *pmctf_error_string(int code)
{
    switch (code) {
    case 0: return "success";
    case PMCTF_EINVAL: return "pmctf: invalid argument (null pointer, non-positive size or bad flag)";
    case PMCTF_ESHAPE: return "pmctf: unsupported shape";
    case PMCTF_EWORKSPACE: return "pmctf: workspace too small";
    case PMCTF_ETIMEOUT: return "pmctf: a tensor-core kernel on this device gave up waiting for its MMAs (outputs of that launch are incomplete); pmctf_tc_clear_error() re-arms the device";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "pmctf: unknown error";
    }
}

int pmctf_pack_pu_weights(const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                          const float *b3, const float *w4, const float *b4, float *packed, void *stream)
{
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w4 || !b4 || !packed) return PMCTF_EINVAL;
    if (((uintptr_t)packed & 15) != 0) return PMCTF_EINVAL;   // the tensor-core kernel fetches the operand images by TMA bulk copies
    if (pmctf_tc_error_flag() != 0) return PMCTF_ETIMEOUT;
    pack_pu_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(w1, b1, w2, b2, w3, b3, w4, b4, packed);
    const int e = PMCTF_LAUNCHED();
    if (e) return e;
    // the tensor-core kernel takes the small fp32 parameters as kernel arguments: keep a host copy per packed block
    // (one 40 KB read-back and stream synchronisation per weight version)
    return pmctf::register_packed_weights(packed, (cudaStream_t)stream);
}

int pmctf_flow_warp(const float *im, const float *flow, const float *lin_x, const float *lin_y, float *out, int N,
                    int C, int H, int W, int flowN, float sign, int round_out, void *stream)
{
    if (!im || !flow || !lin_x || !lin_y || !out || N <= 0 || C <= 0) return PMCTF_EINVAL;
    if (H < 2 || W < 2 || flowN < 1 || (N % flowN) || H > 65535 || N > 65535) return PMCTF_ESHAPE;
    const float sx = (float)(((double)W - 1.0) / 2.0), sy = (float)(((double)H - 1.0) / 2.0);
    if (H >= 8 && (H + WR - 1) / WR <= 65535 && (long long)H * W < (1LL << 30) && (long long)N * C < (1LL << 31)) {
        flow_warp4_kernel<<<dim3((W + 127) / 128, (H + WR - 1) / WR, N), 128, 0, (cudaStream_t)stream>>>(im, flow, lin_x, lin_y, out, N, C, H, W,
                                                                                                 N / flowN, sign, sx, sy, round_out);
    } else {
        flow_warp_kernel<<<dim3((W + 255) / 256, H, N), 256, 0, (cudaStream_t)stream>>>(im, flow, lin_x, lin_y, out, N, C, H, W, flowN, sign,
                                                                                     sx, sy, round_out);
    }
    return PMCTF_LAUNCHED();
}

int pmctf_chroma_mv_down(const float *mv, float *out, int N, int H, int W, void *stream)
{
    if (!mv || !out || N <= 0) return PMCTF_EINVAL;
    if (H < 2 || W < 2 || (H & 1) || (W & 1) || H / 2 > 65535 || 2 * N > 65535) return PMCTF_ESHAPE;
    dim3 grid((W / 2 + 255) / 256, H / 2, 2 * N);
    chroma_mv_down_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mv, out, 2 * N, H, W);
    return PMCTF_LAUNCHED();
}

int pmctf_lift_step(const pmctf_step_t *step, void *stream)
{
    if (!step) return PMCTF_EINVAL;
    return run_step(*step, (cudaStream_t)stream);
}

int pmctf_predict_update(const float *x, const float *pu_packed, float in_mul, float *out, int N, int H, int W,
                         int conv_mode, void *stream)
{
    if (!x || !out || !pu_packed) return PMCTF_EINVAL;
    pmctf_step_t s = blank_step(N, H, W);
    s.conv_mode = conv_mode;
    s.src_kind = PMCTF_SRC_PLANE; s.mode = PMCTF_MODE_PU;
    s.src = dense(x, H, W); s.out = dense(out, H, W);
    s.pu_packed = pu_packed; s.in_mul = in_mul;
    return run_step(s, (cudaStream_t)stream);
}

int pmctf_temporal_filter(const float *x, const pmctf_temporal_t *t, int which, float *out, int N, int H, int W,
                          void *stream)
{
    if (!x || !out || !t || (which != 0 && which != 1)) return PMCTF_EINVAL;
    pmctf_step_t s = blank_step(N, H, W);
    s.conv_mode = t->conv_mode;
    s.src_kind = PMCTF_SRC_PLANE; s.mode = PMCTF_MODE_FILTER;
    s.src = dense(x, H, W); s.out = dense(out, H, W);
    s.pu_packed = which == 0 ? t->P_t_packed : t->U_t_packed;
    s.out_mul = t->lossy ? (which == 0 ? t->scale_p : t->scale_u) : 1.0f;
    s.round_tmp = !t->lossy;
    return run_step(s, (cudaStream_t)stream);
}

static pmctf_step_t mctf_step(const pmctf_plane_t *src, const pmctf_plane_t *base, const pmctf_plane_t *out,
                              const pmctf_plane_t *pred, const float *mv, int mv_n, int mv_down, float mv_sign,
                              const float *lin_x, const float *lin_y, const float *packed, float out_mul, float sign,
                              int lossy, int N, int H, int W, int conv_mode)
{
    pmctf_step_t s = blank_step(N, H, W);
    s.conv_mode = conv_mode;
    s.src_kind = PMCTF_SRC_WARP; s.mode = PMCTF_MODE_ACCUM;
    s.src = *src; s.base = *base; s.out = *out;
    if (pred && pred->p) s.pred = *pred;
    s.mv = mv; s.mv_n = mv_n; s.mv_down = mv_down; s.mv_sign = mv_sign; s.lin_x = lin_x; s.lin_y = lin_y;
    s.round_src = !lossy; s.round_tmp = !lossy;
    s.pu_packed = packed; s.out_mul = lossy ? out_mul : 1.0f; s.sign = sign;
    return s;
}

int pmctf_forward_mctf(const pmctf_plane_t *ref, const pmctf_plane_t *cur, const float *mv, int mv_n, int mv_down,
                       const float *lin_x, const float *lin_y, const pmctf_temporal_t *t, const pmctf_plane_t *L,
                       const pmctf_plane_t *Hh, const pmctf_plane_t *pred, const pmctf_plane_t *inv, int N, int H, int W,
                       void *stream)
{
    if (!ref || !cur || !mv || !t || !L || !Hh || !ref->p || !cur->p || !L->p || !Hh->p) return PMCTF_EINVAL;
    // H_t = cur - predict(warp(ref, mv))                 pMCTF_L.py:301-305
    pmctf_step_t s1 = mctf_step(ref, cur, Hh, pred, mv, mv_n, mv_down, 1.0f, lin_x, lin_y, t->P_t_packed, t->scale_p,
                                -1.0f, t->lossy, N, H, W, t->conv_mode);
    int e = run_step(s1, (cudaStream_t)stream);
    if (e) return e;
    // L_t = ref + update(warp(H_t, -mv))                 pMCTF_L.py:307-311
    pmctf_step_t s2 = mctf_step(Hh, ref, L, inv, mv, mv_n, mv_down, -1.0f, lin_x, lin_y, t->U_t_packed, t->scale_u,
                                1.0f, t->lossy, N, H, W, t->conv_mode);
    return run_step(s2, (cudaStream_t)stream);
}

int pmctf_inverse_mctf(const pmctf_plane_t *L, const pmctf_plane_t *Hh, const float *mv, int mv_n, int mv_down,
                       const float *lin_x, const float *lin_y, const pmctf_temporal_t *t, const pmctf_plane_t *ref,
                       const pmctf_plane_t *cur, int N, int H, int W, void *stream)
{
    if (!L || !Hh || !mv || !t || !ref || !cur || !L->p || !Hh->p || !ref->p || !cur->p) return PMCTF_EINVAL;
    // ref = L - update(warp(H, -mv))                     pMCTF_L.py:320-324
    pmctf_step_t s1 = mctf_step(Hh, L, ref, nullptr, mv, mv_n, mv_down, -1.0f, lin_x, lin_y, t->U_t_packed, t->scale_u,
                                -1.0f, t->lossy, N, H, W, t->conv_mode);
    int e = run_step(s1, (cudaStream_t)stream);
    if (e) return e;
    // cur = H + predict(warp(ref, mv))                   pMCTF_L.py:325-329
    pmctf_step_t s2 = mctf_step(ref, Hh, cur, nullptr, mv, mv_n, mv_down, 1.0f, lin_x, lin_y, t->P_t_packed, t->scale_p,
                                1.0f, t->lossy, N, H, W, t->conv_mode);
    return run_step(s2, (cudaStream_t)stream);
}

// one spatial lifting step on logical planes: out = (base/bd1/bd2 + sign*(skip + 0.1*256*PU(skip/256))) * fm
struct SpatialDivs {
    float sd1 = 1.0f, sd2 = 1.0f, bd1 = 1.0f, bd2 = 1.0f;
    int div_group_n = 0x7fffffff; // planes >= this index divide the base by bd1_g1 instead of bd1
    float bd1_g1 = 1.0f;
};

static int spatial_step(const pmctf_iwave_t *p, int which, const pmctf_plane_t &src, const pmctf_plane_t &base,
                        const pmctf_plane_t &out, float sign, float final_mul, const pmctf_plane_t *aux, float aux_mul,
                        int n, int h, int w, cudaStream_t st, const SpatialDivs &dv = SpatialDivs())
{
    pmctf_step_t s = blank_step(n, h, w);
    s.conv_mode = p->conv_mode;
    s.src_kind = PMCTF_SRC_SKIP3; s.mode = PMCTF_MODE_ACCUM;
    s.src = src; s.src_div1 = dv.sd1; s.src_div2 = dv.sd2;
    s.tap0 = p->tap[which][0]; s.tap1 = p->tap[which][1]; s.tap2 = p->tap[which][2]; s.tap_bias = p->bias[which];
    s.pu_packed = p->pu_packed + (long long)which * PMCTF_PU_PACKED_FLOATS;
    s.in_mul = 1.0f / p->dynamic_range; // exact: dynamic_range is a power of two (lifting_1d.py:62,108)
    s.post_mul = p->dynamic_range;
    s.round_tmp = !p->lossy;
    s.base = base; s.base_div1 = dv.bd1; s.base_div2 = dv.bd2; s.sign = sign; s.final_mul = final_mul;
    s.out = out;
    if (aux) { s.aux = *aux; s.aux_mul = aux_mul; }
    return run_step(s, st, dv.div_group_n, dv.bd1_g1);
}

static pmctf_plane_t phase(const pmctf_plane_t &x, int odd)
{ // split: lifting_1d.py:10-13
    pmctf_plane_t p = x;
    p.p = x.p + (odd ? x.rs : 0);
    p.rs = 2 * x.rs;
    return p;
}

static bool pow2_range(float r) { return r >= 1.0f && r <= 65536.0f && (float)(int)r == r && (((int)r) & ((int)r - 1)) == 0; }

// forward_lift on logical planes; ws holds n*h2*w floats (the unscaled high band)
static int iwave_forward(const pmctf_plane_t &x, const pmctf_iwave_t *p, const pmctf_plane_t &l, const pmctf_plane_t &hh,
                         float *ws, int n, int h2, int w, cudaStream_t st)
{
    if (!pow2_range(p->dynamic_range)) return PMCTF_EINVAL;
    pmctf_plane_t xe = phase(x, 0), xo = phase(x, 1);
    pmctf_plane_t hu = dense(ws, h2, w);
    const float sl = p->lossy ? p->scale_l : 1.0f, sh = p->lossy ? p->scale_h : 1.0f;
    int e;
    // P1: x_o += f(x_e)   lifting_1d.py:104-112
    if ((e = spatial_step(p, 0, xe, xo, hu, +1, 1.0f, nullptr, 1, n, h2, w, st))) return e;
    // U1: x_e += f(x_o)   :114-122
    if ((e = spatial_step(p, 1, hu, xe, l, +1, 1.0f, nullptr, 1, n, h2, w, st))) return e;
    // P2                  :124-132
    if ((e = spatial_step(p, 2, l, hu, hu, +1, 1.0f, nullptr, 1, n, h2, w, st))) return e;
    // U2 + scaling        :134-143   (l *= scale_l in the epilogue, h * scale_h copied through)
    return spatial_step(p, 3, hu, l, l, +1, sl, &hh, sh, n, h2, w, st);
}

// backward_lift; ws holds 2*n*h2*w floats.  l_div / l_div_g1 / h_div: fused dequantise (pWave.py:191-202)
static int iwave_backward(const pmctf_plane_t &l, float l_div, int div_group_n, float l_div_g1, const pmctf_plane_t &hh,
                          float h_div, const pmctf_iwave_t *p, const pmctf_plane_t &x, float *ws, int n, int h2, int w,
                          cudaStream_t st)
{
    if (!pow2_range(p->dynamic_range)) return PMCTF_EINVAL;
    pmctf_plane_t xe = phase(x, 0), xo = phase(x, 1);
    pmctf_plane_t lw = dense(ws, h2, w), hw = dense(ws + (long long)n * h2 * w, h2, w);
    const float sl = p->lossy ? p->scale_l : 1.0f, sh = p->lossy ? p->scale_h : 1.0f;
    int e;
    // l = l/scale_l - f(h/scale_h)      lifting_1d.py:148-159
    SpatialDivs dv;
    dv.sd1 = h_div; dv.sd2 = sh; dv.bd1 = l_div; dv.bd2 = sl; dv.div_group_n = div_group_n; dv.bd1_g1 = l_div_g1;
    if ((e = spatial_step(p, 3, hh, l, lw, -1, 1.0f, &hw, 1.0f, n, h2, w, st, dv))) return e;
    if ((e = spatial_step(p, 2, lw, hw, hw, -1, 1.0f, nullptr, 1, n, h2, w, st))) return e; // :161-168
    if ((e = spatial_step(p, 1, hw, lw, lw, -1, 1.0f, nullptr, 1, n, h2, w, st))) return e; // :170-177
    // P1 + merge (:179-189, :16-22): odd rows = h - f(l), even rows = l copied through
    return spatial_step(p, 0, lw, hw, xo, -1, 1.0f, &xe, 1.0f, n, h2, w, st);
}

int pmctf_iwave1d_forward(const pmctf_plane_t *x, const pmctf_iwave_t *p, const pmctf_plane_t *l, const pmctf_plane_t *h,
                          int n, int h2, int w, float *workspace, long long workspace_floats, void *stream)
{
    if (!x || !p || !l || !h || !x->p || !l->p || !h->p || !p->pu_packed || !workspace || n <= 0) return PMCTF_EINVAL;
    if (h2 < 2 || w < 1) return PMCTF_ESHAPE;
    if (workspace_floats < (long long)n * h2 * w) return PMCTF_EWORKSPACE;
    return iwave_forward(*x, p, *l, *h, workspace, n, h2, w, (cudaStream_t)stream);
}

int pmctf_iwave1d_backward(const pmctf_plane_t *l, const pmctf_plane_t *h, const pmctf_iwave_t *p, const pmctf_plane_t *x,
                           int n, int h2, int w, float *workspace, long long workspace_floats, void *stream)
{
    if (!x || !p || !l || !h || !x->p || !l->p || !h->p || !p->pu_packed || !workspace || n <= 0) return PMCTF_EINVAL;
    if (h2 < 2 || w < 1) return PMCTF_ESHAPE;
    if (workspace_floats < 2LL * n * h2 * w) return PMCTF_EWORKSPACE;
    return iwave_backward(*l, 1.0f, 0x7fffffff, 1.0f, *h, 1.0f, p, *x, workspace, n, h2, w, (cudaStream_t)stream);
}

long long pmctf_lift2d_workspace(int N, int H, int W) { return 2LL * N * H * W; }

// transposed logical view of a dense [n, rows, cols] buffer: logical (y', x') = physical (x', y')
static pmctf_plane_t transposed(const float *p, int rows, int cols)
{
    pmctf_plane_t t = {};
    t.p = const_cast<float *>(p); t.bs = (long long)rows * cols; t.rs = 1; t.cs = cols;
    return t;
}

int pmctf_lift2d_forward(const float *x, const pmctf_iwave_t *p, float *ll, float *lh, float *hl, float *hh, float *l_out,
                         float *h_out, int N, int H, int W, float *workspace, long long workspace_floats, void *stream)
{
    if (!x || !p || !ll || !lh || !hl || !hh || !workspace || !p->pu_packed || N <= 0) return PMCTF_EINVAL;
    if ((H & 1) || (W & 1) || H < 4 || W < 4) return PMCTF_ESHAPE;
    if (workspace_floats < pmctf_lift2d_workspace(N, H, W)) return PMCTF_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int h2 = H / 2, w2 = W / 2;
    const long long half = (long long)N * h2 * W;
    float *ws0 = workspace, *lbuf = l_out ? l_out : workspace + half, *hbuf = h_out ? h_out : workspace + 2 * half;
    float *ws1 = workspace + 3 * half;
    // rows: wavelet_transform.py:29
    int e = iwave_forward(dense(x, H, W), p, dense(lbuf, h2, W), dense(hbuf, h2, W), ws0, N, h2, W, st);
    if (e) return e;
    // columns of l and of h as ONE batch of 2N transposed logical planes [W, h2] (:32-40):
    // group 0 = l -> (ll, lh), group 1 = h -> (hl, hh)
    pmctf_plane_t xt = two_groups(transposed(lbuf, h2, W), hbuf, N);
    pmctf_plane_t lo = two_groups(transposed(ll, h2, w2), hl, N), ho = two_groups(transposed(lh, h2, w2), hh, N);
    return iwave_forward(xt, p, lo, ho, ws1, 2 * N, w2, h2, st);
}

int pmctf_lift2d_backward_q(const float *ll, const float *lh, const float *hl, const float *hh, float ll_div, float q,
                            const pmctf_iwave_t *p, float *x, int N, int H, int W, float *workspace,
                            long long workspace_floats, void *stream)
{
    if (!x || !p || !ll || !lh || !hl || !hh || !workspace || !p->pu_packed || N <= 0) return PMCTF_EINVAL;
    if (!(ll_div > 0.0f) || !(q > 0.0f)) return PMCTF_EINVAL;
    if ((H & 1) || (W & 1) || H < 4 || W < 4) return PMCTF_ESHAPE;
    if (workspace_floats < pmctf_lift2d_workspace(N, H, W)) return PMCTF_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int h2 = H / 2, w2 = W / 2;
    const long long half = (long long)N * h2 * W;
    float *lbuf = workspace, *hbuf = workspace + half, *ws = workspace + 2 * half;
    // columns (wavelet_transform.py:46-54) as ONE batch of 2N transposed logical planes:
    // group 0 = (ll, lh) -> l, group 1 = (hl, hh) -> h.  The fused dequantise (pWave.py:191-202)
    // divides ll by ll_div and the three detail bands by q.
    pmctf_plane_t li = two_groups(transposed(ll, h2, w2), hl, N), hi = two_groups(transposed(lh, h2, w2), hh, N);
    pmctf_plane_t xt = two_groups(transposed(lbuf, h2, W), hbuf, N);
    int e = iwave_backward(li, ll_div, N, q, hi, q, p, xt, ws, 2 * N, w2, h2, st);
    if (e) return e;
    // rows (:56)
    return iwave_backward(dense(lbuf, h2, W), 1.0f, 0x7fffffff, 1.0f, dense(hbuf, h2, W), 1.0f, p, dense(x, H, W), ws, N,
                          h2, W, st);
}

int pmctf_lift2d_backward(const float *ll, const float *lh, const float *hl, const float *hh, const pmctf_iwave_t *p,
                          float *x, int N, int H, int W, float *workspace, long long workspace_floats, void *stream)
{
    return pmctf_lift2d_backward_q(ll, lh, hl, hh, 1.0f, 1.0f, p, x, N, H, W, workspace, workspace_floats, stream);
}

int pmctf_quantize(const float *s, float q, float clip, int lossy, int do_round, float *out, long long n, void *stream)
{
    if (n == 0) return 0;
    if (!s || !out || n < 0) return PMCTF_EINVAL;
    const int vec = ((((uintptr_t)s | (uintptr_t)out) & 15) == 0);
    long long blocks = ((vec ? n / 4 + 3 : n) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    quantize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(s, q, clip, lossy, do_round, out, n, vec);
    return PMCTF_LAUNCHED();
}

int pmctf_dequantize(const float *s_hat, float q, int lossy, float *out, long long n, void *stream)
{
    if (n == 0) return 0;
    if (!s_hat || !out || n < 0) return PMCTF_EINVAL;
    const int vec = ((((uintptr_t)s_hat | (uintptr_t)out) & 15) == 0);
    long long blocks = ((vec ? n / 4 + 3 : n) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    dequantize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(s_hat, q, lossy, out, n, vec);
    return PMCTF_LAUNCHED();
}

int pmctf_quantize_stats(const float *s, float q, float clip, int lossy, float *out, int planes, long long plane_elems,
                         unsigned long long *stats, void *stream)
{
    if (planes == 0 || plane_elems == 0) return 0;
    if (!s || !out || !stats || planes < 0 || plane_elems < 0 || planes > 65535) return PMCTF_EINVAL;
    const int vec = ((((uintptr_t)s | (uintptr_t)out) & 15) == 0) && (plane_elems % 4 == 0 || planes == 1);
    long long bx = ((vec ? plane_elems / 4 + 3 : plane_elems) + 255) / 256;
    const long long cap = (148 * 16 + planes - 1) / planes;
    if (bx > cap) bx = cap;
    quantize_stats_kernel<<<dim3((unsigned)bx, planes), 256, 0, (cudaStream_t)stream>>>(s, q, clip, lossy, out, plane_elems, stats, vec);
    return PMCTF_LAUNCHED();
}

int pmctf_quantize_code(const float *s, const float *q_per_plane, float clip, int lossy, int dequant, float *out, short *sym16,
                        long long sym16_plane_stride, int planes, long long plane_elems, unsigned long long *stats, void *stream)
{
    if (planes == 0 || plane_elems == 0) return 0;
    if (!s || !out || !q_per_plane || planes < 0 || plane_elems < 0 || planes > 65535) return PMCTF_EINVAL;
    if (sym16 && sym16_plane_stride < plane_elems) return PMCTF_EINVAL;
    if (clip > 32767.0f && sym16) return PMCTF_EINVAL;   // the symbols must fit the int16 the entropy coder takes (pWave.py:55-58)
    int vec = ((((uintptr_t)s | (uintptr_t)out) & 15) == 0) && (plane_elems % 4 == 0 || planes == 1);
    if (sym16 && ((((uintptr_t)sym16) & 7) != 0 || (sym16_plane_stride % 4 != 0 && planes > 1))) vec = 0;
    long long bx = ((vec ? plane_elems / 4 + 3 : plane_elems) + 255) / 256;
    const long long cap = (148 * 16 + planes - 1) / planes;
    if (bx > cap) bx = cap;
    quantize_code_kernel<<<dim3((unsigned)bx, planes), 256, 0, (cudaStream_t)stream>>>(s, q_per_plane, clip, lossy, dequant, out, sym16,
                                                                                      sym16_plane_stride, plane_elems, stats, vec);
    return PMCTF_LAUNCHED();
}

int pmctf_unpack_u8(const unsigned char *src, float *dst, int n, int h0, int w0, int hp, int wp, void *stream)
{
    if (n == 0) return 0;
    if (!src || !dst || n < 0 || h0 <= 0 || w0 <= 0) return PMCTF_EINVAL;
    if (hp < h0 || wp < w0 || (wp & 3) || hp > 65535 || n > 65535) return PMCTF_ESHAPE;
    unpack_u8_kernel<<<dim3((wp / 4 + 255) / 256, (hp + UNPACK_ROWS - 1) / UNPACK_ROWS, n), 256, 0, (cudaStream_t)stream>>>(src, dst, h0, w0, hp, wp);
    return PMCTF_LAUNCHED();
}

int pmctf_frame_sse(const float *rec, const unsigned char *orig, int n, int h0, int w0, int hp, int wp,
                    unsigned long long *sse, void *stream)
{
    if (n == 0) return 0;
    if (!rec || !orig || !sse || n < 0 || h0 <= 0 || w0 <= 0) return PMCTF_EINVAL;
    if (hp < h0 || wp < w0 || h0 > 65535 || n > 65535) return PMCTF_ESHAPE;
    int bx = (148 * 8 + n - 1) / n;   // about eight blocks per SM over all planes
    if (bx > h0) bx = h0;
    const int vec = (w0 % 4 == 0) && (wp % 4 == 0) && ((((uintptr_t)rec) & 15) == 0) && ((((uintptr_t)orig) & 3) == 0);
    frame_sse_kernel<<<dim3(bx, n), 256, 0, (cudaStream_t)stream>>>(rec, orig, h0, w0, hp, wp, sse, vec);
    return PMCTF_LAUNCHED();
}

} // extern "C"
