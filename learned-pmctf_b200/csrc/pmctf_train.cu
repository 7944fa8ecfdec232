// pmctf_train.cu -- differentiable primitives of the TRAINING path (BASELINE.json configs[4]: forward + backward through
// warp, lifting and convs).  Evaluation uses the fused tensor-core lifting step (pmctf_lift_tc.cu); under autograd the
// modules compose these un-fused fp32 kernels instead, exactly like the reference composes conv2d / grid_sample:
//
//   conv3x3_kernel<CIN,COUT>   y = conv3x3(x, w) + b, zero padding, NCHW (layers.py:54-56).  The data gradient is the
//                              same kernel run with the transposed, flipped weights.
//   wgrad3x3_kernel<CIN,COUT>  dL/dw[co][ci][ky][kx] = sum g[co](p) x[ci](p + k), dL/db[co] = sum g[co]   (atomics)
//   flow_warp_bwd_kernel       adjoint of flow_warp (video_net.py:32-55 / ATen grid_sampler_2d backward, bilinear, border,
//                              align_corners=True): scatter-add into dL/dim, dL/dflow
// Floating-point, tolerance-tested against torch autograd (tests/test_gpu_train.py); no exactness contract here.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pmctf_b200.h"

namespace pmctf {
namespace train {

constexpr int TX = 32, TY = 8; // output tile per CTA, one thread per pixel

// Fused epilogues of the PredictUpdate forward / backward chain (lifting_1d.py:36-49): what the reference leaves to separate
// element-wise ATen kernels over the 16-channel maps.  idx addresses y / y2 / aux / aux2 alike ([N,COUT,H,W]).
//   EPI_NONE      y = acc
//   EPI_TANH      y = tanh(acc)                               conv2: a2 = tanh(conv2(a1))
//   EPI_DUAL      y = acc, y2 = tanh(acc)                     conv1: c1 and a1 = tanh(c1)
//   EPI_ADD       y = acc + aux                               conv3: r = c1 + conv3(a2)
//   EPI_DTANH     y = acc * (1 - aux^2)                       data gradient through a tanh whose OUTPUT is aux
//   EPI_DTANH_ADD y = acc * (1 - aux^2) + aux2                the same plus the gradient of the residual branch
enum { EPI_NONE = 0, EPI_TANH = 1, EPI_DUAL = 2, EPI_ADD = 3, EPI_DTANH = 4, EPI_DTANH_ADD = 5 };
struct Epi {
    const float *aux, *aux2;
    float *y2;
    int mode;
};
__device__ __forceinline__ void epilogue(const Epi &e, float acc, float *y, long long idx)
{
    switch (e.mode) {
    case EPI_TANH: y[idx] = tanhf(acc); break;
    case EPI_DUAL: y[idx] = acc; e.y2[idx] = tanhf(acc); break;
    case EPI_ADD: y[idx] = acc + __ldg(e.aux + idx); break;
    case EPI_DTANH: { const float a = __ldg(e.aux + idx); y[idx] = acc * (1.0f - a * a); break; }
    case EPI_DTANH_ADD: { const float a = __ldg(e.aux + idx); y[idx] = acc * (1.0f - a * a) + __ldg(e.aux2 + idx); break; }
    default: y[idx] = acc;
    }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(TX * TY) conv3x3_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                                                          float *__restrict__ y, int H, int W, const Epi epi)
{
    __shared__ float xs[CIN][TY + 2][TX + 2];
    __shared__ __align__(16) float ws[CIN * 9][COUT]; // [ci*9 + k][co]
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int n = blockIdx.z, y0 = blockIdx.y * TY, x0 = blockIdx.x * TX;
    for (int i = tid; i < CIN * 9 * COUT; i += TX * TY) {
        const int co = i % COUT, r = i / COUT; // r = ci*9 + k
        ws[r][co] = w[(co * CIN + r / 9) * 9 + r % 9];
    }
    const float *xp = x + (long long)n * CIN * H * W;
    for (int i = tid; i < CIN * (TY + 2) * (TX + 2); i += TX * TY) {
        const int ci = i / ((TY + 2) * (TX + 2)), rem = i % ((TY + 2) * (TX + 2));
        const int r = rem / (TX + 2), c = rem % (TX + 2);
        const int gy = y0 - 1 + r, gx = x0 - 1 + c;
        xs[ci][r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? xp[((long long)ci * H + gy) * W + gx] : 0.0f;
    }
    __syncthreads();
    const int gy = y0 + threadIdx.y, gx = x0 + threadIdx.x;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = b ? b[co] : 0.0f;
#pragma unroll 1
    for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float v = xs[ci][threadIdx.y + k / 3][threadIdx.x + k % 3];
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[co] = fmaf(ws[ci * 9 + k][co], v, acc[co]);
        }
    }
    if (gy < H && gx < W) {
        const long long base = (long long)n * COUT * H * W + (long long)gy * W + gx;
#pragma unroll
        for (int co = 0; co < COUT; ++co) epilogue(epi, acc[co], y, base + (long long)co * H * W);
    }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(256) wgrad3x3_kernel(const float *__restrict__ x, const float *__restrict__ g, float *__restrict__ gw,
                                                       float *__restrict__ gb, int N, int H, int W)
{
    // persistent CTAs: each walks its share of the 32x8 tiles keeping the partial sums in registers, so the global atomics
    // (which all land on the same few hundred addresses) are paid once per CTA instead of once per tile
    constexpr int PAIRS = CIN * COUT, SPLIT = 256 / PAIRS;
    __shared__ float xs[CIN][TY + 2][TX + 2];
    __shared__ float gs[COUT][TY][TX];
    __shared__ float red[(SPLIT > 1) ? 256 * 10 : 1];
    const int tid = threadIdx.x;
    const int pair = tid % PAIRS, part = tid / PAIRS;   // one (co, ci) pair per thread, a tile's pixels split over `SPLIT` parts
    const int co = pair / CIN, ci = pair % CIN;
    const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
    const int per_img = tiles_x * tiles_y, total = per_img * N;
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, bsum = 0.0f;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int n = t / per_img, rem = t - n * per_img;
        const int y0 = (rem / tiles_x) * TY, x0 = (rem % tiles_x) * TX;
        const float *xp = x + (long long)n * CIN * H * W;
        const float *gp = g + (long long)n * COUT * H * W;
        __syncthreads();
        for (int i = tid; i < CIN * (TY + 2) * (TX + 2); i += 256) {
            const int c_ = i / ((TY + 2) * (TX + 2)), r2 = i % ((TY + 2) * (TX + 2));
            const int r = r2 / (TX + 2), c = r2 % (TX + 2);
            const int gy = y0 - 1 + r, gx = x0 - 1 + c;
            xs[c_][r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? xp[((long long)c_ * H + gy) * W + gx] : 0.0f;
        }
        for (int i = tid; i < COUT * TY * TX; i += 256) {
            const int c_ = i / (TY * TX), r2 = i % (TY * TX);
            const int r = r2 / TX, c = r2 % TX;
            const int gy = y0 + r, gx = x0 + c;
            gs[c_][r][c] = (gy < H && gx < W) ? gp[((long long)c_ * H + gy) * W + gx] : 0.0f;
        }
        __syncthreads();
        for (int p = part; p < TY * TX; p += SPLIT) {
            const int r = p / TX, c = p % TX;
            const float gv = gs[co][r][c];
            bsum += gv;
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = fmaf(gv, xs[ci][r + k / 3][c + k % 3], acc[k]);
        }
    }
    if (SPLIT > 1) { // fold the parts of a pair inside the CTA
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 9; ++k) red[tid * 10 + k] = acc[k];
        red[tid * 10 + 9] = bsum;
        __syncthreads();
        if (part == 0) {
            for (int q = 1; q < SPLIT; ++q) {
#pragma unroll
                for (int k = 0; k < 9; ++k) acc[k] += red[(pair + q * PAIRS) * 10 + k];
                bsum += red[(pair + q * PAIRS) * 10 + 9];
            }
        }
    }
    if (part == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) atomicAdd(gw + (co * CIN + ci) * 9 + k, acc[k]);
        if (gb && ci == 0) atomicAdd(gb + co, bsum);
    }
}

// ---- 16 -> 16 specialisations (94 % of the training FLOPs): two pixels / two channel pairs per thread and packed fp32x2 FMAs,
//      which halve both the shared-memory loads per FMA and the issue slots ----------------------------------------------
__device__ __forceinline__ float2 ffma2s(float2 a, float b, float2 c)
{
    unsigned long long d;
    const float2 bb = make_float2(b, b);
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long *>(&a)), "l"(*reinterpret_cast<const unsigned long long *>(&bb)),
          "l"(*reinterpret_cast<const unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&d);
}

constexpr int TX2 = 64; // conv16x16: 64 x 8 output tile, thread = 2 horizontally adjacent pixels x 16 output channels

__global__ void __launch_bounds__(256) conv3x3_16x16_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                                                            float *__restrict__ y, int H, int W, const Epi epi)
{
    extern __shared__ __align__(16) float dsm[];                         // 51.5 KB: opt-in dynamic shared memory
    float(*ws)[16] = reinterpret_cast<float(*)[16]>(dsm);                // [ci*9 + k][co]
    float(*xs)[TY + 2][TX2 + 2] = reinterpret_cast<float(*)[TY + 2][TX2 + 2]>(dsm + 144 * 16);
    const int tid = threadIdx.x;
    const int n = blockIdx.z, y0 = blockIdx.y * TY, x0 = blockIdx.x * TX2;
    for (int i = tid; i < 144 * 16; i += 256) {
        const int co = i % 16, r = i / 16;
        ws[r][co] = w[(co * 16 + r / 9) * 9 + r % 9];
    }
    const float *xp = x + (long long)n * 16 * H * W;
    for (int i = tid; i < 16 * (TY + 2) * (TX2 + 2); i += 256) {
        const int ci = i / ((TY + 2) * (TX2 + 2)), rem = i % ((TY + 2) * (TX2 + 2));
        const int r = rem / (TX2 + 2), c = rem % (TX2 + 2);
        const int gy = y0 - 1 + r, gx = x0 - 1 + c;
        xs[ci][r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? xp[((long long)ci * H + gy) * W + gx] : 0.0f;
    }
    __syncthreads();
    const int ty = tid / 32, tx = tid % 32;
    float2 acc[16]; // (pixel 2tx, pixel 2tx+1) per output channel
#pragma unroll
    for (int co = 0; co < 16; ++co) acc[co] = b ? make_float2(b[co], b[co]) : make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int ci = 0; ci < 16; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = xs[ci][ty + ky][2 * tx + j];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float2 vv = make_float2(v[kx], v[kx + 1]);
                const float4 *wr = reinterpret_cast<const float4 *>(ws[ci * 9 + ky * 3 + kx]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wq = wr[q];
                    acc[4 * q + 0] = ffma2s(vv, wq.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = ffma2s(vv, wq.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = ffma2s(vv, wq.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = ffma2s(vv, wq.w, acc[4 * q + 3]);
                }
            }
        }
    }
    const int gy = y0 + ty, gx = x0 + 2 * tx;
    if (gy < H && gx < W) {
        const long long base = (long long)n * 16 * H * W + (long long)gy * W + gx;
        const bool two = gx + 1 < W;
#pragma unroll
        for (int co = 0; co < 16; ++co) {
            epilogue(epi, acc[co].x, y, base + (long long)co * H * W);
            if (two) epilogue(epi, acc[co].y, y, base + (long long)co * H * W + 1);
        }
    }
}

// wgrad 16 x 16: thread = (2 output channels) x (2 input channels) x 9 taps, the tile's pixels split over 4 parts
__global__ void __launch_bounds__(256) wgrad3x3_16x16_kernel(const float *__restrict__ x, const float *__restrict__ g, float *__restrict__ gw,
                                                             float *__restrict__ gb, int N, int H, int W)
{
    __shared__ float xs[16][TY + 2][TX + 2];
    __shared__ float gs[16][TY][TX];
    __shared__ float red[2304 + 16];
    const int tid = threadIdx.x;
    const int blk = tid % 64, part = tid / 64;       // 8 x 8 blocks of (co pair, ci pair)
    const int co0 = 2 * (blk / 8), ci0 = 2 * (blk % 8);
    const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
    const int per_img = tiles_x * tiles_y, total = per_img * N;
    float2 acc[2][9]; // [ci offset][tap] -> (co0, co0+1)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[a][k] = make_float2(0.0f, 0.0f);
    float2 bsum = make_float2(0.0f, 0.0f);
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int n = t / per_img, rem = t - n * per_img;
        const int y0 = (rem / tiles_x) * TY, x0 = (rem % tiles_x) * TX;
        const float *xp = x + (long long)n * 16 * H * W;
        const float *gp = g + (long long)n * 16 * H * W;
        __syncthreads();
        for (int i = tid; i < 16 * (TY + 2) * (TX + 2); i += 256) {
            const int c_ = i / ((TY + 2) * (TX + 2)), r2 = i % ((TY + 2) * (TX + 2));
            const int r = r2 / (TX + 2), c = r2 % (TX + 2);
            const int gy = y0 - 1 + r, gx = x0 - 1 + c;
            xs[c_][r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? xp[((long long)c_ * H + gy) * W + gx] : 0.0f;
        }
        for (int i = tid; i < 16 * TY * TX; i += 256) {
            const int c_ = i / (TY * TX), r2 = i % (TY * TX);
            const int r = r2 / TX, c = r2 % TX;
            const int gy = y0 + r, gx = x0 + c;
            gs[c_][r][c] = (gy < H && gx < W) ? gp[((long long)c_ * H + gy) * W + gx] : 0.0f;
        }
        __syncthreads();
        for (int p = part; p < TY * TX; p += 4) {
            const int r = p / TX, c = p % TX;
            const float2 gv = make_float2(gs[co0][r][c], gs[co0 + 1][r][c]);
            bsum.x += gv.x; bsum.y += gv.y;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int k = 0; k < 9; ++k) acc[a][k] = ffma2s(gv, xs[ci0 + a][r + k / 3][c + k % 3], acc[a][k]);
        }
    }
    // fold the 4 pixel parts through shared memory, then one atomic per output per CTA
    __syncthreads();
    for (int i = tid; i < 2304 + 16; i += 256) red[i] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            atomicAdd(&red[((co0) * 16 + ci0 + a) * 9 + k], acc[a][k].x);
            atomicAdd(&red[((co0 + 1) * 16 + ci0 + a) * 9 + k], acc[a][k].y);
        }
    if (ci0 == 0) {
        atomicAdd(&red[2304 + co0], bsum.x);
        atomicAdd(&red[2304 + co0 + 1], bsum.y);
    }
    __syncthreads();
    for (int i = tid; i < 2304; i += 256) atomicAdd(gw + i, red[i]);
    if (gb && tid < 16) atomicAdd(gb + tid, red[2304 + tid]);
}

__global__ void __launch_bounds__(256) flow_warp_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ im,
                                                            const float *__restrict__ flow, const float *__restrict__ lin_x,
                                                            const float *__restrict__ lin_y, float *__restrict__ gim,
                                                            float *__restrict__ gflow, int N, int C, int H, int W, int flowN, float sign,
                                                            float sx, float sy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const long long plane = (long long)H * W;
    const int fn = n / (N / flowN);
    const float *fb = flow + (long long)fn * 2 * plane + (long long)y * W + x;
    const float fx = sign * fb[0], fy = sign * fb[plane];
    // forward coordinate arithmetic of flow_warp_kernel
    const float gxn = lin_x[x] + fx / sx, gyn = lin_y[y] + fy / sy;
    const float ixu = (gxn + 1.0f) * sx, iyu = (gyn + 1.0f) * sy;
    const float ix = fminf(fmaxf(ixu, 0.0f), (float)(W - 1)), iy = fminf(fmaxf(iyu, 0.0f), (float)(H - 1));
    const float mx = (ixu >= 0.0f && ixu <= (float)(W - 1)) ? 1.0f : 0.0f; // clip_coordinates_set_grad
    const float my = (iyu >= 0.0f && iyu <= (float)(H - 1)) ? 1.0f : 0.0f;
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wx = ix - x0f, wy = iy - y0f;
    const int x0i = (int)x0f, y0i = (int)y0f;
    const bool x1ok = x0i + 1 <= W - 1, y1ok = y0i + 1 <= H - 1;
    float gix = 0.0f, giy = 0.0f;
    for (int c = 0; c < C; ++c) {
        const long long base = ((long long)n * C + c) * plane;
        const float go = gout[base + (long long)y * W + x];
        const float *p = im + base + (long long)y0i * W + x0i;
        const float vnw = p[0], vne = x1ok ? p[1] : 0.0f, vsw = y1ok ? p[W] : 0.0f, vse = (x1ok && y1ok) ? p[W + 1] : 0.0f;
        if (gim) {
            float *q = gim + base + (long long)y0i * W + x0i;
            atomicAdd(q, go * (1.0f - wx) * (1.0f - wy));
            if (x1ok) atomicAdd(q + 1, go * wx * (1.0f - wy));
            if (y1ok) atomicAdd(q + W, go * (1.0f - wx) * wy);
            if (x1ok && y1ok) atomicAdd(q + W + 1, go * wx * wy);
        }
        gix += go * ((vne - vnw) * (1.0f - wy) + (vse - vsw) * wy);
        giy += go * ((vsw - vnw) * (1.0f - wx) + (vse - vne) * wx);
    }
    if (gflow) { // d ix / d fx = 1 inside the image, 0 where the coordinate was clipped
        float *gf = gflow + (long long)fn * 2 * plane + (long long)y * W + x;
        atomicAdd(gf, sign * gix * mx);
        atomicAdd(gf + plane, sign * giy * my);
    }
}

} // namespace train
} // namespace pmctf

using namespace pmctf::train;

static int launch_conv3x3(const float *x, const float *w, const float *b, float *y, const Epi &epi, int N, int cin, int cout, int H, int W,
                          void *stream)
{
    if (!x || !w || !y || N <= 0 || H <= 0 || W <= 0) return PMCTF_EINVAL;
    dim3 grid((W + TX - 1) / TX, (H + TY - 1) / TY, N), block(TX, TY);
    if (grid.y > 65535 || grid.z > 65535) return PMCTF_ESHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    if (cin == 1 && cout == 16) conv3x3_kernel<1, 16><<<grid, block, 0, st>>>(x, w, b, y, H, W, epi);
    else if (cin == 16 && cout == 16) {
        constexpr int SMEM = (144 * 16 + 16 * (TY + 2) * (TX2 + 2)) * 4;
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(conv3x3_16x16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
            if (e != cudaSuccess) return (int)e;
            configured = true;
        }
        conv3x3_16x16_kernel<<<dim3((W + TX2 - 1) / TX2, (H + TY - 1) / TY, N), 256, SMEM, st>>>(x, w, b, y, H, W, epi);
    }
    else if (cin == 16 && cout == 1) conv3x3_kernel<16, 1><<<grid, block, 0, st>>>(x, w, b, y, H, W, epi);
    else if (cin == 1 && cout == 1) conv3x3_kernel<1, 1><<<grid, block, 0, st>>>(x, w, b, y, H, W, epi);
    else return PMCTF_ESHAPE;
    return (int)cudaGetLastError();
}

extern "C" int pmctf_conv3x3(const float *x, const float *w, const float *b, float *y, int N, int cin, int cout, int H, int W, void *stream)
{
    const Epi none = {nullptr, nullptr, nullptr, EPI_NONE};
    return launch_conv3x3(x, w, b, y, none, N, cin, cout, H, W, stream);
}

extern "C" int pmctf_conv3x3_fused(const float *x, const float *w, const float *b, float *y, float *y2, const float *aux, const float *aux2, int mode,
                                   int N, int cin, int cout, int H, int W, void *stream)
{
    if (mode < EPI_NONE || mode > EPI_DTANH_ADD) return PMCTF_EINVAL;
    if ((mode == EPI_DUAL && !y2) || (mode >= EPI_ADD && !aux) || (mode == EPI_DTANH_ADD && !aux2)) return PMCTF_EINVAL;
    const Epi e = {aux, aux2, y2, mode};
    return launch_conv3x3(x, w, b, y, e, N, cin, cout, H, W, stream);
}

extern "C" int pmctf_conv3x3_wgrad(const float *x, const float *g, float *gw, float *gb, int N, int cin, int cout, int H, int W, void *stream)
{
    if (!x || !g || !gw || N <= 0 || H <= 0 || W <= 0) return PMCTF_EINVAL;
    const long long tiles = (long long)((W + TX - 1) / TX) * ((H + TY - 1) / TY) * N;
    if (tiles > 0x7fffffffLL) return PMCTF_ESHAPE;
    const unsigned grid = (unsigned)(tiles < 592 ? tiles : 592); // 4 CTAs per SM on 148 SMs
    cudaStream_t st = (cudaStream_t)stream;
    if (cin == 1 && cout == 16) wgrad3x3_kernel<1, 16><<<grid, 256, 0, st>>>(x, g, gw, gb, N, H, W);
    else if (cin == 16 && cout == 16) wgrad3x3_16x16_kernel<<<grid, 256, 0, st>>>(x, g, gw, gb, N, H, W);
    else if (cin == 16 && cout == 1) wgrad3x3_kernel<16, 1><<<grid, 256, 0, st>>>(x, g, gw, gb, N, H, W);
    else return PMCTF_ESHAPE;
    return (int)cudaGetLastError();
}

extern "C" int pmctf_flow_warp_bwd(const float *gout, const float *im, const float *flow, const float *lin_x, const float *lin_y, float *gim,
                                   float *gflow, int N, int C, int H, int W, int flowN, float sign, void *stream)
{
    if (!gout || !im || !flow || !lin_x || !lin_y || N <= 0 || C <= 0) return PMCTF_EINVAL;
    if (H < 2 || W < 2 || flowN < 1 || (N % flowN) || H > 65535 || N > 65535) return PMCTF_ESHAPE;
    dim3 grid((W + 255) / 256, H, N);
    flow_warp_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gout, im, flow, lin_x, lin_y, gim, gflow, N, C, H, W, flowN, sign,
                                                                (float)(((double)W - 1.0) / 2.0), (float)(((double)H - 1.0) / 2.0));
    return (int)cudaGetLastError();
}
