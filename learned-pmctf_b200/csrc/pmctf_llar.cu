// pmctf_llar.cu -- the autoregressive entropy-parameter model of the LL subband (pMCTF/layers/context_fusion.py:56-204,
// ContextFusionSubband as pWave++ builds it: 128 features, no context input) in its SEQUENTIAL form, the one the bitstream path
// needs (pWave.py:531-584: every coefficient's (scale, mean) depend on the coefficients coded before it).
//
// The reference walks the band in Python, ~25 small ATen calls per coefficient (forward_sequential, :160-204): seconds per
// plane on any GPU.  Here one kernel evaluates a coefficient's whole network -- the 1 -> 128 type-A masked convolution, two masked
// residual blocks, the type-B masked convolution and the three 1x1 layers, on the 3x3 windows of five channel-last history
// planes -- with one thread per output channel and the causal taps' weights streamed from L2 in a [tap][ci][co] layout
// (coalesced over the output channel).  The ENCODER knows every symbol in advance, so it runs the whole band in ONE launch (a
// loop over the coefficients inside the kernel, one CTA per plane of the batch) and emits int16 symbols + scale-table indexes;
// the DECODER launches once per coefficient, alternating with the host's rANS decode of that coefficient.  Both sides execute
// the same device function, so their parameters agree bit for bit (the full-plane convolutions of the training / rate-estimate
// path agree with it to fp32 rounding only, which is why the reference, too, uses the sequential form on both sides).
//
// One deliberate difference to the reference's encoder: the history it conditions on is the RECONSTRUCTED value
// round(symbol + mean), i.e. exactly what the decoder will have, instead of the original quantised coefficient
// (pWave.py:541-547 pads `symbols` = y and never writes ll_hat back).  The two differ only when y - mean ends in exactly .5; the
// reference's decoder drifts from its encoder in that case, this one cannot.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include <cooperative_groups.h>

#include "pmctf_b200.h"

namespace pmctf {
void count_launch();
namespace llar {
namespace cg = cooperative_groups;

constexpr int F = 128;          // features
constexpr int NT = 256;         // 32 groups of four output channels x 8 slices of the reduction dimension
constexpr int KSL = 8;
// causal taps in (dy, dx) relative to the current coefficient, padded-plane coordinates: type B = 5 taps, type A = the first 4
__constant__ int c_dy[5] = {-1, -1, -1, 0, 0};
__constant__ int c_dx[5] = {-1, 0, 1, -1, 0};

struct Net {
    const float *w_in;           // [4][128]   maskedConv1 (1 -> 128), causal taps only, [tap][co]
    const float *b_in;           // [128]
    const float *w[5];           // [5][128][128] = [tap][ci][co]: res0.conv1, res0.conv2, res1.conv1, res1.conv2, maskedConv2
    const float *b[5];
    const float *w1[2];          // [128][128] = [ci][co]: convs.0, convs.1
    const float *b1[2];
    const float *w_out;          // [2][128]: convs.2
    const float *b_out;          // [2]
};

struct Run {
    float *Y;                    // [B][H+2][W+2] reconstructed band, zero border, zero where not yet coded
    float *hist[5];              // [B][H+2][W+2][128] channel-last inputs of the five masked 128 -> 128 convolutions
    int B, H, W;
    float log_min, log_step, top;
    // encoder
    const float *yq;             // [B][H][W] quantised band, or null
    short *sym16, *idx16;        // [B][H*W]
    // decoder: one coefficient per launch
    int pos;                     // coefficient index h * W + w
    const float *prev;           // [B] reconstructed value of coefficient pos - 1 (written into Y first), host-mapped
    float *out_mean;             // [B] host-mapped
    short *out_idx;              // [B] host-mapped
};

struct Smem {
    float in[5 * F];             // staged inputs of a layer: [tap][ci] or [ci]
    float part[KSL * F];         // partial sums of the reduction slices
    float out[F];                // a layer's output
    float cur[F], first[F];      // running activation, maskedConv1's output (the outer skip, context_fusion.py:119)
};

__device__ __forceinline__ float lrelu(float v) { return v >= 0.0f ? v : v * 0.2f; }

// out[co] = b[co] + sum_k w[k][co] * in[k], K = 640 (five taps x 128 channels) or 128: thread = four adjacent output channels x one
// eighth of k (16-byte weight loads, eight in flight), partial sums folded in a fixed order -- the weights stream from L2, so the
// layer is as fast as the loads one SM can keep in flight
__device__ __forceinline__ void gemv(const float *__restrict__ w, const float *__restrict__ b, int K, Smem &s)
{
    const int cg = threadIdx.x & 31, ks = threadIdx.x >> 5, per = K / KSL;
    const float4 *wp = reinterpret_cast<const float4 *>(w) + (long long)(ks * per) * (F / 4) + cg;
    const float *ip = s.in + ks * per;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int j = 0; j < per; ++j) {
        const float4 wv = __ldg(wp + (long long)j * (F / 4));
        const float x = ip[j];
        acc.x = fmaf(wv.x, x, acc.x);
        acc.y = fmaf(wv.y, x, acc.y);
        acc.z = fmaf(wv.z, x, acc.z);
        acc.w = fmaf(wv.w, x, acc.w);
    }
    *reinterpret_cast<float4 *>(s.part + ks * F + 4 * cg) = acc;
    __syncthreads();
    if (threadIdx.x < F) {
        float v = b[threadIdx.x];
#pragma unroll
        for (int k = 0; k < KSL; ++k) v += s.part[k * F + threadIdx.x];
        s.out[threadIdx.x] = v;
    }
    __syncthreads();
}

// stage the five causal taps of a channel-last history plane around (h, w)
__device__ __forceinline__ void stage_taps(const float *__restrict__ hist, int h, int wd, int Wp, Smem &s)
{
    for (int i = threadIdx.x; i < 5 * F; i += NT) {
        const int t = i / F, ci = i - t * F;
        s.in[i] = hist[((long long)(h + 1 + c_dy[t]) * Wp + (wd + 1 + c_dx[t])) * F + ci];
    }
    __syncthreads();
}

// parameters of coefficient (h, w) of plane `bi`; every thread returns the same (scale, mean)
__device__ void coefficient(const Net &n, const Run &r, int bi, int h, int wd, Smem &s, float &scale, float &mean)
{
    const int tid = threadIdx.x, Hp = r.H + 2, Wp = r.W + 2;
    const float *Y = r.Y + (long long)bi * Hp * Wp;
    const long long plane = (long long)Hp * Wp * F, here = ((long long)(h + 1) * Wp + (wd + 1)) * F;
    if (tid < F) {   // maskedConv1 (type A): the four causal neighbours of the reconstructed band
        float t = n.b_in[tid];
#pragma unroll
        for (int k = 0; k < 4; ++k) t = fmaf(__ldg(n.w_in + k * F + tid), Y[(long long)(h + 1 + c_dy[k]) * Wp + (wd + 1 + c_dx[k])], t);
        s.cur[tid] = s.first[tid] = t;
    }
    __syncthreads();
    for (int k = 0; k < 2; ++k) {      // MaskResidual (context_fusion.py:30-41)
        float *h1 = r.hist[2 * k] + bi * plane, *h2 = r.hist[2 * k + 1] + bi * plane;
        if (tid < F) h1[here + tid] = s.cur[tid];
        __syncthreads();
        stage_taps(h1, h, wd, Wp, s);
        gemv(n.w[2 * k], n.b[2 * k], 5 * F, s);
        if (tid < F) h2[here + tid] = lrelu(s.out[tid]);
        __syncthreads();
        stage_taps(h2, h, wd, Wp, s);
        gemv(n.w[2 * k + 1], n.b[2 * k + 1], 5 * F, s);
        if (tid < F) s.cur[tid] = s.out[tid] + s.cur[tid];
        __syncthreads();
    }
    float *h5 = r.hist[4] + bi * plane;
    if (tid < F) h5[here + tid] = s.cur[tid] + s.first[tid];
    __syncthreads();
    stage_taps(h5, h, wd, Wp, s);
    gemv(n.w[4], n.b[4], 5 * F, s);
    for (int k = 0; k < 2; ++k) {      // convs.0, convs.1 (1x1) behind LeakyReLU
        if (tid < F) s.in[tid] = lrelu(s.out[tid]);
        __syncthreads();
        gemv(n.w1[k], n.b1[k], F, s);
    }
    // convs.2: 128 -> 2 on lrelu(out), a fixed-order tree over the channels (identical on both sides)
    if (tid < F) {
        const float t = lrelu(s.out[tid]);
        s.part[tid] = t * __ldg(n.w_out + tid);
        s.part[F + tid] = t * __ldg(n.w_out + F + tid);
    }
    __syncthreads();
    for (int st = F / 2; st > 0; st >>= 1) {
        if (tid < st) {
            s.part[tid] += s.part[tid + st];
            s.part[F + tid] += s.part[F + tid + st];
        }
        __syncthreads();
    }
    scale = s.part[0] + __ldg(n.b_out);
    mean = s.part[F] + __ldg(n.b_out + 1);
    __syncthreads();
}

__device__ __forceinline__ short table_index(float scale, const Run &r)
{
    float v = (logf(fmaxf(scale, 1e-5f)) - r.log_min) / r.log_step;
    v = fminf(fmaxf(v, 0.0f), r.top);
    return (short)(int)v;
}

// encoder: the whole band of plane blockIdx.x in raster order
__global__ void __launch_bounds__(NT) llar_encode_kernel(const Net n, const Run r)
{
    __shared__ Smem s;
    const int bi = blockIdx.x, Wp = r.W + 2;
    float *Y = r.Y + (long long)bi * (r.H + 2) * Wp;
    for (int h = 0; h < r.H; ++h)
        for (int w = 0; w < r.W; ++w) {
            float scale, mean;
            coefficient(n, r, bi, h, w, s, scale, mean);
            if (threadIdx.x == 0) {
                const long long p = (long long)bi * r.H * r.W + (long long)h * r.W + w;
                const float sym = rintf(rintf(r.yq[p]) - mean);          // pWave.py:549-553: round(round(y) - mean)
                Y[(long long)(h + 1) * Wp + (w + 1)] = rintf(sym + mean);  // what the decoder will reconstruct
                r.sym16[p] = (short)(int)fminf(fmaxf(sym, -30000.0f), 30000.0f);
                r.idx16[p] = table_index(scale, r);
            }
            __syncthreads();
        }
}

// decoder: one coefficient; the previous coefficient's reconstruction arrives from the host
__global__ void __launch_bounds__(NT) llar_decode_kernel(const Net n, const Run r)
{
    __shared__ Smem s;
    const int bi = blockIdx.x, Wp = r.W + 2;
    float *Y = r.Y + (long long)bi * (r.H + 2) * Wp;
    if (r.pos > 0 && threadIdx.x == 0) {
        const int q = r.pos - 1;
        Y[(long long)(q / r.W + 1) * Wp + (q % r.W + 1)] = r.prev[bi];
    }
    __syncthreads();
    float scale, mean;
    coefficient(n, r, bi, r.pos / r.W, r.pos % r.W, s, scale, mean);
    if (threadIdx.x == 0) {
        r.out_mean[bi] = mean;
        r.out_idx[bi] = table_index(scale, r);
        __threadfence_system();
    }
}

// ---- the same network on ALL coefficients of a band at once ----------------------------------------------------------------
// When every coefficient is known beforehand (the encoder; the rate-estimate path), nothing in the model is sequential except
// the history it conditions on.  The kernels below evaluate the network layer by layer over the whole band with EXACTLY the
// arithmetic of coefficient() -- the same eight reduction slices per output, each one fma chain in the same order, folded onto
// the bias in the same order, the same multiply-then-tree for the closing 128 -> 2 layer -- so their (scale, mean) equal the
// sequential form's bit for bit (tests/test_gpu_llar.py), at a few hundred microseconds per band instead of 35 us per coefficient.
// The encoder's history is the band as the decoder will reconstruct it, round(symbol + mean), which is only known once the mean
// is: the parallel pass SPECULATES that it equals round(y) (true unless round(y) - mean ends in exactly .5), checks every
// coefficient, and reports a mismatch through `mismatch` so that the caller can fall back to the sequential encoder.
constexpr int TP = 32;          // coefficients per CTA: thread = four output channels x coefficients cs, cs + 8, cs + 16, cs + 24
constexpr int IN_P = 84;        // row pitch of the staged inputs (a slice holds 80 or 16 values per coefficient)

struct TileSmem {
    float in[TP][IN_P];
    float out[TP][F];
    float red[TP][F];
    float sc[TP], mn[TP];
};

// v[i] = b + slice_0 + ... + slice_7 for the tile's coefficients; stage(k0, n) fills s.in[c][0..n) with inputs k0..k0+n-1 of every
// coefficient c of the tile
template <class Stage>
__device__ __forceinline__ void tile_gemv(const float *__restrict__ w, const float *__restrict__ b, int K, TileSmem &s, float4 (&v)[4], Stage stage)
{
    const int cg = threadIdx.x & 31, cs = threadIdx.x >> 5, per = K / KSL;
    const float4 bv = __ldg(reinterpret_cast<const float4 *>(b) + cg);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = bv;
    for (int sl = 0; sl < KSL; ++sl) {
        __syncthreads();                // the previous slice's inputs have been consumed
        stage(sl * per, per);
        __syncthreads();
        const float4 *wp = reinterpret_cast<const float4 *>(w) + (long long)(sl * per) * (F / 4) + cg;
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int j = 0; j < per; ++j) {
            const float4 wv = __ldg(wp + (long long)j * (F / 4));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x = s.in[cs + 8 * i][j];
                acc[i].x = fmaf(wv.x, x, acc[i].x);
                acc[i].y = fmaf(wv.y, x, acc[i].y);
                acc[i].z = fmaf(wv.z, x, acc[i].z);
                acc[i].w = fmaf(wv.w, x, acc[i].w);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i].x += acc[i].x;
            v[i].y += acc[i].y;
            v[i].z += acc[i].z;
            v[i].w += acc[i].w;
        }
    }
}

// band -> padded plane of the history (round_in: the encoder's round(y)); maskedConv1 (type A) -> hist[0]
__global__ void __launch_bounds__(256) llar_par_fill_kernel(const Run r, const float *__restrict__ x, int round_in)
{
    const long long total = (long long)r.B * r.H * r.W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int bi = (int)(i / ((long long)r.H * r.W)), p = (int)(i - (long long)bi * r.H * r.W);
        const int h = p / r.W, w = p - h * r.W;
        const float v = x[i];
        r.Y[((long long)bi * (r.H + 2) + h + 1) * (r.W + 2) + w + 1] = round_in ? rintf(v) : v;
    }
}

__global__ void __launch_bounds__(F) llar_par_in_kernel(const Net n, const Run r)
{
    const int bi = blockIdx.y, tid = threadIdx.x, Hp = r.H + 2, Wp = r.W + 2, HW = r.H * r.W;
    const float *Y = r.Y + (long long)bi * Hp * Wp;
    float *h0 = r.hist[0] + (long long)bi * Hp * Wp * F;
    float wk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wk[k] = __ldg(n.w_in + k * F + tid);
    const float bb = n.b_in[tid];
    for (int p = blockIdx.x * TP; p < min(HW, (int)(blockIdx.x + 1) * TP); ++p) {
        const int h = p / r.W, wd = p - h * r.W;
        float t = bb;
#pragma unroll
        for (int k = 0; k < 4; ++k) t = fmaf(wk[k], Y[(long long)(h + 1 + c_dy[k]) * Wp + (wd + 1 + c_dx[k])], t);
        h0[((long long)(h + 1) * Wp + (wd + 1)) * F + tid] = t;
    }
}

// one masked 128 -> 128 layer on a tile of TP coefficients.  MODE 1: dst = lrelu(v); 2: dst = v + add1; 3: dst = (v + add1) + add2
// (the positions' own entries of earlier history planes); 4: maskedConv2 and everything behind it
struct ParOut {
    short *sym16, *idx16;
    float *scales, *means;
    int *mismatch;
};

template <int MODE>
__global__ void __launch_bounds__(NT) llar_par_layer_kernel(const Net n, const Run r, int li, int src, int dst, int add1, int add2, const ParOut o)
{
    __shared__ TileSmem s;
    const int bi = blockIdx.y, tid = threadIdx.x, Hp = r.H + 2, Wp = r.W + 2, HW = r.H * r.W, p0 = blockIdx.x * TP;
    const long long plane = (long long)Hp * Wp * F;
    const float *hs = r.hist[src] + bi * plane;
    float4 v[4];
    tile_gemv(n.w[li], n.b[li], 5 * F, s, v, [&](int k0, int per) {
        for (int i = tid; i < TP * per; i += NT) {
            const int c = i / per, j = i - c * per, k = k0 + j, t = k >> 7, ci = k & (F - 1), p = p0 + c;
            float x = 0.0f;
            if (p < HW) {
                const int h = p / r.W, wd = p - h * r.W;
                x = hs[((long long)(h + 1 + c_dy[t]) * Wp + (wd + 1 + c_dx[t])) * F + ci];
            }
            s.in[c][j] = x;
        }
    });
    const int cg = tid & 31, cs = tid >> 5;
    if constexpr (MODE != 4) {
        float *hd = r.hist[dst] + bi * plane;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int p = p0 + cs + 8 * i;
            if (p >= HW) continue;
            const int h = p / r.W, wd = p - h * r.W;
            const long long here = ((long long)(h + 1) * Wp + (wd + 1)) * F + 4 * cg;
            float4 t = v[i];
            if (MODE == 1) {
                t = make_float4(lrelu(t.x), lrelu(t.y), lrelu(t.z), lrelu(t.w));
            } else {
                const float4 a = *reinterpret_cast<const float4 *>(r.hist[add1] + bi * plane + here);
                t = make_float4(t.x + a.x, t.y + a.y, t.z + a.z, t.w + a.w);
                if (MODE == 3) {
                    const float4 f = *reinterpret_cast<const float4 *>(r.hist[add2] + bi * plane + here);
                    t = make_float4(t.x + f.x, t.y + f.y, t.z + f.z, t.w + f.w);
                }
            }
            *reinterpret_cast<float4 *>(hd + here) = t;
        }
    } else {
    // the 1x1 layers behind LeakyReLU, per coefficient
    for (int k = 0; k < 2; ++k) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4 *>(&s.out[cs + 8 * i][4 * cg]) = v[i];
        tile_gemv(n.w1[k], n.b1[k], F, s, v, [&](int k0, int per) {
            for (int i = tid; i < TP * per; i += NT) {
                const int c = i / per, j = i - c * per;
                s.in[c][j] = lrelu(s.out[c][k0 + j]);
            }
        });
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4 *>(&s.out[cs + 8 * i][4 * cg]) = v[i];
    __syncthreads();
    // convs.2: multiply, then the fixed-order tree of coefficient()
    for (int oi = 0; oi < 2; ++oi) {
        for (int i = tid; i < TP * F; i += NT) {
            const int c = i >> 7, ch = i & (F - 1);
            s.red[c][ch] = lrelu(s.out[c][ch]) * __ldg(n.w_out + oi * F + ch);
        }
        __syncthreads();
        for (int st = F / 2; st > 0; st >>= 1) {
            for (int i = tid; i < TP * st; i += NT) {
                const int c = i / st, ch = i - c * st;
                s.red[c][ch] += s.red[c][ch + st];
            }
            __syncthreads();
        }
        if (tid < TP) (oi == 0 ? s.sc : s.mn)[tid] = s.red[tid][0] + __ldg(n.b_out + oi);
        __syncthreads();
    }
    if (tid < TP && p0 + tid < HW) {
        const long long p = (long long)bi * HW + p0 + tid;
        const float scale = s.sc[tid], mean = s.mn[tid];
        if (o.scales) o.scales[p] = scale;
        if (o.means) o.means[p] = mean;
        if (o.sym16) {
            const float yr = rintf(r.yq[p]);
            const float sym = rintf(yr - mean);
            if (rintf(sym + mean) != yr && o.mismatch) atomicOr(o.mismatch, 1);   // the speculated history is not what the decoder will see
            o.sym16[p] = (short)(int)fminf(fmaxf(sym, -30000.0f), 30000.0f);
            o.idx16[p] = table_index(scale, r);
        }
    }
    }
}

// ---- the DECODER of a whole band in one launch -------------------------------------------------------------------------------
// The decoder is sequential by definition (a coefficient's parameters need the coefficients before it), so what can be removed
// is everything around the arithmetic: (1) the host round trip per coefficient -- the rANS decoder of the band's symbols runs on
// the device, inside the kernel; (2) the weight traffic -- coefficient() streams 1.8 MB of weights from L2 through ONE SM per
// coefficient (>= 14 us at the L2 -> SM rate).  Here a CLUSTER of eight CTAs works on one plane: CTA r owns reduction slice r of
// every masked 128 -> 128 layer (the contract's eight slices: 80 inputs x 128 outputs = 40 KB per layer, 200 KB for the five
// layers, RESIDENT in its shared memory for the whole band), computes that slice's partial sums (one fma chain per output, the
// same chain as gemv()) and stores them into EVERY CTA's shared memory through the cluster's distributed shared memory; after
// one cluster barrier every CTA folds the eight partials onto the bias in the contract's order and applies the layer's epilogue
// itself -- eight identical copies of the 128 values (the (0, 0) tap of the next layer, the (0, -1) tap of the next coefficient),
// nothing to broadcast, ONE barrier per layer (seven per coefficient, an eighth behind the rANS step).  CTA 0 also writes the
// values to the global history planes: the taps of the next row, which every CTA prefetches one coefficient ahead into a
// four-column ring.  The first version (CTA 0 folds and broadcasts: two barriers per layer) took 99 ms per 1080p band.
// Same values as coefficient() bit for bit (tests/test_gpu_llar.py).
constexpr int CL = 8;                                   // cluster size = reduction slices of the contract
constexpr int SL = 5 * F / KSL;                         // 80 inputs per slice of a masked layer
constexpr int C_W = 0, C_UP = C_W + 5 * SL * F, C_CP = C_UP + 5 * 4 * F, C_BC = C_CP + 2 * 5 * F, C_PART = C_BC + 3 * F,
              C_IN = C_PART + 2 * CL * F, C_RED = C_IN + F, C_MISC = C_RED + 2 * F, C_FLOATS = C_MISC + 16;
constexpr int M_YLEFT = 4;                              // misc[0..3]: ring of the reconstructed row above, misc[4]: the left neighbour
static_assert(C_FLOATS * 4 <= 227 * 1024, "cluster decoder: shared memory");

struct DevRans {
    const unsigned int *words;      // the sub-stream's 32-bit words (device copy)
    long long nwords;
    unsigned long long *state;      // [0] rANS state x, [1] index of the next word, [2] error flag, [3] turn counter (planes of a batch)
    const int *cdf, *sizes, *offsets;
    int num, stride;
};

// one symbol of table `ti` (decode_part of pmctf_rans.cu / rans.cpp:279-331 on the device; the tables are non-decreasing, so the
// reference's linear scan is an upper-bound search)
__device__ int rans_decode_one(const DevRans &d, int ti, unsigned long long &x, unsigned long long &pos, bool &bad)
{
    auto next = [&]() -> unsigned long long {
        if ((long long)pos >= d.nwords) { bad = true; return 0ull; }
        return (unsigned long long)d.words[pos++];
    };
    auto get_raw = [&]() -> unsigned int {
        const unsigned int v = (unsigned int)(x & 15ull);
        x >>= 4;
        if (x < (1ull << 31)) x = (x << 32) | next();
        return v;
    };
    if (ti < 0 || ti >= d.num) { bad = true; return 0; }
    const int *cdf = d.cdf + (long long)ti * d.stride;
    const int size = d.sizes[ti], escape = size - 2;
    if (escape < 0 || size > d.stride) { bad = true; return 0; }
    const unsigned int target = (unsigned int)(x & 0xFFFFull);
    int lo = 0, hi = size - 1;          // s = number of entries of cdf[1 .. size-1] that are <= target
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((unsigned int)cdf[mid] <= target) lo = mid; else hi = mid - 1;
    }
    const int sI = lo;
    if (sI + 1 >= size) { bad = true; return 0; }
    const unsigned int start = (unsigned int)cdf[sI], freq = (unsigned int)(cdf[sI + 1] - cdf[sI]);
    x = (unsigned long long)freq * (x >> 16) + (x & 0xFFFFull) - start;
    if (x < (1ull << 31)) x = (x << 32) | next();
    int v = sI;
    if (v == escape) {
        unsigned int dg = get_raw();
        int digits = (int)dg;
        while (dg == 15u) {
            dg = get_raw();
            digits += (int)dg;
            if (digits > 64) { bad = true; return 0; }
        }
        unsigned int raw = 0;
        for (int j = 0; j < digits; ++j) raw |= get_raw() << (j * 4);
        v = (int)(raw >> 1);
        v = (raw & 1u) ? -v - 1 : v + escape;
    }
    return (int)(short)(v + d.offsets[ti]);
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(F) llar_cluster_decode_kernel(const Net n, const Run r, const DevRans d, float *__restrict__ out)
{
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), bi = blockIdx.y, tid = threadIdx.x;
    float *wres = sm + C_W, *up = sm + C_UP, *cp = sm + C_CP, *bc = sm + C_BC, *part = sm + C_PART, *in = sm + C_IN, *red = sm + C_RED,
          *misc = sm + C_MISC;
    const int Hp = r.H + 2, Wp = r.W + 2, HW = r.H * r.W;
    const long long plane = (long long)Hp * Wp * F;
    float *Y = r.Y + (long long)bi * Hp * Wp;
    float *hist[5];
#pragma unroll
    for (int L = 0; L < 5; ++L) hist[L] = r.hist[L] + bi * plane;
    // this CTA's slice of the masked layers: resident for the whole band; of the two 1x1 layers: 16 inputs per output, in registers
    for (int L = 0; L < 5; ++L)
        for (int j = 0; j < SL; ++j) wres[(L * SL + j) * F + tid] = __ldg(n.w[L] + (long long)(SL * rank + j) * F + tid);
    float w1a[16], w1b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        w1a[j] = __ldg(n.w1[0] + (16 * rank + j) * F + tid);
        w1b[j] = __ldg(n.w1[1] + (16 * rank + j) * F + tid);
    }
    float win[4], bl[5];
#pragma unroll
    for (int k = 0; k < 4; ++k) win[k] = __ldg(n.w_in + k * F + tid);
#pragma unroll
    for (int L = 0; L < 5; ++L) bl[L] = __ldg(n.b[L] + tid);
    const float bin = __ldg(n.b_in + tid), b1a = __ldg(n.b1[0] + tid), b1b = __ldg(n.b1[1] + tid);
    const float wo0 = __ldg(n.w_out + tid), wo1 = __ldg(n.w_out + F + tid);
    float *rpart[CL];       // every CTA receives every slice's partial sums and folds them itself (same order, same values):
#pragma unroll              // one cluster barrier per layer instead of two, nothing to broadcast
    for (int q = 0; q < CL; ++q) rpart[q] = cluster.map_shared_rank(part, q);
    for (int i = tid; i < C_FLOATS - C_UP; i += F) sm[C_UP + i] = 0.0f;
    cluster.sync();
    bool bad = false;
    unsigned long long rx = 0, rpos = 0;    // a single plane keeps the reader in the registers of the decoding thread
    if (rank == 0 && tid == 0 && r.B == 1) {
        rx = d.state[0];
        rpos = d.state[1];
    }
    int pp = 0;             // the partial sums alternate between two buffers: a CTA may already send a layer's while another still folds the one before
    // one exchange: this CTA's partial to all, barrier, fold onto the bias in slice order
    auto exchange = [&](float acc, float bias) -> float {
#pragma unroll
        for (int q = 0; q < CL; ++q) rpart[q][(pp * CL + rank) * F + tid] = acc;
        cluster.sync();
        float v = bias;
#pragma unroll
        for (int q = 0; q < CL; ++q) v += part[(pp * CL + q) * F + tid];
        pp ^= 1;
        return v;
    };
    for (int h = 0; h < r.H; ++h) {
        // row start: the left neighbour is the zero border; the rings of the row above hold columns -1, 0, 1 (padded row h = image row h - 1)
#pragma unroll
        for (int L = 0; L < 5; ++L) {
            cp[(1 * 5 + L) * F + tid] = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) up[(L * 4 + c) * F + tid] = __ldcg(hist[L] + ((long long)h * Wp + c) * F + tid);
        }
        if (tid < 3) misc[tid] = __ldcg(Y + (long long)h * Wp + tid);
        if (tid == 3) misc[M_YLEFT] = 0.0f;
        __syncthreads();
        for (int w = 0; w < r.W; ++w) {
            const int ci = w & 1, pi = ci ^ 1;
            const long long here = ((long long)(h + 1) * Wp + (w + 1)) * F;
            float pf[5], pfy = 0.0f;
            const bool pre = w + 2 <= r.W;      // column w + 2 of the row above: needed by the next coefficient
#pragma unroll
            for (int L = 0; L < 5; ++L) pf[L] = pre ? __ldcg(hist[L] + ((long long)h * Wp + (w + 3)) * F + tid) : 0.0f;
            if (pre && tid == 0) pfy = __ldcg(Y + (long long)h * Wp + (w + 3));
            {                                   // maskedConv1 (type A) on the reconstructed band
                float t = bin;
                t = fmaf(win[0], misc[w & 3], t);
                t = fmaf(win[1], misc[(w + 1) & 3], t);
                t = fmaf(win[2], misc[(w + 2) & 3], t);
                t = fmaf(win[3], misc[M_YLEFT], t);
                cp[(ci * 5 + 0) * F + tid] = t;
                if (rank == 0) hist[0][here + tid] = t;
            }
            __syncthreads();
            for (int L = 0; L < 5; ++L) {
                if (tid < SL) {
                    const int k = SL * rank + tid, t = k >> 7, c = k & (F - 1);
                    float x;
                    if (t < 3) x = up[(L * 4 + ((w + t) & 3)) * F + c];
                    else if (t == 3) x = cp[(pi * 5 + L) * F + c];
                    else x = cp[(ci * 5 + L) * F + c];
                    in[tid] = x;
                }
                __syncthreads();
                float acc = 0.0f;
#pragma unroll 8
                for (int j = 0; j < SL; ++j) acc = fmaf(wres[(L * SL + j) * F + tid], in[j], acc);
                const float v = exchange(acc, bl[L]);
                float o;
                if (L == 0 || L == 2) o = lrelu(v);
                else if (L == 1) o = v + cp[(ci * 5 + 0) * F + tid];
                else if (L == 3) o = (v + cp[(ci * 5 + 2) * F + tid]) + cp[(ci * 5 + 0) * F + tid];
                else o = v;
                if (L < 4) {
                    cp[(ci * 5 + L + 1) * F + tid] = o;
                    if (rank == 0) hist[L + 1][here + tid] = o;
                } else {
                    bc[tid] = o;
                }
                __syncthreads();
            }
            for (int k = 0; k < 2; ++k) {       // convs.0, convs.1 behind LeakyReLU
                if (tid < 16) in[tid] = lrelu(bc[k * F + 16 * rank + tid]);
                __syncthreads();
                float acc = 0.0f;
#pragma unroll
                for (int j = 0; j < 16; ++j) acc = fmaf(k == 0 ? w1a[j] : w1b[j], in[j], acc);
                bc[(k + 1) * F + tid] = exchange(acc, k == 0 ? b1a : b1b);
                __syncthreads();
            }
            if (rank == 0) {                    // convs.2 (multiply, fixed-order tree), the symbol, the reconstruction
                const float t = lrelu(bc[2 * F + tid]);
                red[tid] = t * wo0;
                red[F + tid] = t * wo1;
                __syncthreads();
                for (int st = F / 2; st > 0; st >>= 1) {
                    if (tid < st) {
                        red[tid] += red[tid + st];
                        red[F + tid] += red[F + tid + st];
                    }
                    __syncthreads();
                }
                if (tid == 0) {
                    const float scale = red[0] + __ldg(n.b_out), mean = red[F] + __ldg(n.b_out + 1);
                    const int ti = table_index(scale, r);
                    volatile unsigned long long *stt = d.state;
                    const unsigned long long turn = (unsigned long long)(h * r.W + w) * (unsigned long long)r.B + (unsigned long long)bi;
                    if (r.B > 1) {              // the planes of a batch share the stream: plane order inside a coefficient
                        while (stt[3] != turn) { }
                        __threadfence();
                        rx = stt[0];
                        rpos = stt[1];
                    }
                    const int sym = rans_decode_one(d, ti, rx, rpos, bad);
                    if (r.B > 1) {
                        stt[0] = rx;
                        stt[1] = rpos;
                        __threadfence();
                        stt[3] = turn + 1ull;
                    }
                    const float rec = rintf((float)sym + mean);
                    Y[(long long)(h + 1) * Wp + (w + 1)] = rec;
                    out[(long long)bi * HW + h * r.W + w] = rec;
#pragma unroll
                    for (int q = 0; q < CL; ++q) cluster.map_shared_rank(misc, q)[M_YLEFT] = rec;   // the next coefficient's left neighbour
                }
            }
            if (pre) {
#pragma unroll
                for (int L = 0; L < 5; ++L) up[(L * 4 + ((w + 3) & 3)) * F + tid] = pf[L];
                if (tid == 0) misc[(w + 3) & 3] = pfy;
            }
            cluster.sync();                     // the reconstruction has reached every CTA
        }
    }
    if (rank == 0 && tid == 0) {
        if (r.B == 1) {
            d.state[0] = rx;
            d.state[1] = rpos;
        }
        if (bad) d.state[2] = 1ull;
    }
}

// OIHW masked weights -> the causal-tap layouts above.  src [128][cin][3][3]; dst [taps][cin][128]
__global__ void llar_pack_kernel(const float *__restrict__ w, int cin, int taps, float *__restrict__ out)
{
    const int total = taps * cin * F;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % F, ci = (i / F) % cin, t = i / (F * cin);
        const int ky = c_dy[t] + 1, kx = c_dx[t] + 1;
        out[i] = w[(((long long)co * cin + ci) * 3 + ky) * 3 + kx];
    }
}
// [co][ci] (1x1, OIHW with 1x1 kernel) -> [ci][co]
__global__ void llar_transpose_kernel(const float *__restrict__ w, float *__restrict__ out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F * F; i += gridDim.x * blockDim.x) out[(i % F) * F + i / F] = w[i];
}

} // namespace llar
} // namespace pmctf

using namespace pmctf;

extern "C" {

int pmctf_llar_pack(const float *w, int cin, int taps, float *out, void *stream)
{
    if (!w || !out || !((cin == 1 && taps == 4) || (cin == llar::F && taps == 5) || (cin == llar::F && taps == 0))) return PMCTF_EINVAL;
    if (taps == 0)
        llar::llar_transpose_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(w, out);
    else
        llar::llar_pack_kernel<<<(taps * cin * llar::F + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, cin, taps, out);
    count_launch();
    return (int)cudaGetLastError();
}

static int fill(const pmctf_llar_t *p, llar::Net &n, llar::Run &r)
{
    if (!p || !p->w_in || !p->b_in || !p->w_out || !p->b_out || !p->Y || p->B <= 0 || p->H <= 0 || p->W <= 0) return PMCTF_EINVAL;
    n.w_in = p->w_in; n.b_in = p->b_in; n.w_out = p->w_out; n.b_out = p->b_out;
    for (int i = 0; i < 5; ++i) {
        if (!p->w[i] || !p->b[i] || !p->hist[i]) return PMCTF_EINVAL;
        n.w[i] = p->w[i]; n.b[i] = p->b[i]; r.hist[i] = p->hist[i];
    }
    for (int i = 0; i < 2; ++i) {
        if (!p->w1[i] || !p->b1[i]) return PMCTF_EINVAL;
        n.w1[i] = p->w1[i]; n.b1[i] = p->b1[i];
    }
    if (!(p->log_scale_step > 0.0f) || p->scale_levels < 1 || p->scale_levels > 32767) return PMCTF_EINVAL;
    r.Y = p->Y; r.B = p->B; r.H = p->H; r.W = p->W; r.log_min = p->log_scale_min; r.log_step = p->log_scale_step; r.top = (float)(p->scale_levels - 1);
    r.yq = nullptr; r.sym16 = nullptr; r.idx16 = nullptr; r.pos = 0; r.prev = nullptr; r.out_mean = nullptr; r.out_idx = nullptr;
    return 0;
}

int pmctf_llar_encode(const pmctf_llar_t *p, const float *yq, short *sym16, short *idx16, void *stream)
{
    llar::Net n;
    llar::Run r;
    int e = fill(p, n, r);
    if (e) return e;
    if (!yq || !sym16 || !idx16) return PMCTF_EINVAL;
    r.yq = yq; r.sym16 = sym16; r.idx16 = idx16;
    llar::llar_encode_kernel<<<p->B, llar::NT, 0, (cudaStream_t)stream>>>(n, r);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_llar_forward(const pmctf_llar_t *p, const float *x, int round_in, short *sym16, short *idx16, float *scales, float *means,
                       int *mismatch, void *stream)
{
    llar::Net n;
    llar::Run r;
    int e = fill(p, n, r);
    if (e) return e;
    if (!x || ((sym16 == nullptr) != (idx16 == nullptr)) || (!sym16 && !scales && !means)) return PMCTF_EINVAL;
    const long long tiles = ((long long)p->H * p->W + llar::TP - 1) / llar::TP;
    if (tiles > 0x7fffffffLL || p->B > 65535) return PMCTF_EINVAL;
    r.yq = x;
    const llar::ParOut o{sym16, idx16, scales, means, mismatch};
    const llar::ParOut none{nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((unsigned)tiles, (unsigned)p->B);
    const long long total = (long long)p->B * p->H * p->W;
    llar::llar_par_fill_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 4096), 256, 0, st>>>(r, x, round_in);
    llar::llar_par_in_kernel<<<grid, llar::F, 0, st>>>(n, r);
    llar::llar_par_layer_kernel<1><<<grid, llar::NT, 0, st>>>(n, r, 0, 0, 1, 0, 0, none);   // res0.conv1:  hist1 = lrelu(.)
    llar::llar_par_layer_kernel<2><<<grid, llar::NT, 0, st>>>(n, r, 1, 1, 2, 0, 0, none);   // res0.conv2:  hist2 = . + hist0
    llar::llar_par_layer_kernel<1><<<grid, llar::NT, 0, st>>>(n, r, 2, 2, 3, 0, 0, none);   // res1.conv1:  hist3 = lrelu(.)
    llar::llar_par_layer_kernel<3><<<grid, llar::NT, 0, st>>>(n, r, 3, 3, 4, 2, 0, none);   // res1.conv2:  hist4 = (. + hist2) + hist0
    llar::llar_par_layer_kernel<4><<<grid, llar::NT, 0, st>>>(n, r, 4, 4, 0, 0, 0, o);      // maskedConv2, convs.0-2, symbols
    for (int i = 0; i < 7; ++i) count_launch();
    return (int)cudaGetLastError();
}

int pmctf_llar_decode_band(const pmctf_llar_t *p, const unsigned int *words, long long nwords, unsigned long long *state, const int *cdfs,
                           int cdf_num, int cdf_stride, const int *cdfs_sizes, const int *offsets, float *out, void *stream)
{
    llar::Net n;
    llar::Run r;
    int e = fill(p, n, r);
    if (e) return e;
    if (!words || nwords < 0 || !state || !cdfs || !cdfs_sizes || !offsets || cdf_num <= 0 || cdf_stride < 2 || !out) return PMCTF_EINVAL;
    if (p->B > 16 || (long long)p->H * p->W > 0x7fffffffLL) return PMCTF_ESHAPE;   // every plane's cluster must be resident: they take turns on the stream
    const llar::DevRans d{words, nwords, state, cdfs, cdfs_sizes, offsets, cdf_num, cdf_stride};
    constexpr int SMEM = llar::C_FLOATS * 4;
    cudaError_t ce = cudaFuncSetAttribute(llar::llar_cluster_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (ce != cudaSuccess) return (int)ce;
    llar::llar_cluster_decode_kernel<<<dim3(llar::CL, (unsigned)p->B), llar::F, SMEM, (cudaStream_t)stream>>>(n, r, d, out);
    count_launch();
    return (int)cudaGetLastError();
}

int pmctf_llar_decode_step(const pmctf_llar_t *p, int pos, const float *prev, float *out_mean, short *out_idx, void *stream)
{
    llar::Net n;
    llar::Run r;
    int e = fill(p, n, r);
    if (e) return e;
    if (pos < 0 || pos >= p->H * p->W || !out_mean || !out_idx || (pos > 0 && !prev)) return PMCTF_EINVAL;
    r.pos = pos; r.prev = prev; r.out_mean = out_mean; r.out_idx = out_idx;
    llar::llar_decode_kernel<<<p->B, llar::NT, 0, (cudaStream_t)stream>>>(n, r);
    count_launch();
    return (int)cudaGetLastError();
}

} // extern "C"
