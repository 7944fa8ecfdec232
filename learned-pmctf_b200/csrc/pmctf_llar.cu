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

#include "pmctf_b200.h"

namespace pmctf {
void count_launch();
namespace llar {

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
