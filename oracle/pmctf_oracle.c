/*
 * pmctf_oracle.c -- CPU restatement of the reference's MCTF + pWave++ lifting hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (learned-pmctf_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: the reference (FAU-LMS/Learned-pMCTF) ships NO golden vectors or tests
 * (SURVEY.md section 4).  The oracle is therefore pinned against outputs of the reference
 * itself, generated in the build container by oracle/make_golden.py (imports the
 * unmodified reference from /root/reference, CPU fp32) and committed under tests/golden/.
 *
 * What is restated (all citations relative to /root/reference):
 *   flow_warp / torch_warp             pMCTF/layers/video/video_net.py:32-55   (+ ATen CPU grid_sampler_2d)
 *   bilineardownsacling(mv)/2          pMCTF/layers/video/video_net.py:66-71   (+ ATen CPU upsample_bilinear2d)
 *   PredictUpdate                      pMCTF/layers/lifting_1d.py:25-49
 *   TemporalLifting.predict/update     pMCTF/layers/video/wavelet_transform_temporal_mctf.py:27-45
 *   pMCTF.forward_MCTF / inverse_MCTF  pMCTF/models/video/pMCTF_L.py:297-330
 *   split / merge                      pMCTF/layers/lifting_1d.py:10-22
 *   iWave1D.forward_lift/backward_lift pMCTF/layers/lifting_1d.py:103-189
 *   LiftingScheme2D fwd/bwd            pMCTF/layers/wavelet_transform.py:25-57
 *   quantise / dequantise              pMCTF/models/pWave.py:184-202, layers/layers.py:71-92
 *
 * Arithmetic contract (shared, as a SPECIFICATION, with the CUDA kernels -- the code is
 * written independently on each side):
 *   - every convolution output is one sequential fp32 FMA chain: acc = bias, then
 *     acc = fma(w[co][ci][ky][kx], in[ci][y+ky-1][x+kx-1], acc) in (ci, ky, kx) order;
 *   - tanh is the deterministic routine tanh_det() below (IEEE +,*,fma,/ and rint only);
 *   - every other elementwise op is a single correctly rounded fp32 operation in the
 *     reference's own order (no contraction: build with -ffp-contract=off).
 *   The reference's MKLDNN convolutions and Sleef tanh use a different (unspecified)
 *   summation order, so oracle-vs-reference agreement is to ~1e-5 on the 0..255 scale
 *   (measured in tests/test_oracle_golden.py); oracle-vs-CUDA agreement is bit-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define NCH 16

typedef struct {
    const float *w1, *b1; /* (16,1,3,3), (16) */
    const float *w2, *b2; /* (16,16,3,3), (16) */
    const float *w3, *b3; /* (16,16,3,3), (16) */
    const float *w4, *b4; /* (1,16,3,3), (1)  */
} orc_pu_t;

typedef struct {
    /* conv_P1, conv_U1, conv_P2, conv_U2: weight (1,1,3,1) + bias (1)   lifting_1d.py:71-89 */
    float tap[4][3];
    float bias[4];
    orc_pu_t pu[4]; /* P_1, U_1, P_2, U_2 in application order          lifting_1d.py:93-96 */
    float scale_l, scale_h; /* lifting_1d.py:98-101 */
    float dynamic_range;    /* 256.0  lifting_1d.py:62 */
    int lossy;
} orc_iwave_t;

/* ------------------------------------------------------------------------------------------ */
/* deterministic tanh (arithmetic contract; table include/pmctf_tanh_table.h, routine specified in
 * tools/gen_tanh_table.py): second-order expansion around the nearest multiple of 1/128 with the derivatives
 * expressed through T = tanh(node).  Absolute error <= 1.2e-7; exact 0 at 0; odd; saturates at |x| >= 9.        */
#include "../include/pmctf_tanh_table.h"
static const uint32_t tanh_bits[PMCTF_TANH_ENTRIES] = {PMCTF_TANH_TABLE_VALUES};
static inline float tanh_det(float x)
{
    const float *tab = (const float *)tanh_bits;
    float m = fmaxf(-fabsf(x), -PMCTF_TANH_XMAX);
    float fi = rintf(m * -PMCTF_TANH_STEPS);            /* rint(128 |x|), ties to even */
    float e = fmaf(fi, 1.0f / PMCTF_TANH_STEPS, m);     /* node - |x|, exact */
    float T = tab[(int32_t)fi];
    float Q = fmaf(T, T, -1.0f);
    float R = T * Q;
    float G = fmaf(e, R, Q);
    float y = fmaf(e, G, T);
    return copysignf(y, x);
}

ORC_API void orc_tanh(const float *x, float *y, long n)
{
    for (long i = 0; i < n; ++i) y[i] = tanh_det(x[i]);
}

/* torch.linspace(-1, 1, n) as the scalar (CUDA-kernel) formula; ATen's vectorised CPU kernel
 * differs in the last bit and depends on the host's SIMD width, so tests that compare with
 * the CPU-generated goldens pass the golden's own tables instead.  video_net.py:36-39 */
ORC_API void orc_linspace(float *out, int n)
{
    const float start = -1.0f, end = 1.0f;
    const float step = (end - start) / (float)(n - 1);
    const int half = n / 2;
    for (int i = 0; i < n; ++i)
        out[i] = (i < half) ? start + step * (float)i : end - step * (float)(n - i - 1);
}

/* ------------------------------------------------------------------------------------------ */
/* flow_warp: video_net.py:32-55.  im [N,C,H,W], flow [flowN,2,H,W] (flowN==1: broadcast over N,
 * the `.tile` of pMCTF_L.py:299-300), sign multiplies the flow (-mv_hat: pMCTF_L.py:307).      */
ORC_API void orc_flow_warp(const float *im, const float *flow, const float *lin_x, const float *lin_y,
                           float *out, int N, int C, int H, int W, int flowN, float sign, int round_out)
{
    const float sx = (float)(((double)W - 1.0) / 2.0); /* python float, cast at the tensor op */
    const float sy = (float)(((double)H - 1.0) / 2.0);
    const float wmax = (float)(W - 1), hmax = (float)(H - 1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y) {
            const float *fxp = flow + ((size_t)(flowN == 1 ? 0 : n) * 2 + 0) * H * W + (size_t)y * W;
            const float *fyp = fxp + (size_t)H * W;
            for (int x = 0; x < W; ++x) {
                float gx = lin_x[x] + (sign * fxp[x]) / sx; /* video_net.py:42-45 */
                float gy = lin_y[y] + (sign * fyp[x]) / sy;
                /* ATen grid_sampler unnormalize (align_corners=True) + border clip */
                float ix = (gx + 1.0f) * sx;
                float iy = (gy + 1.0f) * sy;
                ix = fminf(fmaxf(ix, 0.0f), wmax);
                iy = fminf(fmaxf(iy, 0.0f), hmax);
                float x0 = floorf(ix), y0 = floorf(iy);
                float w = ix - x0, e = 1.0f - w, nn = iy - y0, s = 1.0f - nn;
                float nw = s * e, ne = s * w, sw = nn * e, se = nn * w;
                int x0i = (int)x0, y0i = (int)y0;
                int x1ok = x0i + 1 <= W - 1, y1ok = y0i + 1 <= H - 1;
                for (int c = 0; c < C; ++c) {
                    const float *p = im + ((size_t)n * C + c) * H * W;
                    float vnw = p[(size_t)y0i * W + x0i];
                    float vne = x1ok ? p[(size_t)y0i * W + x0i + 1] : 0.0f;
                    float vsw = y1ok ? p[(size_t)(y0i + 1) * W + x0i] : 0.0f;
                    float vse = (x1ok && y1ok) ? p[(size_t)(y0i + 1) * W + x0i + 1] : 0.0f;
                    float acc = vnw * nw;
                    acc = fmaf(vne, ne, acc);
                    acc = fmaf(vsw, sw, acc);
                    acc = fmaf(vse, se, acc);
                    if (round_out) acc = rintf(acc); /* lossless: pMCTF_L.py:302-303 */
                    out[((size_t)n * C + c) * H * W + (size_t)y * W + x] = acc;
                }
            }
        }
}

/* bilineardownsacling(mv) / 2: video_net.py:66-71, used at pMCTF_L.py:317,336,401.
 * mv [N,2,H,W] -> [N,2,H/2,W/2]; ATen CPU order is ((v00+v01)+v10)+v11 with weights 1/4.       */
ORC_API void orc_chroma_mv_down(const float *mv, float *out, int N, int H, int W)
{
    const int h2 = H / 2, w2 = W / 2;
    for (int p = 0; p < N * 2; ++p)
        for (int y = 0; y < h2; ++y)
            for (int x = 0; x < w2; ++x) {
                const float *a = mv + (size_t)p * H * W + (size_t)(2 * y) * W + 2 * x;
                float v = ((a[0] * 0.25f + a[1] * 0.25f) + a[W] * 0.25f) + a[W + 1] * 0.25f;
                out[(size_t)p * h2 * w2 + (size_t)y * w2 + x] = v / 2.0f;
            }
}

/* ------------------------------------------------------------------------------------------ */
/* PredictUpdate: lifting_1d.py:25-49.  One plane x[H][W] (already multiplied by in_mul by the
 * caller) -> out[H][W].  Band-tiled so the 16-channel intermediates stay in cache.             */
#define XB 64 /* x block held in registers */

/* generic 3x3 layer on a band: in[cin][rows_in][Wp], out[cout][rows_out][Wp]; Wp = W+2 with a
 * zero column on each side; out row r uses in rows r, r+1, r+2 (in starts one row earlier).   */
static void conv3x3_band(const float *in, int cin, int rows_in, const float *wgt, const float *bias,
                         float *out, int cout, int rows_out, int W, int Wp)
{
    for (int co = 0; co < cout; ++co)
        for (int r = 0; r < rows_out; ++r)
            for (int xb = 0; xb < W; xb += XB) {
                const int nx = (W - xb < XB) ? W - xb : XB;
                float acc[XB];
                for (int i = 0; i < XB; ++i) acc[i] = bias[co];
                for (int ci = 0; ci < cin; ++ci)
                    for (int ky = 0; ky < 3; ++ky) {
                        const float *row = in + ((size_t)ci * rows_in + r + ky) * Wp + xb;
                        for (int kx = 0; kx < 3; ++kx) {
                            const float wv = wgt[((co * cin + ci) * 3 + ky) * 3 + kx];
                            if (nx == XB) {
                                for (int i = 0; i < XB; ++i) acc[i] = __builtin_fmaf(wv, row[i + kx], acc[i]);
                            } else {
                                for (int i = 0; i < nx; ++i) acc[i] = __builtin_fmaf(wv, row[i + kx], acc[i]);
                            }
                        }
                    }
                float *o = out + ((size_t)co * rows_out + r) * Wp + 1 + xb;
                for (int i = 0; i < nx; ++i) o[i] = acc[i];
            }
}

/* ---- "exact" convolution mode (PMCTF_CONV_TENSOR of include/pmctf_b200.h) ----------------------------------
 * conv2 / conv3 of PredictUpdate take tanh outputs in [-1, 1].  In this mode they are evaluated in exact integer
 * arithmetic, which makes the result independent of any summation order (so a tensor-core implementation can be
 * bit-identical):  V = rint(a * 2^22);  Wq = rint(w * 2^Sw) with Sw = 22 - e, max|w| = m * 2^e, m in [0.5, 1);
 * S = sum V * Wq (exact, |S| < 2^52);  out = fma((float)S, 2^-(22+Sw), bias)   -- (float)S is one RN rounding.
 * conv4 of this mode sums per-tap partials (conv4_partials_band below).                                          */
static int g_conv_mode = 0;
ORC_API void orc_set_conv_mode(int mode) { g_conv_mode = mode; }
ORC_API int orc_get_conv_mode(void) { return g_conv_mode; }

typedef struct {
    int32_t w[NCH * NCH * 9];
    float down;
} orc_qconv_t;

static void quantise_weights(const float *wgt, orc_qconv_t *q)
{
    float m = 0.0f;
    for (int i = 0; i < NCH * NCH * 9; ++i) m = fmaxf(m, fabsf(wgt[i]));
    int e = 0;
    if (m > 0.0f) frexpf(m, &e);
    const int sw = 22 - e;
    const float up = ldexpf(1.0f, sw);
    for (int i = 0; i < NCH * NCH * 9; ++i) q->w[i] = (int32_t)rintf(wgt[i] * up);
    q->down = ldexpf(1.0f, -(22 + sw));
}

/* same band geometry as conv3x3_band, 16 -> 16 channels */
static void conv3x3_band_exact(const float *in, int rows_in, const orc_qconv_t *q, const float *bias, float *out,
                               int rows_out, int W, int Wp, int32_t *vin)
{
    const size_t nin = (size_t)NCH * rows_in * Wp;
    for (size_t i = 0; i < nin; ++i) vin[i] = (int32_t)rintf(in[i] * 4194304.0f);
    for (int co = 0; co < NCH; ++co)
        for (int r = 0; r < rows_out; ++r) {
            float *o = out + ((size_t)co * rows_out + r) * Wp + 1;
            for (int x = 0; x < W; ++x) {
                int64_t S = 0;
                for (int ci = 0; ci < NCH; ++ci)
                    for (int ky = 0; ky < 3; ++ky) {
                        const int32_t *row = vin + ((size_t)ci * rows_in + r + ky) * Wp + x;
                        const int32_t *wq = q->w + ((co * NCH + ci) * 3 + ky) * 3;
                        S += (int64_t)row[0] * wq[0] + (int64_t)row[1] * wq[1] + (int64_t)row[2] * wq[2];
                    }
                o[x] = __builtin_fmaf((float)S, q->down, bias[co]);
            }
        }
}

/* conv4 (16 -> 1) of the exact mode: per-tap partial sums T_k(p) = sum_ci w4[ci][k] * a3[ci][p] (one fma chain over ci,
 * starting from 0), then out = ((b4 + T_0) + T_1) + ... + T_8 over the 3x3 neighbourhood, k = 3*ky + kx.  The order
 * is a specification of this mode (the tensor-core kernel forms the partials where a pixel's 16 channels sit in
 * registers); the reference's own summation order is unspecified.  Band geometry as conv3x3_band.                  */
static void conv4_partials_band(const float *in, int rows_in, const float *w4, float b4, float *out, int rows_out, int W, int Wp)
{
    for (int r = 0; r < rows_out; ++r) {
        float *o = out + (size_t)r * Wp + 1;
        for (int x = 0; x < W; ++x) {
            float acc = b4;
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    float t = 0.0f;
                    for (int ci = 0; ci < NCH; ++ci)
                        t = __builtin_fmaf(w4[(ci * 3 + ky) * 3 + kx], in[((size_t)ci * rows_in + r + ky) * Wp + x + kx], t);
                    acc = acc + t;
                }
            o[x] = acc;
        }
    }
}

static void band_fix(float *buf, int ch, int rows, int Wp, int W, int row0_img, int H, int do_tanh)
{
    /* zero padding semantics of every layer: rows outside the image and the two border columns
     * are 0 (layers.py:54-56, padding=1 zeros); optional tanh (lifting_1d.py:39,42).           */
    for (int c = 0; c < ch; ++c)
        for (int r = 0; r < rows; ++r) {
            float *p = buf + ((size_t)c * rows + r) * Wp;
            const int yi = row0_img + r;
            if (yi < 0 || yi >= H) {
                memset(p, 0, sizeof(float) * Wp);
                continue;
            }
            p[0] = 0.0f;
            p[W + 1] = 0.0f;
            if (do_tanh)
                for (int x = 1; x <= W; ++x) p[x] = tanh_det(p[x]);
        }
}

#define BAND 16

ORC_API void orc_predict_update(const float *x, const orc_pu_t *pu, float *out, int N, int H, int W, float in_mul)
{
    const int Wp = W + 2;
    const int nb = (H + BAND - 1) / BAND;
    const int exact = g_conv_mode == 1;
    orc_qconv_t q2, q3;
    if (exact) {
        quantise_weights(pu->w2, &q2);
        quantise_weights(pu->w3, &q3);
    }
#pragma omp parallel
    {
        int32_t *vin = exact ? (int32_t *)malloc(sizeof(int32_t) * NCH * (BAND + 6) * Wp) : NULL;
        float *xin = (float *)malloc(sizeof(float) * (BAND + 8) * Wp);
        float *c1 = (float *)malloc(sizeof(float) * NCH * (BAND + 6) * Wp);
        float *a1 = (float *)malloc(sizeof(float) * NCH * (BAND + 6) * Wp);
        float *a2 = (float *)malloc(sizeof(float) * NCH * (BAND + 4) * Wp);
        float *a3 = (float *)malloc(sizeof(float) * NCH * (BAND + 2) * Wp);
        float *o4 = (float *)malloc(sizeof(float) * BAND * Wp);
#pragma omp for collapse(2) schedule(dynamic)
        for (int n = 0; n < N; ++n)
            for (int b = 0; b < nb; ++b) {
                const int y0 = b * BAND;
                const int rows = (H - y0 < BAND) ? H - y0 : BAND;
                const float *xp = x + (size_t)n * H * W;
                /* input rows y0-4 .. y0+rows+4, zero outside */
                for (int r = 0; r < rows + 8; ++r) {
                    float *d = xin + (size_t)r * Wp;
                    const int yi = y0 - 4 + r;
                    if (yi < 0 || yi >= H) {
                        memset(d, 0, sizeof(float) * Wp);
                    } else {
                        d[0] = 0.0f;
                        d[W + 1] = 0.0f;
                        for (int xx = 0; xx < W; ++xx) d[1 + xx] = xp[(size_t)yi * W + xx] * in_mul;
                    }
                }
                /* conv1 -> c1 rows y0-3.. (rows+6) */
                conv3x3_band(xin, 1, rows + 8, pu->w1, pu->b1, c1, NCH, rows + 6, W, Wp);
                memcpy(a1, c1, sizeof(float) * NCH * (rows + 6) * Wp);
                band_fix(c1, NCH, rows + 6, Wp, W, y0 - 3, H, 0);
                band_fix(a1, NCH, rows + 6, Wp, W, y0 - 3, H, 1); /* tanh(conv1) */
                if (exact) conv3x3_band_exact(a1, rows + 6, &q2, pu->b2, a2, rows + 4, W, Wp, vin);
                else conv3x3_band(a1, NCH, rows + 6, pu->w2, pu->b2, a2, NCH, rows + 4, W, Wp);
                band_fix(a2, NCH, rows + 4, Wp, W, y0 - 2, H, 1); /* tanh(conv2) */
                if (exact) conv3x3_band_exact(a2, rows + 4, &q3, pu->b3, a3, rows + 2, W, Wp, vin);
                else conv3x3_band(a2, NCH, rows + 4, pu->w3, pu->b3, a3, NCH, rows + 2, W, Wp);
                /* x = conv1 + conv3   lifting_1d.py:45 */
                for (int c = 0; c < NCH; ++c)
                    for (int r = 0; r < rows + 2; ++r) {
                        float *p3 = a3 + ((size_t)c * (rows + 2) + r) * Wp;
                        const float *p1 = c1 + ((size_t)c * (rows + 6) + r + 2) * Wp;
                        for (int xx = 1; xx <= W; ++xx) p3[xx] = p1[xx] + p3[xx];
                    }
                band_fix(a3, NCH, rows + 2, Wp, W, y0 - 1, H, 0);
                if (exact) conv4_partials_band(a3, rows + 2, pu->w4, pu->b4[0], o4, rows, W, Wp);
                else conv3x3_band(a3, NCH, rows + 2, pu->w4, pu->b4, o4, 1, rows, W, Wp);
                for (int r = 0; r < rows; ++r)
                    memcpy(out + (size_t)n * H * W + (size_t)(y0 + r) * W, o4 + (size_t)r * Wp + 1, sizeof(float) * W);
            }
        free(xin); free(c1); free(a1); free(a2); free(a3); free(o4); free(vin);
    }
}

/* TemporalLifting.predict_filter / update_filter: wavelet_transform_temporal_mctf.py:27-45.
 * out = (x + PU(x)*0.1) * scale   (lossy);  out = x + round(PU(x)*0.1)   (lossless)          */
ORC_API void orc_temporal_filter(const float *x, const orc_pu_t *pu, float scale, int lossy,
                                 float *out, int N, int H, int W)
{
    const size_t n = (size_t)N * H * W;
    float *t = (float *)malloc(sizeof(float) * n);
    orc_predict_update(x, pu, t, N, H, W, 1.0f);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        float tmp = t[i] * 0.1f;
        if (!lossy) tmp = rintf(tmp);
        float v = x[i] + tmp;
        if (lossy) v = v * scale;
        out[i] = v;
    }
    free(t);
}

/* pMCTF.forward_MCTF: pMCTF_L.py:297-312.  Planes [N,H,W] (C=1), mv [mvN,2,H,W].
 * pred / inv may be NULL.                                                                     */
ORC_API void orc_forward_mctf(const float *ref, const float *cur, const float *mv, int mvN,
                              const float *lin_x, const float *lin_y, const orc_pu_t *P_t, const orc_pu_t *U_t,
                              float scale_p, float scale_u, int lossy,
                              float *L, float *Hh, float *pred_out, float *inv_out, int N, int H, int W)
{
    const size_t n = (size_t)N * H * W;
    float *wbuf = (float *)malloc(sizeof(float) * n);
    float *pred = (float *)malloc(sizeof(float) * n);
    orc_flow_warp(ref, mv, lin_x, lin_y, wbuf, N, 1, H, W, mvN, 1.0f, !lossy);
    orc_temporal_filter(wbuf, P_t, scale_p, lossy, pred, N, H, W);
    for (size_t i = 0; i < n; ++i) Hh[i] = cur[i] - pred[i];
    if (pred_out) memcpy(pred_out, pred, sizeof(float) * n);
    orc_flow_warp(Hh, mv, lin_x, lin_y, wbuf, N, 1, H, W, mvN, -1.0f, !lossy);
    orc_temporal_filter(wbuf, U_t, scale_u, lossy, pred, N, H, W);
    for (size_t i = 0; i < n; ++i) L[i] = ref[i] + pred[i];
    if (inv_out) memcpy(inv_out, pred, sizeof(float) * n);
    free(wbuf); free(pred);
}

/* pMCTF.inverse_MCTF: pMCTF_L.py:314-330 (mv already down-scaled by the caller if chroma).     */
ORC_API void orc_inverse_mctf(const float *L, const float *Hh, const float *mv, int mvN,
                              const float *lin_x, const float *lin_y, const orc_pu_t *P_t, const orc_pu_t *U_t,
                              float scale_p, float scale_u, int lossy,
                              float *ref, float *cur, int N, int H, int W)
{
    const size_t n = (size_t)N * H * W;
    float *wbuf = (float *)malloc(sizeof(float) * n);
    float *f = (float *)malloc(sizeof(float) * n);
    orc_flow_warp(Hh, mv, lin_x, lin_y, wbuf, N, 1, H, W, mvN, -1.0f, !lossy);
    orc_temporal_filter(wbuf, U_t, scale_u, lossy, f, N, H, W);
    for (size_t i = 0; i < n; ++i) ref[i] = L[i] - f[i];
    orc_flow_warp(ref, mv, lin_x, lin_y, wbuf, N, 1, H, W, mvN, 1.0f, !lossy);
    orc_temporal_filter(wbuf, P_t, scale_p, lossy, f, N, H, W);
    for (size_t i = 0; i < n; ++i) cur[i] = Hh[i] + f[i];
    free(wbuf); free(f);
}

/* ------------------------------------------------------------------------------------------ */
/* one lifting step of iWave1D on contiguous phases [N,h,W]: lifting_1d.py:104-112 (and the
 * three repeats).  dst = dst + sign * (skip + 0.1 * (PU(skip/256) * 256)),
 * skip = conv(3,1)(reflect-pad-rows(src)) + bias.                                              */
static void lift_step(const float *src, float *dst, const float tap[3], float bias, const orc_pu_t *pu,
                      float dyn, int lossy, float sign, int N, int h, int W)
{
    const size_t n = (size_t)N * h * W;
    float *skip = (float *)malloc(sizeof(float) * n);
    float *l = (float *)malloc(sizeof(float) * n);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < N; ++b)
        for (int y = 0; y < h; ++y) {
            /* ReflectionPad2d((0,0,1,1)): row -1 -> row 1, row h -> row h-2   lifting_1d.py:91 */
            const int ym = (y == 0) ? 1 : y - 1;
            const int yp = (y == h - 1) ? h - 2 : y + 1;
            const float *r0 = src + ((size_t)b * h + ym) * W;
            const float *r1 = src + ((size_t)b * h + y) * W;
            const float *r2 = src + ((size_t)b * h + yp) * W;
            float *o = skip + ((size_t)b * h + y) * W;
            for (int x = 0; x < W; ++x) {
                float acc = bias;
                acc = __builtin_fmaf(tap[0], r0[x], acc);
                acc = __builtin_fmaf(tap[1], r1[x], acc);
                acc = __builtin_fmaf(tap[2], r2[x], acc);
                o[x] = acc;
            }
        }
    /* skip / 256 is an exact power-of-two scaling == * (1/256) */
    orc_predict_update(skip, pu, l, N, h, W, 1.0f / dyn);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        float lv = l[i] * dyn;
        float tmp = skip[i] + lv * 0.1f;
        if (!lossy) tmp = rintf(tmp);
        dst[i] = (sign > 0) ? dst[i] + tmp : dst[i] - tmp;
    }
    free(skip); free(l);
}

/* iWave1D.forward_lift: lifting_1d.py:103-145.  x [N,H,W] -> l,h [N,H/2,W] (split along rows). */
ORC_API void orc_iwave1d_forward(const float *x, const orc_iwave_t *p, float *l, float *hh, int N, int H, int W)
{
    const int h = H / 2;
    for (int b = 0; b < N; ++b)
        for (int y = 0; y < h; ++y) { /* split: lifting_1d.py:10-13 */
            memcpy(l + ((size_t)b * h + y) * W, x + ((size_t)b * H + 2 * y) * W, sizeof(float) * W);
            memcpy(hh + ((size_t)b * h + y) * W, x + ((size_t)b * H + 2 * y + 1) * W, sizeof(float) * W);
        }
    lift_step(l, hh, p->tap[0], p->bias[0], &p->pu[0], p->dynamic_range, p->lossy, +1.0f, N, h, W); /* P1 -> x_o */
    lift_step(hh, l, p->tap[1], p->bias[1], &p->pu[1], p->dynamic_range, p->lossy, +1.0f, N, h, W); /* U1 -> x_e */
    lift_step(l, hh, p->tap[2], p->bias[2], &p->pu[2], p->dynamic_range, p->lossy, +1.0f, N, h, W); /* P2 */
    lift_step(hh, l, p->tap[3], p->bias[3], &p->pu[3], p->dynamic_range, p->lossy, +1.0f, N, h, W); /* U2 */
    if (p->lossy) { /* lifting_1d.py:141-143 */
        const size_t n = (size_t)N * h * W;
        for (size_t i = 0; i < n; ++i) { l[i] = l[i] * p->scale_l; hh[i] = hh[i] * p->scale_h; }
    }
}

/* iWave1D.backward_lift: lifting_1d.py:147-189 (merge: :16-22).                                */
ORC_API void orc_iwave1d_backward(const float *l_in, const float *h_in, const orc_iwave_t *p, float *x, int N, int H, int W)
{
    const int h = H / 2;
    const size_t n = (size_t)N * h * W;
    float *l = (float *)malloc(sizeof(float) * n);
    float *hh = (float *)malloc(sizeof(float) * n);
    for (size_t i = 0; i < n; ++i) {
        l[i] = p->lossy ? l_in[i] / p->scale_l : l_in[i];
        hh[i] = p->lossy ? h_in[i] / p->scale_h : h_in[i];
    }
    lift_step(hh, l, p->tap[3], p->bias[3], &p->pu[3], p->dynamic_range, p->lossy, -1.0f, N, h, W); /* U2 */
    lift_step(l, hh, p->tap[2], p->bias[2], &p->pu[2], p->dynamic_range, p->lossy, -1.0f, N, h, W); /* P2 */
    lift_step(hh, l, p->tap[1], p->bias[1], &p->pu[1], p->dynamic_range, p->lossy, -1.0f, N, h, W); /* U1 */
    lift_step(l, hh, p->tap[0], p->bias[0], &p->pu[0], p->dynamic_range, p->lossy, -1.0f, N, h, W); /* P1 */
    for (int b = 0; b < N; ++b)
        for (int y = 0; y < h; ++y) {
            memcpy(x + ((size_t)b * H + 2 * y) * W, l + ((size_t)b * h + y) * W, sizeof(float) * W);
            memcpy(x + ((size_t)b * H + 2 * y + 1) * W, hh + ((size_t)b * h + y) * W, sizeof(float) * W);
        }
    free(l); free(hh);
}

static void transpose(const float *a, float *b, int N, int H, int W)
{ /* [N,H,W] -> [N,W,H]: the permute(0,1,3,2) of wavelet_transform.py:32-40 */
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) b[((size_t)n * W + x) * H + y] = a[((size_t)n * H + y) * W + x];
}

/* LiftingScheme2D.forward_lift_2d: wavelet_transform.py:25-43.  x [N,H,W] -> 4 x [N,H/2,W/2].
 * l_out/h_out ([N,H/2,W], the row-pass outputs, returned in the reference dict as transposed
 * views) may be NULL.                                                                         */
ORC_API void orc_lift2d_forward(const float *x, const orc_iwave_t *p, float *ll, float *lh, float *hl, float *hh,
                                float *l_out, float *h_out, int N, int H, int W)
{
    const int h2 = H / 2, w2 = W / 2;
    const size_t nh = (size_t)N * h2 * W, nq = (size_t)N * h2 * w2;
    float *l = (float *)malloc(sizeof(float) * nh), *h = (float *)malloc(sizeof(float) * nh);
    float *t = (float *)malloc(sizeof(float) * nh);
    float *a = (float *)malloc(sizeof(float) * nq), *b = (float *)malloc(sizeof(float) * nq);
    orc_iwave1d_forward(x, p, l, h, N, H, W);
    if (l_out) memcpy(l_out, l, sizeof(float) * nh);
    if (h_out) memcpy(h_out, h, sizeof(float) * nh);
    transpose(l, t, N, h2, W);                 /* [N,W,h2] */
    orc_iwave1d_forward(t, p, a, b, N, W, h2); /* -> [N,w2,h2] each; lift_v is lift_h (:20-21) */
    transpose(a, ll, N, w2, h2);
    transpose(b, lh, N, w2, h2);
    transpose(h, t, N, h2, W);
    orc_iwave1d_forward(t, p, a, b, N, W, h2);
    transpose(a, hl, N, w2, h2);
    transpose(b, hh, N, w2, h2);
    free(l); free(h); free(t); free(a); free(b);
}

/* LiftingScheme2D.backward_lift_2d: wavelet_transform.py:45-57.                                */
ORC_API void orc_lift2d_backward(const float *ll, const float *lh, const float *hl, const float *hh,
                                 const orc_iwave_t *p, float *x, int N, int H, int W)
{
    const int h2 = H / 2, w2 = W / 2;
    const size_t nh = (size_t)N * h2 * W, nq = (size_t)N * h2 * w2;
    float *l = (float *)malloc(sizeof(float) * nh), *h = (float *)malloc(sizeof(float) * nh);
    float *t = (float *)malloc(sizeof(float) * nh);
    float *a = (float *)malloc(sizeof(float) * nq), *b = (float *)malloc(sizeof(float) * nq);
    transpose(ll, a, N, h2, w2); /* [N,w2,h2] */
    transpose(lh, b, N, h2, w2);
    orc_iwave1d_backward(a, b, p, t, N, W, h2); /* [N,W,h2] */
    transpose(t, l, N, W, h2);
    transpose(hl, a, N, h2, w2);
    transpose(hh, b, N, h2, w2);
    orc_iwave1d_backward(a, b, p, t, N, W, h2);
    transpose(t, h, N, W, h2);
    orc_iwave1d_backward(l, h, p, x, N, H, W);
    free(l); free(h); free(t); free(a); free(b);
}

/* quantize_subband (+ optional RoundNoGradient): pWave.py:184-189,256-257,337; layers.py:71-92.
 * out = [rint](clamp(s * q, -clip, clip)); torch.round is half-to-even == rintf.               */
ORC_API void orc_quantize(const float *s, float q, float clip, int lossy, int do_round, float *out, long n)
{
    for (long i = 0; i < n; ++i) {
        float v = lossy ? s[i] * q : s[i];
        v = fminf(fmaxf(v, -clip), clip);
        if (do_round) v = rintf(v);
        out[i] = v;
    }
}

/* dequantize_subbands: pWave.py:191-202.  out = s_hat / q                                      */
ORC_API void orc_dequantize(const float *s_hat, float q, int lossy, float *out, long n)
{
    for (long i = 0; i < n; ++i) out[i] = lossy ? s_hat[i] / q : s_hat[i];
}

/* ---- PostProcess (pMCTF/layers/postprocessing.py:20-44) in plain fp32 ---------------------------
 * nn.Conv2d(3x3, zero padding 1) on planar [C][H][W] maps: acc = bias, then fma over (ci, ky, kx) in that order.  The CUDA path
 * evaluates the 64 -> 64 layers on the tensor cores with bf16 operands, so this is a TOLERANCE oracle (1e-3 on the [0,1] pixel
 * scale), pinned to outputs of the reference's own module (tests/golden/postprocess.npz).                                      */
static void conv3x3_planar(const float *in, int ci_n, const float *w, const float *b, int co_n, float *out, int H, int W,
                           float slope, const float *res)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int co = 0; co < co_n; ++co)
        for (int y = 0; y < H; ++y) {
            float *o = out + ((size_t)co * H + y) * W;
            for (int x = 0; x < W; ++x) o[x] = b[co];
            for (int ci = 0; ci < ci_n; ++ci)
                for (int ky = 0; ky < 3; ++ky) {
                    const int yy = y + ky - 1;
                    if (yy < 0 || yy >= H) continue;
                    const float *ip = in + ((size_t)ci * H + yy) * W;
                    const float *wp = w + (((size_t)co * ci_n + ci) * 3 + ky) * 3;
                    for (int kx = 0; kx < 3; ++kx) {
                        const float wv = wp[kx];
                        const int x0 = kx == 0 ? 1 : 0, x1 = kx == 2 ? W - 1 : W;
                        for (int x = x0; x < x1; ++x) o[x] = __builtin_fmaf(wv, ip[x + kx - 1], o[x]);
                    }
                }
            if (res) {
                const float *rp = res + ((size_t)co * H + y) * W;
                for (int x = 0; x < W; ++x) o[x] = o[x] + rp[x];
            }
            if (slope != 1.0f)
                for (int x = 0; x < W; ++x) o[x] = o[x] >= 0.0f ? o[x] : o[x] * slope;
        }
}

typedef struct {
    const float *conv1_w, *conv1_b;     /* [64,1,3,3] */
    const float *res_w[12], *res_b[12]; /* resBlocks.{i}.conv1 / conv2: [64,64,3,3] */
    const float *conv2_w, *conv2_b, *conv3_w, *conv3_b;
} orc_postprocess_t;

/* y = PostProcess(x * in_mul) * out_mul on [N,1,H,W]  (pWave.py:300: in_mul = 1/256, out_mul = 256) */
ORC_API void orc_postprocess(const float *x, const orc_postprocess_t *p, float in_mul, float out_mul, float *y, int N, int H, int W)
{
    const size_t px = (size_t)H * W;
    float *xs = (float *)malloc(sizeof(float) * px), *c1 = (float *)malloc(sizeof(float) * 64 * px);
    float *a = (float *)malloc(sizeof(float) * 64 * px), *t = (float *)malloc(sizeof(float) * 64 * px);
    float *o = (float *)malloc(sizeof(float) * px);
    for (int n = 0; n < N; ++n) {
        for (size_t i = 0; i < px; ++i) xs[i] = x[n * px + i] * in_mul;
        conv3x3_planar(xs, 1, p->conv1_w, p->conv1_b, 64, c1, H, W, 1.0f, NULL);
        memcpy(a, c1, sizeof(float) * 64 * px);
        for (int r = 0; r < 6; ++r) {   /* ResBlock: conv1 -> LeakyReLU(0.2) -> conv2, + input (:14-18) */
            conv3x3_planar(a, 64, p->res_w[2 * r], p->res_b[2 * r], 64, t, H, W, 0.2f, NULL);
            float *nx = (float *)malloc(sizeof(float) * 64 * px);
            conv3x3_planar(t, 64, p->res_w[2 * r + 1], p->res_b[2 * r + 1], 64, nx, H, W, 1.0f, a);
            memcpy(a, nx, sizeof(float) * 64 * px);
            free(nx);
        }
        conv3x3_planar(a, 64, p->conv2_w, p->conv2_b, 64, t, H, W, 1.0f, c1);   /* conv2(tmp) + conv1 (:40) */
        conv3x3_planar(t, 64, p->conv3_w, p->conv3_b, 1, o, H, W, 1.0f, NULL);
        for (size_t i = 0; i < px; ++i) y[n * px + i] = (xs[i] + o[i]) * out_mul;   /* x + tmp (:43) */
    }
    free(xs); free(c1); free(a); free(t); free(o);
}

/* one layer, planar in / out (tests of the single tensor-core convolution) */
ORC_API void orc_conv3x3(const float *in, int ci, const float *w, const float *b, int co, float *out, int H, int W, float slope,
                         const float *res)
{
    conv3x3_planar(in, ci, w, b, co, out, H, W, slope, res);
}

/* ---- generic nn.Conv2d(k x k, stride 1, zero padding k/2, groups) on ONE planar map [ci][H][W] -> [co][H][W], plain fp32: acc = bias,
 * then fma over (ci, ky, kx).  Building block of the fp32 TOLERANCE oracle of the entropy-parameter networks
 * (oracle/ctx_oracle.py: pMCTF/layers/context_fusion_4step.py, layers/video/layers.py:113-172). */
ORC_API void orc_conv2d(const float *in, int ci_n, const float *w, const float *b, int co_n, int k, int groups, float *out, int H, int W)
{
    const int cig = ci_n / groups, cog = co_n / groups, r = k / 2;
#pragma omp parallel for collapse(2) schedule(static)
    for (int co = 0; co < co_n; ++co)
        for (int y = 0; y < H; ++y) {
            float *o = out + ((size_t)co * H + y) * W;
            const int g = co / cog;
            for (int x = 0; x < W; ++x) o[x] = b ? b[co] : 0.0f;
            for (int c = 0; c < cig; ++c)
                for (int ky = 0; ky < k; ++ky) {
                    const int yy = y + ky - r;
                    if (yy < 0 || yy >= H) continue;
                    const float *ip = in + ((size_t)(g * cig + c) * H + yy) * W;
                    const float *wp = w + (((size_t)co * cig + c) * k + ky) * k;
                    for (int kx = 0; kx < k; ++kx) {
                        const float wv = wp[kx];
                        const int d = kx - r, x0 = d < 0 ? -d : 0, x1 = d > 0 ? W - d : W;
                        for (int x = x0; x < x1; ++x) o[x] = __builtin_fmaf(wv, ip[x + d], o[x]);
                    }
                }
        }
}

/* ---- the LL subband's autoregressive entropy-parameter model, coefficient by coefficient in raster order: ContextFusionSubband
 * (pMCTF/layers/context_fusion.py:56-204, forward_sequential) as driven by pWave._compress_subband_ar (pMCTF/models/pWave.py:531-553).
 * fp32 CONTRACT shared as a specification with csrc/pmctf_llar.cu:
 *   - a masked 3x3 layer reads the causal taps (dy, dx) = (-1,-1), (-1,0), (-1,1), (0,-1) [type A] + (0,0) [type B] of a zero-padded map;
 *   - its reduction index k = tap * 128 + ci (1x1 layers: k = ci) is cut into EIGHT equal slices; a slice is one fma chain from 0 in k
 *     order; out = ((bias + slice_0) + slice_1) + ... + slice_7;
 *   - the 1 -> 128 input layer is one fma chain from the bias over its four taps;
 *   - residual blocks (context_fusion.py:30-41): t = lrelu(conv1(x)); x = conv2(t) + x; after the two blocks x = x + first;
 *   - tail: lrelu, 1x1, lrelu, 1x1, lrelu, then the 128 -> 2 layer as t[c] * w[c] products folded by the tree r[i] += r[i + s],
 *     s = 64, 32, ..., 1, plus its bias;
 *   - symbol = rint(rint(y) - mean); the history the next coefficients see is the RECONSTRUCTION rint(symbol + mean).
 * Weights in their state_dict (OIHW) layout; masked taps are simply never read.  w128: the five masked 128 -> 128 layers
 * (residualBlocks.0.conv1, .0.conv2, .1.conv1, .1.conv2, maskedConv2), each [128][128][3][3]; w1: convs.0, convs.1 [128][128]; wout [2][128]. */
static inline float orc_lrelu(float v) { return v >= 0.0f ? v : v * 0.2f; }

ORC_API void orc_llar_encode(const float *w_in, const float *b_in, const float *w128, const float *b128, const float *w1, const float *b1,
                             const float *wout, const float *bout, const float *yq, int H, int W, float *scales, float *means, float *symbols,
                             float *recon)
{
    enum { F = 128, KS = 8 };
    static const int DY[5] = {-1, -1, -1, 0, 0}, DX[5] = {-1, 0, 1, -1, 0};
    const int Hp = H + 2, Wp = W + 2;
    float *Y = (float *)calloc((size_t)Hp * Wp, sizeof(float));
    float *hist = (float *)calloc((size_t)5 * Hp * Wp * F, sizeof(float));   /* inputs of the five masked layers, channel-last */
    float in[5 * F], out[F], cur[F], first[F], red0[F], red1[F];
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
            const size_t here = ((size_t)(h + 1) * Wp + (w + 1)) * F;
            for (int co = 0; co < F; ++co) {
                float t = b_in[co];
                for (int k = 0; k < 4; ++k) t = __builtin_fmaf(w_in[(size_t)co * 9 + (DY[k] + 1) * 3 + (DX[k] + 1)], Y[(size_t)(h + 1 + DY[k]) * Wp + (w + 1 + DX[k])], t);
                cur[co] = first[co] = t;
            }
            for (int L = 0; L < 5; ++L) {
                float *plane = hist + (size_t)L * Hp * Wp * F;
                if (L == 4)
                    for (int c = 0; c < F; ++c) cur[c] = cur[c] + first[c];
                if (L == 1 || L == 3) {
                    for (int c = 0; c < F; ++c) plane[here + c] = orc_lrelu(out[c]);
                } else {
                    for (int c = 0; c < F; ++c) plane[here + c] = cur[c];
                }
                for (int t = 0; t < 5; ++t)
                    for (int c = 0; c < F; ++c) in[t * F + c] = plane[((size_t)(h + 1 + DY[t]) * Wp + (w + 1 + DX[t])) * F + c];
                const float *wl = w128 + (size_t)L * F * F * 9, *bl = b128 + (size_t)L * F;
                const int per = 5 * F / KS;
                for (int co = 0; co < F; ++co) {
                    float v = bl[co];
                    for (int s = 0; s < KS; ++s) {
                        float acc = 0.0f;
                        for (int j = 0; j < per; ++j) {
                            const int k = s * per + j, t = k / F, ci = k % F;
                            acc = __builtin_fmaf(wl[((size_t)co * F + ci) * 9 + (DY[t] + 1) * 3 + (DX[t] + 1)], in[k], acc);
                        }
                        v += acc;
                    }
                    out[co] = v;
                }
                if (L == 1 || L == 3)
                    for (int c = 0; c < F; ++c) cur[c] = out[c] + cur[c];
            }
            for (int k1 = 0; k1 < 2; ++k1) {
                for (int c = 0; c < F; ++c) in[c] = orc_lrelu(out[c]);
                const float *wl = w1 + (size_t)k1 * F * F, *bl = b1 + (size_t)k1 * F;
                const int per = F / KS;
                for (int co = 0; co < F; ++co) {
                    float v = bl[co];
                    for (int s = 0; s < KS; ++s) {
                        float acc = 0.0f;
                        for (int j = 0; j < per; ++j) acc = __builtin_fmaf(wl[(size_t)co * F + s * per + j], in[s * per + j], acc);
                        v += acc;
                    }
                    cur[co] = v;          /* scratch: out is still being read through `in` only */
                }
                for (int c = 0; c < F; ++c) out[c] = cur[c];
            }
            for (int c = 0; c < F; ++c) {
                const float t = orc_lrelu(out[c]);
                red0[c] = t * wout[c];
                red1[c] = t * wout[F + c];
            }
            for (int st = F / 2; st > 0; st >>= 1)
                for (int c = 0; c < st; ++c) {
                    red0[c] += red0[c + st];
                    red1[c] += red1[c + st];
                }
            const float scale = red0[0] + bout[0], mean = red1[0] + bout[1];
            const float sym = rintf(rintf(yq[(size_t)h * W + w]) - mean), rec = rintf(sym + mean);
            Y[(size_t)(h + 1) * Wp + (w + 1)] = rec;
            scales[(size_t)h * W + w] = scale;
            means[(size_t)h * W + w] = mean;
            symbols[(size_t)h * W + w] = sym;
            recon[(size_t)h * W + w] = rec;
        }
    free(Y);
    free(hist);
}
