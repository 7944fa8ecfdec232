// pmctf_common.cuh -- device helpers shared by the lifting-step kernels (FFMA and tensor-core versions).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pmctf_b200.h"
#include "pmctf_tanh_table.h"

namespace pmctf {

// ------------------------------------------------------------------------------------------
// deterministic tanh (arithmetic contract: table include/pmctf_tanh_table.h, routine specified in
// tools/gen_tanh_table.py, restated independently in oracle/pmctf_oracle.c): second-order expansion around the
// nearest multiple of 1/128, derivatives from T = tanh(node); one 4-byte shared-memory lookup per evaluation
constexpr int TANH_BULK_BYTES = ((PMCTF_TANH_ENTRIES * 4 + 15) / 16) * 16;   // the table as one 16-byte-granular bulk copy
__device__ __align__(16) const unsigned int g_tanh_bits[TANH_BULK_BYTES / 4] = {PMCTF_TANH_TABLE_VALUES};
constexpr int TANH_SMEM_BYTES = ((PMCTF_TANH_ENTRIES * 4 + 127) / 128) * 128;

__device__ __forceinline__ void load_tanh_table(float *tab_smem, int tid, int nthreads)
{
    for (int i = tid; i < PMCTF_TANH_ENTRIES; i += nthreads) tab_smem[i] = __uint_as_float(g_tanh_bits[i]);
}

__device__ __forceinline__ float tanh_det(float x, const float *__restrict__ tab)
{
    const float m = fmaxf(-fabsf(x), -PMCTF_TANH_XMAX);
    const float fi = rintf(m * -PMCTF_TANH_STEPS);
    const float e = fmaf(fi, 1.0f / PMCTF_TANH_STEPS, m);
#if defined(PMCTF_WHATIF) && (PMCTF_WHATIF & 2)
    const float T = fi * 0.0008f;
#else
    const float T = tab[(int)fi];
#endif
    const float Q = fmaf(T, T, -1.0f);
    const float R = T * Q;
    const float G = fmaf(e, R, Q);
    const float y = fmaf(e, G, T);
    return copysignf(y, x);
}

// two independent IEEE fp32 FMAs in one instruction (sm_100 FFMA2): acc.x = fma(a.x, b, acc.x), acc.y = fma(a.y, b, acc.y).
// Bit-identical to two fmaf() calls; it only halves the issue slots of the CUDA-core convolution chains.
__device__ __forceinline__ float2 ffma2(float2 a, float b, float2 acc)
{
    unsigned long long d;
    const float2 bb = make_float2(b, b);
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long *>(&a)), "l"(*reinterpret_cast<const unsigned long long *>(&bb)),
          "l"(*reinterpret_cast<const unsigned long long *>(&acc)));
    return *reinterpret_cast<float2 *>(&d);
}

__device__ __forceinline__ float2 fma2v(float2 a, float2 b, float2 c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long *>(&a)), "l"(*reinterpret_cast<const unsigned long long *>(&b)),
          "l"(*reinterpret_cast<const unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 mul2v(float2 a, float2 b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long *>(&a)), "l"(*reinterpret_cast<const unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&d);
}

__device__ __forceinline__ float2 add2v(float2 a, float2 b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long *>(&a)), "l"(*reinterpret_cast<const unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&d);
}

// default since the r2w11 A/B (profiles/r2_whatif.txt): +1.3 % on the tensor-core step, same values; -DPMCTF_TANH_NO_IMAD restores the
// shift / mask / add addressing
#if !defined(PMCTF_TANH_NO_IMAD) && !defined(PMCTF_TANH_IMAD)
#define PMCTF_TANH_IMAD 1
#endif
// PMCTF_TANH_IMAD: (shared-memory address of the table) - 4 * 0x4B400000, made opaque so that it stays ONE register and the
// entry's address is one shift-add of the rounded float's bits (see tanh_det2)
__device__ __forceinline__ const float *tanh_table_base(const float *tab_smem)
{
    // + entry 0 of the table in global memory (tanh(0) = +0.0f: the bits are 0), which the assembler cannot fold: otherwise it
    // splits the constant back into an immediate offset of the load plus an add per lookup
    const unsigned int zero = *reinterpret_cast<const volatile unsigned int *>(g_tanh_bits);
    const unsigned int tb = (unsigned int)__cvta_generic_to_shared(tab_smem) - 0x2D000000u + zero;
    return reinterpret_cast<const float *>((size_t)tb);
}

// tanh_det of two values at once on the packed fp32x2 pipe: the same values per element.  rint(128 |x|) and the table index come
// from the 1.5 * 2^23 trick (fma(m, -128, M) = M + rint(128 |x|) exactly, ties to even like rintf) instead of FRND + F2I, which
// keeps the conversion (XU) pipe out of the hot loop; the sign arrangement of the contract needs no packed negation.
__device__ __forceinline__ float2 tanh_det2(float2 x, const float *__restrict__ tab)
{
    constexpr float M = 12582912.0f; // 1.5 * 2^23
    const float2 m = make_float2(fmaxf(-fabsf(x.x), -PMCTF_TANH_XMAX), fmaxf(-fabsf(x.y), -PMCTF_TANH_XMAX));
#if defined(PMCTF_WHATIF) && (PMCTF_WHATIF & 16)   // timing experiment: third-order expansion on a 1/32 grid (wrong values from the 1/128 table)
    {
        const float2 tm = fma2v(m, make_float2(-32.0f, -32.0f), make_float2(M, M));
        const float2 fi = add2v(tm, make_float2(-M, -M));
        const float2 e = fma2v(fi, make_float2(1.0f / 32.0f, 1.0f / 32.0f), m);
        const float2 T = make_float2(tab[__float_as_int(tm.x) & 0x7FF], tab[__float_as_int(tm.y) & 0x7FF]);
        const float2 Q = fma2v(T, T, make_float2(-1.0f, -1.0f));
        const float2 R = mul2v(T, Q);
        const float2 Qc = mul2v(Q, make_float2(0.6666667f, 0.6666667f));
        const float2 S = fma2v(Q, Q, Qc);
        const float2 G2 = fma2v(e, S, R);
        const float2 G = fma2v(e, G2, Q);
        const float2 y = fma2v(e, G, T);
        return make_float2(copysignf(y.x, x.x), copysignf(y.y, x.y));
    }
#endif
    const float2 tm = fma2v(m, make_float2(-PMCTF_TANH_STEPS, -PMCTF_TANH_STEPS), make_float2(M, M));
    const float2 fi = add2v(tm, make_float2(-M, -M));
    const float2 e = fma2v(fi, make_float2(1.0f / PMCTF_TANH_STEPS, 1.0f / PMCTF_TANH_STEPS), m);
#if defined(PMCTF_WHATIF) && (PMCTF_WHATIF & 2)
    const float2 T = make_float2(fi.x * 0.0008f, fi.y * 0.0008f);
#else
#if defined(PMCTF_TANH_IMAD)
    // the float's bits are 0x4B400000 + n (n = the table index, <= 1152): the shared-memory address of entry n is ONE shift-add,
    // (bits << 2) + (table address - 4 * 0x4B400000) in wrap-around 32-bit arithmetic, instead of a shift, a mask and an add
    const unsigned int tb = (unsigned int)(size_t)tab;     // PMCTF_TANH_IMAD: `tab` carries tanh_table_base(), not a pointer
    float2 T;
    asm("ld.shared.f32 %0, [%1];" : "=f"(T.x) : "r"((unsigned int)__float_as_int(tm.x) * 4u + tb));
    asm("ld.shared.f32 %0, [%1];" : "=f"(T.y) : "r"((unsigned int)__float_as_int(tm.y) * 4u + tb));
#else
    const float2 T = make_float2(tab[__float_as_int(tm.x) & 0x7FF], tab[__float_as_int(tm.y) & 0x7FF]);
#endif
#endif
    const float2 Q = fma2v(T, T, make_float2(-1.0f, -1.0f));
    const float2 R = mul2v(T, Q);
    const float2 G = fma2v(e, R, Q);
    const float2 y = fma2v(e, G, T);
    return make_float2(copysignf(y.x, x.x), copysignf(y.y, x.y));
}

struct PlaneD {
    float *p;
    long long gs, bs, rs, cs;
    int group_n; // 0: single level (n * bs)
};

__device__ __forceinline__ long long plane_off(const PlaneD &pl, int n)
{
    if (pl.group_n <= 0) return (long long)n * pl.bs;
    const int g = n / pl.group_n;
    return (long long)g * pl.gs + (long long)(n - g * pl.group_n) * pl.bs;
}

struct StepD {
    int n, div_group_n, h, w; // div_group_n: planes >= this index use base_div1_g1
    int mode;
    PlaneD src;
    float src_div1, src_div2;
    const float *mv;
    int mv_share, mv_down, mv_h, mv_w; // mv_share = n / mv_n
    float mv_sign, sx, sy;
    const float *lin_x, *lin_y;
    int round_src;
    float tap0, tap1, tap2, tap_bias;
    const float *pu_packed;
    float in_mul, post_mul, out_mul;
    int round_tmp;
    PlaneD base;
    float base_div1, base_div1_g1, base_div2, sign, final_mul; // base_div1_g1: divisor for plane group >= 1
    PlaneD out, pred, aux;
    float aux_mul;
};

// bilinear border-clamped backward warp of one sample: video_net.py:42-50 + ATen grid_sampler_2d
// (align_corners=True, padding_mode=border), op for op as in oracle/pmctf_oracle.c:orc_flow_warp
__device__ __forceinline__ float warp_sample(const float *__restrict__ im, long long rs, long long cs, int H, int W,
                                             float lx, float ly, float fx, float fy, float sx, float sy)
{
    float gx = lx + fx / sx;
    float gy = ly + fy / sy;
    float ix = (gx + 1.0f) * sx;
    float iy = (gy + 1.0f) * sy;
    ix = fminf(fmaxf(ix, 0.0f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.0f), (float)(H - 1));
    float x0 = floorf(ix), y0 = floorf(iy);
    float w = ix - x0, e = 1.0f - w, nn = iy - y0, s = 1.0f - nn;
    float nw = s * e, ne = s * w, sw = nn * e, se = nn * w;
    int x0i = (int)x0, y0i = (int)y0;
    bool x1ok = x0i + 1 <= W - 1, y1ok = y0i + 1 <= H - 1;
    const float *p = im + (long long)y0i * rs + (long long)x0i * cs;
    float vnw = __ldg(p);
    float vne = x1ok ? __ldg(p + cs) : 0.0f;
    float vsw = y1ok ? __ldg(p + rs) : 0.0f;
    float vse = (x1ok && y1ok) ? __ldg(p + rs + cs) : 0.0f;
    float acc = vnw * nw;
    acc = fmaf(vne, ne, acc);
    acc = fmaf(vsw, sw, acc);
    acc = fmaf(vse, se, acc);
    return acc;
}

// motion vector at (y, x) of plane n; mv_down fuses bilineardownsacling(mv)/2 (video_net.py:66-71)
__device__ __forceinline__ void load_mv(const float *__restrict__ mv, int mv_share, int mv_down, int mv_h, int mv_w, int n,
                                        int y, int x, float sign, float &fx, float &fy)
{
    const long long plane = (long long)mv_h * mv_w;
    const float *b = mv + (long long)(n / mv_share) * 2 * plane; // mv_share consecutive planes use one field
    if (!mv_down) {
        fx = sign * __ldg(b + (long long)y * mv_w + x);
        fy = sign * __ldg(b + plane + (long long)y * mv_w + x);
    } else {
        const float *a = b + (long long)(2 * y) * mv_w + 2 * x;
        float2 r0 = __ldg(reinterpret_cast<const float2 *>(a));
        float2 r1 = __ldg(reinterpret_cast<const float2 *>(a + mv_w));
        fx = sign * ((((r0.x * 0.25f + r0.y * 0.25f) + r1.x * 0.25f) + r1.y * 0.25f) / 2.0f);
        a += plane;
        r0 = __ldg(reinterpret_cast<const float2 *>(a));
        r1 = __ldg(reinterpret_cast<const float2 *>(a + mv_w));
        fy = sign * ((((r0.x * 0.25f + r0.y * 0.25f) + r1.x * 0.25f) + r1.y * 0.25f) / 2.0f);
    }
}


} // namespace pmctf
