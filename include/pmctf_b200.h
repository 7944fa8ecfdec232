/*
 * pmctf_b200.h -- C ABI of the B200 (sm_100a) implementation of Learned-pMCTF's
 * motion-compensated temporal lifting + pWave++ spatial lifting hot path.
 *
 * The reference has no FFI for this path: its boundary is the Python nn.Module surface
 * (SURVEY.md section 8b).  Each entry point below replaces the reference function cited beside it
 * (paths relative to the reference repository) and is what a maintainer would bind with ctypes
 * from those modules (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer to fp32 unless noted;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - nothing here allocates, frees or synchronises: the caller owns inputs, outputs and
 *     workspaces and keeps them alive until the stream has drained;
 *   - every function returns 0 on success, a negative PMCTF_E* code for rejected arguments
 *     (nothing was launched) or a positive cudaError_t from the launch;
 *   - there is no CPU fallback.
 *
 * Arithmetic contract (DESIGN.md "Numerics"): fp32 throughout; every convolution output is one
 * sequential fma chain in (ci, ky, kx) order starting from the bias; tanh is the deterministic
 * routine documented in DESIGN.md; all other ops are single correctly rounded fp32 operations
 * in the reference's order.  oracle/pmctf_oracle.c restates the same contract on the CPU and
 * the two agree bit for bit.
 */
#ifndef PMCTF_B200_H
#define PMCTF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PMCTF_ABI_VERSION 2

#define PMCTF_EINVAL (-1)   /* null pointer / non-positive size / unsupported flag */
#define PMCTF_ESHAPE (-2)   /* shape not supported (odd split size, plane too small for reflection) */
#define PMCTF_EWORKSPACE (-3) /* workspace too small */
#define PMCTF_ETIMEOUT (-4) /* an earlier tensor-core launch on this device gave up waiting for its MMAs; nothing was launched */

#define PMCTF_PU_PACKED_FLOATS 10128 /* size of one packed PredictUpdate weight block (fp32 taps + int8 tensor-core operands) */

/* How the two 16->16 convolutions of PredictUpdate (lifting_1d.py:40-44) are evaluated (DESIGN.md "Numerics"):
 *   PMCTF_CONV_TENSOR  exact fixed-point implicit GEMM on the tcgen05 tensor cores (default)
 *   PMCTF_CONV_FFMA    sequential fp32 FMA chains on the CUDA cores
 * Each mode is bit-exact against the oracle run in the same mode.
 * Every descriptor that reaches a PredictUpdate (pmctf_step_t, pmctf_iwave_t, pmctf_temporal_t) carries a conv_mode
 * field: PMCTF_CONV_DEFAULT (= 0, what a zeroed struct holds) follows the process-wide default set by
 * pmctf_set_conv_mode(); the other two values select the arithmetic for that call only. */
#define PMCTF_CONV_DEFAULT 0
#define PMCTF_CONV_FFMA 1
#define PMCTF_CONV_TENSOR 2
#define PMCTF_IWAVE_PACKED_FLOATS (4 * PMCTF_PU_PACKED_FLOATS)

/* A strided view of a batch of single-channel planes (strides in elements).  Element (n, y, x)
 * lives at p[n*bs + y*rs + x*cs] when group_n == 0, and at
 * p[(n / group_n)*gs + (n % group_n)*bs + y*rs + x*cs] otherwise (two-level batch: e.g. the Cb/Cr
 * planes of every second frame of a GOP).  This is how split (lifting_1d.py:10-13), merge
 * (:16-22), the permute(0,1,3,2) views of wavelet_transform.py:32-54 and the ref/cur frame
 * selection of the dyadic GOP schedule (test_pMCTF_flex.py:143-146) are expressed without copies. */
typedef struct {
    float *p;
    long long bs, rs, cs;
    long long gs;
    int group_n;
} pmctf_plane_t;

enum { PMCTF_SRC_PLANE = 0, PMCTF_SRC_WARP = 1, PMCTF_SRC_SKIP3 = 2 };
enum { PMCTF_MODE_ACCUM = 0, PMCTF_MODE_FILTER = 1, PMCTF_MODE_PU = 2 };

/* One fused lifting step:
 *      s   = source(src)                         plane | warp(src, sign*mv) | 3-tap skip conv of src
 *      t   = PU(s * in_mul) * post_mul           PredictUpdate, lifting_1d.py:25-49
 *      tmp = s + t * 0.1   [rint if round_tmp]   lifting_1d.py:108-111 / wavelet_transform_temporal_mctf.py:28-32
 *      r   = tmp * out_mul                       scale_p / scale_u (wavelet_transform_temporal_mctf.py:33-35)
 *      out = (base/base_div1/base_div2 + sign*r) * final_mul     MODE_ACCUM
 *      out = r                                                  MODE_FILTER
 *      out = t                                                  MODE_PU
 *      pred (optional) = r ; aux (optional, SKIP3/PLANE) = (src/src_div1/src_div2) * aux_mul
 * All planes are logical [n, h, w] with arbitrary strides. */
typedef struct {
    int n, h, w;
    int src_kind, mode;
    pmctf_plane_t src;
    float src_div1, src_div2;
    /* PMCTF_SRC_WARP: video_net.py:32-55 */
    const float *mv;      /* [mv_n, 2, mv_h, mv_w]; n/mv_n consecutive planes share a field (mv_n == 1: the
                           * `.tile` of pMCTF_L.py:299-300; mv_n == n/2: batched chroma pairs) */
    int mv_n, mv_down;    /* mv_down: mv is the LUMA field [.,2,2h,2w]; use 2x2 mean / 2 (video_net.py:66-71) */
    float mv_sign;
    const float *lin_x, *lin_y; /* torch.linspace(-1,1,w) / (-1,1,h) tables (video_net.py:36-39) */
    int round_src;        /* lossless: pMCTF_L.py:302-303 */
    /* PMCTF_SRC_SKIP3: conv (3,1) + bias on the row-reflect-padded source, lifting_1d.py:105-106 */
    float tap0, tap1, tap2, tap_bias;
    /* PredictUpdate weights packed by pmctf_pack_pu_weights */
    const float *pu_packed;
    float in_mul, post_mul, out_mul;
    int round_tmp;
    pmctf_plane_t base;
    float base_div1, base_div2, sign, final_mul;
    pmctf_plane_t out, pred, aux;
    float aux_mul;
    int conv_mode;        /* PMCTF_CONV_* for this call (0 = process default) */
} pmctf_step_t;

/* iWave1D parameters (lifting_1d.py:52-101) */
typedef struct {
    float tap[4][3];       /* conv_P1, conv_U1, conv_P2, conv_U2 weights (1,1,3,1) */
    float bias[4];
    const float *pu_packed; /* 4 packed blocks: P_1, U_1, P_2, U_2 */
    float scale_l, scale_h; /* lifting_1d.py:98-101 */
    float dynamic_range;    /* 256: lifting_1d.py:62 */
    int lossy;
    int conv_mode;          /* PMCTF_CONV_* for calls with this descriptor (0 = process default) */
} pmctf_iwave_t;

/* TemporalLifting parameters (wavelet_transform_temporal_mctf.py:11-25) */
typedef struct {
    const float *P_t_packed, *U_t_packed;
    float scale_p, scale_u; /* 1/sqrt(2), 0.5 */
    int lossy;
    int conv_mode;          /* PMCTF_CONV_* for calls with this descriptor (0 = process default) */
} pmctf_temporal_t;

int pmctf_abi_version(void);
const char *pmctf_error_string(int code);
/* Number of CUDA kernels launched by this library so far in this process (instrumentation for bench.py). */
unsigned long long pmctf_launch_count(void);
/* Process-wide DEFAULT of the convolution arithmetic (PMCTF_CONV_FFMA or PMCTF_CONV_TENSOR; used by descriptors whose
 * conv_mode is PMCTF_CONV_DEFAULT); returns 0 or PMCTF_EINVAL.  Atomic; the per-descriptor field is the re-entrant way. */
int pmctf_set_conv_mode(int mode);
int pmctf_get_conv_mode(void);
/* Tensor-core watchdog.  Every mbarrier wait inside the tensor-core step kernel is bounded; a kernel that gives up sets an
 * error word in mapped pinned host memory (per device) and ALL its CTAs leave together.  From then on every call that would
 * launch a tensor-core step on that device -- and pmctf_pack_pu_weights -- returns PMCTF_ETIMEOUT without launching, until
 * pmctf_tc_clear_error().  pmctf_tc_error_flag() reads the word of the current device (no copy, no synchronisation: call it
 * after synchronising the streams whose kernels should be covered); pmctf_tc_inject_timeout() sets it (tests). */
int pmctf_tc_error_flag(void);
int pmctf_tc_clear_error(void);
int pmctf_tc_inject_timeout(void);
/* clock64 stamps of the phases of one CTA of the most recent tensor-core launches (profiling aid; synchronises):
 * [0..6] start, source ready, conv1 done, conv2 done, conv3 done, conv4 done, end; [8..11] MMA issue begin/end of
 * conv2, conv3; [12..13] cycles an epilogue warp waited for accumulators. */
int pmctf_tc_debug_times(long long *out16);
/* MMA throughput probe on an otherwise idle SM: issues reps x 18 kind::i8 MMAs (variant 0: the convolution's block
 * pattern; 1/2/3: N = 16/48/96 with overlapping K chunks; 4: N = 48, disjoint K chunks) and writes
 * {issue cycles, issue+completion cycles, #MMAs} to out3_device. */
int pmctf_tc_mma_probe(int variant, int reps, long long *out3_device, void *stream);

/* Repack one PredictUpdate's 8 tensors (OIHW, as in the state_dict) into the kernel layout.
 * Replaces nothing in the reference; run once per weight version.  `packed` holds PMCTF_PU_PACKED_FLOATS floats and must be
 * 16-byte aligned.  The call also reads the packed block back once
 * (40 KB, synchronises `stream`) and registers its small fp32 parameters (conv1, conv4, biases) under the address of
 * `packed`: the tensor-core step kernel takes them as kernel arguments.  A step launched with a `pu_packed` block that
 * did not come from this call (e.g. a device-side copy of one) is rejected with PMCTF_EINVAL in tensor mode. */
int pmctf_pack_pu_weights(const float *w1, const float *b1, const float *w2, const float *b2,
                          const float *w3, const float *b3, const float *w4, const float *b4,
                          float *packed, void *stream);
/* Forget the parameters registered for `packed` (call before freeing or reusing the block's memory): a later launch that names
 * the address without a fresh pmctf_pack_pu_weights() is rejected with PMCTF_EINVAL instead of running with stale values. */
int pmctf_release_pu_weights(const float *packed);

/* flow_warp(im, flow): pMCTF/layers/video/video_net.py:32-55.
 * im [N,C,H,W], flow [flowN,2,H,W] in pixels (flowN divides N), out [N,C,H,W]. */
int pmctf_flow_warp(const float *im, const float *flow, const float *lin_x, const float *lin_y, float *out,
                    int N, int C, int H, int W, int flowN, float sign, int round_out, void *stream);

/* bilineardownsacling(mv) / 2: video_net.py:66-71 as used at pMCTF_L.py:317,336,401.
 * mv [N,2,H,W] -> out [N,2,H/2,W/2]. */
int pmctf_chroma_mv_down(const float *mv, float *out, int N, int H, int W, void *stream);

/* The generic fused step (see pmctf_step_t). */
int pmctf_lift_step(const pmctf_step_t *step, void *stream);

/* PredictUpdate.forward: lifting_1d.py:36-49.  x, out dense [N,1,H,W]. */
int pmctf_predict_update(const float *x, const float *pu_packed, float in_mul, float *out,
                         int N, int H, int W, int conv_mode, void *stream);

/* TemporalLifting.predict_filter / update_filter: wavelet_transform_temporal_mctf.py:27-45
 * (which = 0 predict, 1 update). */
int pmctf_temporal_filter(const float *x, const pmctf_temporal_t *t, int which, float *out,
                          int N, int H, int W, void *stream);

/* pMCTF.forward_MCTF: pMCTF/models/video/pMCTF_L.py:297-312.  ref, cur, L, H: N logical planes
 * [H, W] each (any strides); mv dense [mv_n,2,H,W] (or the luma field [mv_n,2,2H,2W] with
 * mv_down=1, fusing pMCTF_L.py:401), N/mv_n consecutive planes share one field;
 * pred / inv may be NULL (they only feed the MSE terms, pMCTF_L.py:351,373).
 * L must not alias ref or cur; H may alias cur. */
int pmctf_forward_mctf(const pmctf_plane_t *ref, const pmctf_plane_t *cur, const float *mv, int mv_n, int mv_down,
                       const float *lin_x, const float *lin_y, const pmctf_temporal_t *t,
                       const pmctf_plane_t *L, const pmctf_plane_t *Hh, const pmctf_plane_t *pred,
                       const pmctf_plane_t *inv, int N, int H, int W, void *stream);

/* pMCTF.inverse_MCTF: pMCTF_L.py:314-330 (mv_down=1 == downscale=True). */
int pmctf_inverse_mctf(const pmctf_plane_t *L, const pmctf_plane_t *Hh, const float *mv, int mv_n, int mv_down,
                       const float *lin_x, const float *lin_y, const pmctf_temporal_t *t,
                       const pmctf_plane_t *ref, const pmctf_plane_t *cur, int N, int H, int W, void *stream);

/* iWave1D.forward_lift / backward_lift on strided views: lifting_1d.py:103-189.
 * x is the logical [n, 2*h2, w] input; l, h are logical [n, h2, w] outputs (any strides).
 * workspace: n*h2*w floats (forward: the unscaled high band), 2*n*h2*w floats (backward). */
int pmctf_iwave1d_forward(const pmctf_plane_t *x, const pmctf_iwave_t *p, const pmctf_plane_t *l,
                          const pmctf_plane_t *h, int n, int h2, int w, float *workspace,
                          long long workspace_floats, void *stream);
int pmctf_iwave1d_backward(const pmctf_plane_t *l, const pmctf_plane_t *h, const pmctf_iwave_t *p,
                           const pmctf_plane_t *x, int n, int h2, int w, float *workspace,
                           long long workspace_floats, void *stream);

/* LiftingScheme2D.forward_lift_2d / backward_lift_2d: wavelet_transform.py:25-57.
 * x dense [N,1,H,W]; ll, lh, hl, hh dense [N,1,H/2,W/2]; l_out, h_out dense [N,1,H/2,W] (the
 * row-pass outputs the reference returns as 'l','h', may be NULL); workspace_floats >=
 * pmctf_lift2d_workspace(N,H,W) = 2*N*H*W. */
long long pmctf_lift2d_workspace(int N, int H, int W);
int pmctf_lift2d_forward(const float *x, const pmctf_iwave_t *p, float *ll, float *lh, float *hl, float *hh,
                         float *l_out, float *h_out, int N, int H, int W,
                         float *workspace, long long workspace_floats, void *stream);
int pmctf_lift2d_backward(const float *ll, const float *lh, const float *hl, const float *hh,
                          const pmctf_iwave_t *p, float *x, int N, int H, int W,
                          float *workspace, long long workspace_floats, void *stream);
/* backward_lift_2d with dequantize_subbands (pWave.py:191-202) fused into the first loads:
 * ll is divided by ll_div (q_ll at the coarsest level, 1 elsewhere), lh/hl/hh by q. */
int pmctf_lift2d_backward_q(const float *ll, const float *lh, const float *hl, const float *hh,
                            float ll_div, float q, const pmctf_iwave_t *p, float *x, int N, int H, int W,
                            float *workspace, long long workspace_floats, void *stream);

/* quantize_subband (+ RoundNoGradient): pWave.py:184-189,256-257,337; layers.py:71-92.
 * out = [rint](clamp(s*q, +-clip)).  dequantize_subbands: pWave.py:191-202, out = s_hat / q. */
int pmctf_quantize(const float *s, float q, float clip, int lossy, int do_round, float *out,
                   long long n, void *stream);
int pmctf_dequantize(const float *s_hat, float q, int lossy, float *out, long long n, void *stream);

/* quantize_subbands (pWave.py:168-182) on `planes` dense planes of `plane_elems` coefficients with
 * the rate statistics of the symbols accumulated on the fly: stats[2*p] += sum |sym|,
 * stats[2*p+1] += #nonzero (unsigned 64-bit, caller zeroes them).  These exact integer counters are
 * what the GOP-sharded run gathers over NCCL in place of the entropy model's bit estimate
 * (the entropy model itself is out of scope, section 8f). */
int pmctf_quantize_stats(const float *s, float q, float clip, int lossy, float *out, int planes,
                         long long plane_elems, unsigned long long *stats, void *stream);

/* The coder step of a batch of planes with PER-PLANE steps (the H frames of all temporal levels of a GOP coded together;
 * hp_q_scale differs per level, pMCTF_L.py:343-347): sym = rint(clamp(s * q[p], +-clip)).  q_per_plane: DEVICE array of
 * `planes` floats.  out (fp32) receives the symbols, or with dequant != 0 the dequantised values sym / q[p]
 * (dequantize_subbands, pWave.py:191-202).  sym16 (may be NULL): the symbols as int16 at sym16[p * sym16_plane_stride + i]
 * -- what the reference copies to the host for the entropy coder (entropy_models.py:37-40); 8-byte aligned.  stats (may be
 * NULL) as in pmctf_quantize_stats. */
int pmctf_quantize_code(const float *s, const float *q_per_plane, float clip, int lossy, int dequant, float *out, short *sym16,
                        long long sym16_plane_stride, int planes, long long plane_elems, unsigned long long *stats, void *stream);

/* 8-bit planes [n,h0,w0] (HOST-visible layout of one YUV plane batch, already on the device) ->
 * fp32 planes [n,hp,wp], zero padded bottom/right: np_image_to_tensor + F.pad,
 * test_pMCTF_flex.py:151-192 (padding rule: pMCTF/utils/stream_helper.py:23-32). wp % 4 == 0. */
int pmctf_unpack_u8(const unsigned char *src, float *dst, int n, int h0, int w0, int hp, int wp, void *stream);

/* sse[p] += sum over the un-padded h0 x w0 area of (round(clamp(rec,0,255)) - orig)^2: the PSNR
 * numerators of test_pMCTF_flex.py:300-310 as exact integers (caller zeroes sse). */
int pmctf_frame_sse(const float *rec, const unsigned char *orig, int n, int h0, int w0, int hp, int wp,
                    unsigned long long *sse, void *stream);

/* ---- PostProcess: pMCTF/layers/postprocessing.py:20-44, applied to every reconstructed plane at
 * pWave.py:299-300,347,455,526 as dequantModule(x_hat / 256) * 256.  15 3x3 convolutions (1 -> 64, 6 ResBlocks of two
 * 64 -> 64 with LeakyReLU(0.2), 64 -> 64 + skip, 64 -> 1), 958 kFLOP per pixel.  The thirteen 64 -> 64 layers and the last one
 * run as tcgen05 implicit GEMMs with bf16 operands and fp32 accumulators in TMEM; biases, skip connections and the residual
 * stream are fp32.  Not bit-exact (the tensor core's summation order is unspecified): parity is the north-star tolerance for
 * frames, 1e-3 on the [0,1] pixel scale, against the fp32 oracle.
 * Weights: conv1 fp32 OIHW as in the state_dict; every other layer packed once by pmctf_pp_pack_conv (bf16 operand image,
 * pmctf_pp_packed_bytes(co) bytes, 16-byte aligned); biases fp32. */
typedef struct {
    const float *conv1_w, *conv1_b;          /* [64,1,3,3], [64] */
    const void *res_w[12];                   /* resBlocks.{0..5}.{conv1,conv2} packed (co = 64) */
    const float *res_b[12];
    const void *conv2_w; const float *conv2_b; /* packed (co = 64) */
    const void *conv3_w; const float *conv3_b; /* packed (co = 1), [1] */
} pmctf_postprocess_t;
long long pmctf_pp_packed_bytes(int co);
/* OIHW fp32 [co,64,3,3] (co = 64, or 1..16) -> operand image */
int pmctf_pp_pack_conv(const float *w, int co, void *packed, void *stream);
/* Feature-map layouts between the layers (chunk-planar, so that one thread per pixel is coalesced on both sides of the tensor
 * core): bf16 operands [N][8][H][W][8] (channel c of pixel (y,x) at ((n*8 + c/8)*H*W + y*W + x)*8 + c%8), fp32 [N][16][H][W][4].
 * conv1: x [N,1,H,W] fp32 (scaled by in_mul) -> 64-channel feature map, fp32 and bf16 copies */
int pmctf_pp_conv_in(const float *x, const float *w, const float *b, float in_mul, float *out_f32, void *out_bf16, int N, int H, int W,
                     void *stream);
/* fp32 -> bf16 (RN), n a multiple of 4 */
int pmctf_pp_to_bf16(const float *in, void *out, long long n, void *stream);
/* One 3x3 convolution 64 -> co on the tensor cores: in_bf16 [N][8][H][W][8]; co == 64: out = lrelu_slope(conv + bias [+ residual
 * fp32 [N][16][H][W][4]]) written as fp32 [N][16][H][W][4] and / or bf16 [N][8][H][W][8] (either may be NULL); co == 1:
 * y_plane[N,1,H,W] = (x_plane * in_mul + conv + bias) * out_mul (postprocessing.py:41-44 with the scaling of pWave.py:300). */
int pmctf_pp_conv64(const void *in_bf16, const void *packed_w, const float *bias, int co, const float *residual, float lrelu_slope,
                    float *out_f32, void *out_bf16, const float *x_plane, float in_mul, float out_mul, float *y_plane, int N, int H,
                    int W, void *stream);
/* The whole filter, plane by plane: y = PostProcess(x * in_mul) * out_mul on [N,1,H,W]; workspace (256-byte aligned) of
 * pmctf_postprocess_workspace(H, W) bytes holds one plane's feature maps. */
long long pmctf_postprocess_workspace(int H, int W);
int pmctf_postprocess(const float *x, const pmctf_postprocess_t *p, float in_mul, float out_mul, float *y, int N, int H, int W,
                      void *workspace, long long workspace_bytes, void *stream);

/* ---- four-step entropy-parameter network: pMCTF/layers/context_fusion_4step.py:23-249 ----------
 * ContextFusionFourStep(num_features = 112) is evaluated once per coded subband (pWave.py:259-290, 405-421, 500-512): 22 dense
 * 112 -> 112 3x3 convolutions and one 1x1 per subband (4.97 MFLOP per coefficient).  Those run as tcgen05 CTA-pair implicit GEMMs
 * (cta_group::2, M = 256, N = 112, bf16 operands, fp32 accumulators in TMEM, inputs by TMA tensor loads, all weights of a layer
 * resident in shared memory, half per CTA); the small layers around them are CUDA-core kernels.  Feature maps between the layers
 * are chunk-planar: bf16 operands [N][14][H][W][8], fp32 [N][28][H][W][4].  Not bit-exact against fp32 (bf16 operands): the
 * tests report parameter errors and final-symbol mismatch counts against the fp32 oracle. */
long long pmctf_ctx_packed_bytes(int taps);
/* OIHW fp32 [112,112,k,k] (taps = k*k = 9 or 1) -> bf16 operand images of both CTAs of a pair */
int pmctf_ctx_pack_conv(const float *w, int taps, void *packed, void *stream);
/* nn.Conv2d(1 or 2 -> 112, 3x3, padding 1) on single-channel planes x0 (and x1, or NULL): conv1_context (:47), y_spatial_prior_k.0
 * (:62-86); w [112,cin,3,3]; writes the fp32 map and its bf16 operand copy */
int pmctf_ctx_conv_in(const float *x0, const float *x1, const float *w, const float *b, float *out_f32, void *out_bf16, int N, int H, int W,
                      void *stream);
/* One 112 -> 112 convolution on the tensor cores (ContextResidual.conv1 / conv2 :12-14, DepthConv.conv1 with taps = 1):
 * out = lrelu_slope(conv(in) + bias [+ res] [+ res2]) as fp32 and / or bf16 (either may be NULL) */
int pmctf_ctx_conv112(const void *in_bf16, const void *packed_w, int taps, const float *bias, const float *res, const float *res2,
                      float lrelu_slope, float *out_f32, void *out_bf16, int N, int H, int W, void *stream);
/* The same layer with the 112 -> 2 projection that follows it (y_spatial_prior_k_out.2, :66-70) evaluated in the epilogue, where a
 * pixel's 112 output channels sit in one thread's registers: the feature map is never written, only scales / means [N,1,H,W].
 * head_w [2,112,1,1], head_b [2].  Bit-identical to pmctf_ctx_conv112 (fp32 out) followed by pmctf_ctx_head. */
int pmctf_ctx_conv112_head(const void *in_bf16, const void *packed_w, int taps, const float *bias, const float *res, const float *res2,
                           float lrelu_slope, const float *head_w, const float *head_b, float *scales, float *means, int N, int H, int W,
                           void *stream);
/* lower_level_subband (:49-52): nearest x2 upsampling of prev [N,1,h,w] + nn.Conv2d(1 -> 1, 3x3) -> out [N,1,2h,2w] */
int pmctf_ctx_lower_subband(const float *prev, const float *w, const float *b, float *out, int N, int h, int w_, void *stream);
/* DepthConvBlock(112, 2) behind its first 1x1 convolution (pMCTF/layers/video/layers.py:113-172): t1 = LeakyReLU_0.01(conv1(ctx))
 * (a pmctf_ctx_conv112 call with taps = 1) -> depthwise 3x3 -> 1x1 to 2 channels + adaptor(ctx) -> ConvFFN; all weights fp32 in
 * their state_dict layouts.  scales / means: [N,1,H,W] (chunk(2, dim=1) of the block's output, :172) */
typedef struct {
    const float *dw_w, *dw_b;   /* block.0.depth_conv  [112,1,3,3], [112] */
    const float *pw_w, *pw_b;   /* block.0.conv2       [2,112,1,1], [2] */
    const float *ad_w, *ad_b;   /* block.0.adaptor     [2,112,1,1], [2] */
    const float *f1_w, *f1_b;   /* block.1.conv.0      [8,2,1,1], [8] */
    const float *f2_w, *f2_b;   /* block.1.conv.2      [2,8,1,1], [2] */
} pmctf_ctx_dcb_t;
int pmctf_ctx_dcb_tail(const float *t1, const float *ctx_f32, const pmctf_ctx_dcb_t *p, float *scales, float *means, int N, int H, int W,
                       void *stream);
/* the 112 -> 2 1x1 convolution that ends y_spatial_prior_k_out (:66-70) on an fp32 map; w [2,112,1,1] */
int pmctf_ctx_head(const float *feat, const float *w, const float *b, float *scales, float *means, int N, int H, int W, void *stream);
/* process_with_mask (:115-125) of step `step` (0..3, mask = pixels with 2*(y&1) + (x&1) == step): on the mask
 * x_q = rint(x - mean), x_hat = x_q + mean, s_hat = scale, x_res = x - mean, written into the running planes (step 0 zeroes the
 * rest; x_q / x_res may be NULL).  Decoder form (:209-247): x == NULL and dec_sym = the step's decoded symbols (int16, full
 * plane); x == NULL and dec_sym == NULL: nothing but idx16 is written (what the decoder needs BEFORE it can read the step's
 * symbols).  idx16 / sym16 (may be NULL): the step's full planes as GaussianEncoder.encode derives them from the masked planes
 * (entropy_models.py:37-40,266-275) -- symbol and scale-table index on the mask, 0 and the index of scale 1e-5 elsewhere. */
typedef struct {
    const float *x;
    const short *dec_sym;
    const float *scales, *means;
    float *x_hat, *x_q, *s_hat, *x_res;
    short *sym16, *idx16;
    float log_scale_min, log_scale_step;
    int scale_levels, step, lossy, N, H, W;
} pmctf_ctx_step_t;
int pmctf_ctx_mask_step(const pmctf_ctx_step_t *s, void *stream);

/* ---- the LL subband's autoregressive model in its sequential form (third module):
 * pMCTF/layers/context_fusion.py:56-204 (ContextFusionSubband.forward_sequential) as driven by pWave._compress_subband_ar /
 * _decompress_subband_ar (pWave.py:531-584).  One kernel evaluates a coefficient's whole network from five channel-last history
 * planes; the encoder runs the band in one launch, the decoder one launch per coefficient (see csrc/pmctf_llar.cu).
 * Weights are repacked once by pmctf_llar_pack: (cin = 1, taps = 4) maskedConv1 -> [4][128]; (128, 5) the masked 128 -> 128 layers
 * -> [5][128][128] = [tap][ci][co]; (128, 0) a 1x1 layer -> [ci][co].  w_out / b_out: convs.2 in its state_dict layout [2,128,1,1]. */
typedef struct {
    const float *w_in, *b_in;
    const float *w[5], *b[5];       /* residualBlocks.0.conv1, .0.conv2, .1.conv1, .1.conv2, maskedConv2 */
    const float *w1[2], *b1[2];     /* convs.0, convs.1 */
    const float *w_out, *b_out;     /* convs.2 */
    float *Y;                       /* [B][H+2][W+2] reconstructed band, zero border; zeroed by the caller before the first coefficient */
    float *hist[5];                 /* [B][H+2][W+2][128] each, zeroed by the caller */
    int B, H, W;
    float log_scale_min, log_scale_step;
    int scale_levels;
} pmctf_llar_t;
int pmctf_llar_pack(const float *w, int cin, int taps, float *out, void *stream);
/* encoder: yq [B,1,H,W] quantised band -> sym16 / idx16 [B][H*W] (symbol round(round(y) - mean) and scale-table index per
 * coefficient, raster order) and, in p->Y, the band as the decoder will reconstruct it */
int pmctf_llar_encode(const pmctf_llar_t *p, const float *yq, short *sym16, short *idx16, void *stream);
/* The same network on all coefficients of the band at once (the encoder knows them; so does the rate-estimate path,
 * ContextFusionSubband.forward, context_fusion.py:143-158): seven launches, every output computed with exactly the arithmetic of
 * the sequential form, so scales / means / symbols / indexes equal pmctf_llar_encode's bit for bit.  x [B][H][W]: the band
 * (round_in != 0: the quantised band, rounded here as pWave.py:549-553 does).  Outputs (each may be NULL): sym16 / idx16 as
 * pmctf_llar_encode, scales / means [B][H][W] fp32.  The history is SPECULATED to be round(x); *mismatch (device int, zeroed by the
 * caller) is set when some coefficient's reconstruction round(symbol + mean) differs from it -- the caller then runs
 * pmctf_llar_encode, whose history is the reconstruction by construction.  p->Y, p->hist as above (zero borders). */
int pmctf_llar_forward(const pmctf_llar_t *p, const float *x, int round_in, short *sym16, short *idx16, float *scales, float *means,
                       int *mismatch, void *stream);
/* decoder, the whole band in ONE launch: a cluster of eight CTAs per plane keeps the masked layers' weights resident in shared
 * memory (one reduction slice per CTA, partial sums exchanged through distributed shared memory) and decodes the band's symbols
 * with a device-side rANS decoder (rans.cpp:279-331), so there is no host round trip per coefficient.  words / nwords: device copy
 * of the sub-stream's 32-bit words; state: device u64[4] = {rANS state, index of the next word, error flag (out), 0}, as exported by
 * pmctf_rans_decoder_peek and written back with pmctf_rans_decoder_seek; cdfs [cdf_num][cdf_stride] / cdfs_sizes / offsets: DEVICE
 * copies of the tables pmctf_rans_decode_stream takes; out [B][H*W]: the reconstructed band round(symbol + mean) (also left in
 * p->Y).  Same values as pmctf_llar_decode_step + pmctf_rans_decode_stream coefficient by coefficient.  B <= 16. */
int pmctf_llar_decode_band(const pmctf_llar_t *p, const unsigned int *words, long long nwords, unsigned long long *state, const int *cdfs,
                           int cdf_num, int cdf_stride, const int *cdfs_sizes, const int *offsets, float *out, void *stream);
/* decoder: parameters of coefficient `pos` (raster index); prev [B] = reconstructed value of coefficient pos - 1 (ignored for
 * pos == 0), out_mean [B] / out_idx [B]: HOST-visible (mapped pinned) memory read after synchronising the stream */
int pmctf_llar_decode_step(const pmctf_llar_t *p, int pos, const float *prev, float *out_mean, short *out_idx, void *stream);

/* ---- SpyNet motion estimation: pMCTF/layers/video/video_net.py:74-121 ---------------------------
 * pmctf_pair_conv: nn.Conv2d(cin, cout, k, padding = k/2), k = 1, 3 or 7, between channel-chunked bf16 maps as a tcgen05 CTA-pair
 * implicit GEMM (run-time channel counts; the machine of pmctf_ctx_conv112).  in_bf16 [N][cin_pad/8][H][W][8] (cin_pad a multiple of
 * 16, <= 64; channels beyond the layer's cin hold zeros or meet zero weights), weights packed once by pmctf_pair_pack_conv
 * (OIHW fp32 [cout,cin,k,k] -> pmctf_pair_packed_bytes(k, cin_pad, cout_pad) bytes, 16-byte aligned, cout_pad a multiple of 16,
 * <= 128).  out = lrelu_slope(conv + bias [+ add_nchw]) (slope 0 = ReLU, 1 = identity) written as bf16 [N][cout_pad/8][H][W][8]
 * (padding channels zero) and / or fp32 NCHW [N,cout,H,W]; add_nchw (fp32 NCHW) only with out_nchw. */
long long pmctf_pair_packed_bytes(int ks, int cin_pad, int cout_pad);
int pmctf_pair_pack_conv(const float *w, int cout, int cin, int ks, int cin_pad, int cout_pad, void *packed, void *stream);
int pmctf_pair_conv(const void *in_bf16, const void *packed_w, const float *bias, int ks, int cin_pad, int cout, int cout_pad, float slope,
                    void *out_bf16, float *out_nchw, const float *add_nchw, int N, int H, int W, void *stream);
/* One pyramid level's network input (video_net.py:113-119): flow_up [N,2,H,W] = 2 * bilinear x2 upsampling of flow [N,2,H/2,W/2]
 * (F.interpolate, align_corners=False; flow == NULL: zeros, the coarsest level), im2 warped by it (flow_warp, video_net.py:32-55),
 * and the bf16 operand records [im1 (3 ch), warp(im2) (3), flow_up (2), 8 zeros] as [N][2][H][W][8].  im1, im2: [N,3,H,W]. */
int pmctf_spynet_prep(const float *im1, const float *im2, const float *flow, float *flow_up, void *rec_bf16, int N, int H, int W, void *stream);
/* F.avg_pool2d(x, 2, 2) on `planes` planes of H x W (the image pyramid, video_net.py:104-106) */
int pmctf_avgpool2(const float *in, float *out, long long planes, int H, int W, void *stream);

/* Element-wise glue of the coder, one pass each.  pmctf_lstm_gates: the ConvLSTM cell of the long-term context behind its two
 * convolutions (pMCTF/layers/long_context.py:16-34): a = a_in + a_hid, s = sigmoid(a), c_out = s c + s tanh(a), h_out = s tanh(c_out).
 * pmctf_laplace_bits: CompressionModel.get_y_laplace_bits (pMCTF/entropy_models/gaussian_model.py:37-55): per-element
 * max(-log2(cdf(y + .5) - cdf(y - .5) + 1e-5), 0) of a zero-mean Laplace with scale clamp(sigma, 1e-5, 1e10). */
int pmctf_lstm_gates(const float *a_in, const float *a_hid, const float *c, float *h_out, float *c_out, long long n, void *stream);
int pmctf_laplace_bits(const float *y, const float *sigma, float *bits, long long n, void *stream);

/* ---- entropy-coder boundary ----------------------------------------------------------------
 * HOST functions (plain host pointers, no stream): the 64-bit rANS coder of pMCTF/cpp/rans/rans.cpp:76-168,272-331 behind the
 * sub-stream container of pMCTF/cpp/py_rans/py_rans.cpp:22-225 (what the reference binds as MLCodec_rans.RansEncoder /
 * RansDecoder), and pmf_to_quantized_cdf of pMCTF/cpp/ops/ops.cpp:24-82 (MLCodec_CXX).  The rANS primitives are a restatement
 * of rygorous/ryg_rans rans64.h @ c9d162d9 (un-vendored dependency of the reference); streams are byte-identical to the
 * reference's.  cdfs: [cdf_num][cdf_stride] int32 rows, cdfs_sizes[i] valid entries each (the last interval is the escape to
 * 4-bit bypass digits), offsets[i] = value of the first table entry.  Symbols with a negative index are skipped by the encoder.
 * With multi_thread != 0 or stream_part > 1 every sub-stream has its own worker thread: encode / flush return at once and
 * pmctf_rans_encoded_size / pmctf_rans_get_encoded_stream wait for the queued work. */
int pmctf_pmf_to_quantized_cdf(const float *pmf, int n, int precision, unsigned *cdf /* n + 1 */);
int pmctf_rans_encoder_create(int multi_thread, int stream_part, void **enc);
int pmctf_rans_encoder_destroy(void *enc);
int pmctf_rans_encoder_reset(void *enc);
int pmctf_rans_encode_with_indexes(void *enc, const short *symbols, const short *indexes, long long n, const int *cdfs, int cdf_num,
                                   int cdf_stride, const int *cdfs_sizes, const int *offsets);
/* the same as ceil(n / chunk) consecutive pmctf_rans_encode_with_indexes calls of `chunk` symbols each: every chunk is shared out
 * over the sub-streams on its own (py_rans.cpp:45-59), as the reference's per-coefficient encoder.encode calls of the LL band are
 * (pWave.py:548-553) -- what a decoder asking for `chunk` symbols per call (or pmctf_llar_decode_band) expects */
int pmctf_rans_encode_chunked(void *enc, const short *symbols, const short *indexes, long long n, long long chunk, const int *cdfs,
                              int cdf_num, int cdf_stride, const int *cdfs_sizes, const int *offsets);
int pmctf_rans_encoder_flush(void *enc);
long long pmctf_rans_encoded_size(void *enc);
int pmctf_rans_get_encoded_stream(void *enc, unsigned char *out, long long capacity);
int pmctf_rans_decoder_create(int stream_part, void **dec);
int pmctf_rans_decoder_destroy(void *dec);
int pmctf_rans_decoder_set_stream(void *dec, const unsigned char *bytes, long long n);
/* state of one sub-stream's reader, for decoders that continue on the device (pmctf_llar_decode_band): x = rANS state, pos = index
 * of the next unread 32-bit word, words / nwords = the decoder's own copy of the sub-stream (valid until the next set_stream) */
int pmctf_rans_decoder_parts(void *dec);
int pmctf_rans_decoder_peek(void *dec, int part, unsigned long long *x, long long *pos, long long *nwords, const unsigned int **words);
int pmctf_rans_decoder_seek(void *dec, int part, unsigned long long x, long long pos);
int pmctf_rans_decode_stream(void *dec, const short *indexes, long long n, const int *cdfs, int cdf_num, int cdf_stride,
                             const int *cdfs_sizes, const int *offsets, short *out);
/* DEVICE: what entropy_models.py:37-40 and GaussianEncoder.build_indexes (:266-270) do with two blocking copies per coded step,
 * in one pass: sym16 = int16(clamp(symbols, +-30000)), idx16 = int16(clamp((log(max(scales, 1e-5)) - log_scale_min) /
 * log_scale_step, 0, scale_levels - 1)).  symbols / sym16 may both be NULL (decoder side: indexes only).  Inputs 16-byte,
 * outputs 8-byte aligned. */
int pmctf_gaussian_symbolize(const float *symbols, const float *scales, long long n, float log_scale_min, float log_scale_step,
                             int scale_levels, short *sym16, short *idx16, void *stream);

/* Unit test / timing probe of the tcgen05 (5th-generation tensor core) conventions the lifting convolutions are built
 * on: runs `n_ops` kind::i8 MMAs (M = 128, K = 32, s8 x s8 -> s32 in TMEM) per 128-row block on operands copied to
 * shared memory and returns the raw accumulators out[block][128][out_cols].  Offsets are bytes relative to the staged
 * A / B images; rows are 16-byte records at (r%8)*16 + (r/8)*sbo + chunk*lbo (K-major, no swizzle).  No reference
 * counterpart (the reference calls cuDNN).  err receives 1+block if an MMA never completed. */
typedef struct {
    unsigned a_off, a_lbo, a_sbo;
    unsigned b_off, b_lbo, b_sbo;
    unsigned n, d_col, accumulate;
    unsigned a_unsigned; /* 1: the A bytes are unsigned (u8 x s8), 0: signed */
} pmctf_umma_op_t;
int pmctf_umma_selftest(const signed char *A, int a_bytes, const signed char *B, int b_bytes, const pmctf_umma_op_t *ops,
                        int n_ops, int n_blocks, int block_stride_bytes, int out_cols, int *out, int repeat,
                        long long *cycles, int *err, void *stream);

/* ---- training path (BASELINE.json configs[4]): differentiable fp32 primitives composed by autograd ---------------------
 * Under autograd the modules compose these un-fused kernels the way the reference composes conv2d / grid_sample.
 * conv3x3: nn.Conv2d(cin, cout, 3, padding=1) forward (pMCTF/layers/layers.py:54-56), x [N,cin,H,W], w [cout,cin,3,3], b [cout]
 * or NULL, (cin, cout) in {(1,16), (16,16), (16,1), (1,1)}; its data gradient is the same call with w transposed and flipped.
 * conv3x3_wgrad: gw[cout,cin,3,3] += dL/dw, gb[cout] += dL/db (caller zeroes; gb may be NULL).
 * flow_warp_bwd: adjoint of pmctf_flow_warp; gim [N,C,H,W] += dL/dim, gflow [flowN,2,H,W] += dL/dflow (caller zeroes; either
 * may be NULL). */
int pmctf_conv3x3(const float *x, const float *w, const float *b, float *y, int N, int cin, int cout, int H, int W, void *stream);
/* pmctf_conv3x3 with the element-wise step that follows it in PredictUpdate's forward / backward chain (lifting_1d.py:36-49) fused
 * into the epilogue, over [N,cout,H,W] operands: mode 0 y = conv; 1 y = tanh(conv); 2 y = conv, y2 = tanh(conv); 3 y = conv + aux;
 * 4 y = conv * (1 - aux^2) (data gradient through a tanh with output aux); 5 y = conv * (1 - aux^2) + aux2. */
int pmctf_conv3x3_fused(const float *x, const float *w, const float *b, float *y, float *y2, const float *aux, const float *aux2, int mode,
                        int N, int cin, int cout, int H, int W, void *stream);
int pmctf_conv3x3_wgrad(const float *x, const float *g, float *gw, float *gb, int N, int cin, int cout, int H, int W, void *stream);
int pmctf_flow_warp_bwd(const float *gout, const float *im, const float *flow, const float *lin_x, const float *lin_y, float *gim,
                        float *gflow, int N, int C, int H, int W, int flowN, float sign, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PMCTF_B200_H */
