/*
 * mi_b200.h -- C ABI of the B200-native MAML inner-loop hot path of
 * myungsub/meta-interpolation (libmi_b200.so, sm_100a only).
 *
 * Conventions ("Operator/FFI level"):
 *   - every entry point returns 0 on success or a non-zero cudaError_t /
 *     MI_ERR_* code (the reference's native extensions use the same int-error
 *     convention, dain/my_package/FilterInterpolation/filterinterpolation_cuda.cc:18-60);
 *   - the caller owns every buffer; nothing here allocates, frees or
 *     synchronises; every call is asynchronous on `stream`
 *     (the reference launches on torch.cuda.current_stream(),
 *     sepconv/sepconv_op/sepconv.py:276-291);
 *   - activations are NHWC fp32 with an explicit pixel stride `ld*` (floats,
 *     >= channels), which also expresses channel slices of a concat buffer;
 *   - conv weights are [Cout][kh][kw][ldw] ("KRSC", ldw = Cin rounded up to 4,
 *     pad lanes zero); the reference's OIHW nn.Parameter is a permuted VIEW of
 *     the same storage, so no copy exists between the two.
 *
 * Each function cites the reference interface it replaces.
 */
#ifndef MI_B200_H_
#define MI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mi_stream_t; /* cudaStream_t */

enum { MI_OK = 0, MI_ERR_BAD_ARG = 10001, MI_ERR_UNSUPPORTED = 10002, MI_ERR_WORKSPACE = 10003 };

/* activation kinds fused into conv epilogues and their backward masks */
enum { MI_ACT_NONE = 0, MI_ACT_RELU = 1, MI_ACT_LEAKY = 2, MI_ACT_SIGMOID = 3, MI_ACT_TANH = 4 };

/* conv engines: 0 = pick (tcgen05 when the shape is eligible), 1 = force SIMT fp32, 2 = force tcgen05 TF32 */
enum { MI_ENGINE_AUTO = 0, MI_ENGINE_SIMT = 1, MI_ENGINE_TC = 2 };

/* what the weight-gradient finishing stage does with the reduced gradient g */
enum {
    MI_WG_STORE = 0,     /* grad_out = g                                             */
    MI_WG_ACCUM = 1,     /* grad_out += scale * g           (outer meta-gradient)    */
    MI_WG_SGD_SCALAR = 2,/* w_out = w_in - lr[0] * g        (LSLR, inner_loop_optimizers.py:136-147) */
    MI_WG_SGD_TENSOR = 3 /* w_out = w_in - lr[i] * g[i]     (Meta-SGD, inner_loop_optimizers.py:324-332) */
};

int mi_version(void);
const char* mi_error_string(int code);
/* number of kernels launched by this library since load (bench.py "gpu_launches") */
unsigned long long mi_launch_count(void);
/* 1 if the tcgen05 path was compiled in and the current device is sm_100 */
int mi_tc_available(void);
/* CTAs one persistent tensor-core launch may occupy (default and 0: every SM of the device; the environment variable
 * MI_B200_SM_BUDGET overrides both).  The host side narrows it when several task lanes are in flight (one lane per
 * task, meta_learning_system.py:383 loops over the tasks of a meta-batch): launches of different lanes then run side
 * by side on disjoint SMs and each CTA walks more tiles, so per-CTA set-up is paid fewer times per layer.  Grid sizes
 * are fixed when a launch is captured into a CUDA graph.  Returns the value in force. */
int mi_set_sm_budget(int ctas);
/* Pad-lane policy of the tensor-core convolutions.  NHWC rows are padded to a multiple of 4 channels (16-byte rows for
 * TMA); a bulk tensor store clips its innermost dimension at 16-byte granularity, so storing a 51-channel row through
 * TMA also writes lane 51.  Off (default): lanes past `cout` are never written -- they may be the first channels of a
 * neighbouring concat slice -- and a ragged tail of 1-3 channels is stored with scalar accesses.  On: when the output
 * row stride equals cout rounded up to 4, the caller states that those lanes are padding of this very tensor, and they
 * are overwritten with zeros.  The host side switches it on around calls whose output it allocated itself (ops.py).
 * Returns the previous value. */
int mi_set_pad_lanes_scratch(int on);

/* per-launch CUDA-event timing of the hot kernels (bench.py roofline section).  tag: 0 fprop/dgrad tcgen05 per-tap,
 * 1 wgrad tcgen05 per-tap, 2 fprop/dgrad SIMT, 3 wgrad SIMT, 4 sepconv fwd, 5 sepconv bwd, 6 wgrad finish,
 * 7 fprop/dgrad tcgen05 halo (resident weights), 8 fprop/dgrad tcgen05 halo (streamed weights), 9 wgrad tcgen05
 * filter-column, 10 fprop/dgrad tcgen05 with the filter columns stacked along N (the default 3x3 engine).
 * out[4] = {launches, total ms, algorithmic flops, algorithmic bytes}.  Not usable under graph capture. */
int mi_prof_enable(int on);
int mi_prof_summary(int tag, double* out);

/* ------------------------------------------------------------------ convolution
 * Replaces F.conv2d in MetaConv2dLayer.forward (model_utils.py:360) and its
 * autograd backward; stride 1, padding k/2, dilation 1, groups 1 (the only
 * configuration any of the five backbones uses).
 *
 * y[n,oy,ox,co] = act( bias[co] + sum_{ky,kx,ci} x[n,oy+ky-k/2,ox+kx-k/2,ci] * w[co,ky,kx,ci] )
 */
int mi_conv2d_fprop(const float* x, int ldx, const float* w, int ldw, const float* bias,
                    float* y, int ldy, int n, int h, int wd, int cin, int cout, int k,
                    int act, float slope, int engine, mi_stream_t stream);

/* dx = conv_transpose(dy, w) [* act'(mask_y)] ; accumulate!=0 adds into dx.
 * `wt` is the dgrad weight layout produced by mi_weight_to_dgrad:
 * wt[ci][k-1-ky][k-1-kx][co] = w[co][ky][kx][ci]  (ldwt = Cout rounded up to 4).
 * mask_y (optional) is the POST-activation tensor that fed this conv
 * (dx is then the gradient w.r.t. that producer's pre-activation). */
int mi_conv2d_dgrad(const float* dy, int lddy, const float* wt, int ldwt,
                    float* dx, int lddx, const float* mask_y, int ldmask, int mask_act, float mask_slope,
                    int accumulate, int n, int h, int wd, int cin, int cout, int k,
                    int engine, mi_stream_t stream);

/* round_tf32 != 0: the rotated copy is rounded to the TF32 grid (see "TF32 operand convention" below) */
int mi_weight_to_dgrad(const float* w, int ldw, float* wt, int ldwt, int cin, int cout, int k, int round_tf32,
                       mi_stream_t stream);

/* TF32 operand convention of the tensor-core engine (MI_ENGINE_AUTO / MI_ENGINE_TC).  tcgen05 `kind::tf32` reads the
 * upper 19 bits of each fp32 operand, i.e. it TRUNCATES: a relative bias of -2^-11 per operand that accumulates
 * coherently through a deep conv stack (cuDNN's allow_tf32 path, which the reference runs by default, behaves the
 * same way).  To stay within the fp32 tolerance of the reference's exact path this library keeps every operand ON
 * the TF32 grid by round-to-nearest at the producer, where truncation is then a no-op:
 *   - conv outputs (fprop and dgrad, both CUDA engines unless MI_ENGINE_SIMT is forced) are stored rounded;
 *   - weights are read from rounded copies (mi_round_tf32 / wr_out / mi_weight_to_dgrad(round_tf32=1)), the fp32
 *     master copy that the inner and outer updates modify stays exact;
 *   - activations produced by other kernels are rounded in place by the caller before a conv reads them.
 * y[r][0:c] = rn_tf32(x[r][0:c]) for `rows` rows of strides ldx / ldy (y may alias x). */
int mi_round_tf32(const float* x, int ldx, float* y, int ldy, int c, size_t rows, mi_stream_t stream);

/* workspace bytes mi_conv2d_wgrad needs for this shape */
size_t mi_conv2d_wgrad_workspace(int n, int h, int wd, int cin, int cout, int k, int engine);

/* g_w[co,ky,kx,ci] = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n,oy+ky-k/2,ox+kx-k/2,ci];  g_b[co] = sum dy.
 * The finishing stage applies `mode` (MI_WG_*) to weights and bias alike, so
 * the inner-loop update needs no separate elementwise launch
 * (replaces autograd.grad + update_params, meta_learning_system.py:291-307).
 *   grad_w/grad_b : STORE/ACCUM target (may be NULL in the SGD modes)
 *   w_in,b_in -> w_out,b_out : SGD modes (may alias)
 *   lr_w, lr_b : device pointers; scalar (mode 2) or same shape as w / b (mode 3)
 *   gsum_w,gsum_b : optional running sum of g over inner steps (Meta-SGD outer grad of alpha, SURVEY Appx E4)
 *   scale : multiplies g in ACCUM mode.
 *   wt_out, ldwt : optional (SGD modes only): the updated weight is also written in the dgrad layout of
 *                  mi_weight_to_dgrad, so the next inner step needs no rotation launch.
 *   wr_out : optional (SGD modes only): TF32-rounded copy of w_out in the same layout (what the next fprop reads);
 *            when given, wt_out is rounded too. */
int mi_conv2d_wgrad(const float* x, int ldx, const float* dy, int lddy,
                    int n, int h, int wd, int cin, int cout, int k, int ldw,
                    int mode, float scale,
                    float* grad_w, float* grad_b,
                    const float* w_in, const float* b_in, float* w_out, float* b_out,
                    const float* lr_w, const float* lr_b, float* gsum_w, float* gsum_b,
                    float* wt_out, int ldwt, float* wr_out,
                    void* workspace, size_t workspace_bytes, int engine, mi_stream_t stream);

/* Deferred finishing (north_star: "no extra elementwise launch between inner steps").  Between _begin and _flush the
 * finishing stage of every mi_conv2d_wgrad call -- split-K / bias reduction and the store / accumulate / fused
 * inner-loop update of update_params (inner_loop_optimizers.py:136-147, 324-332), including the rotated (`wt_out`) and
 * TF32-rounded (`wr_out`) copies -- is recorded instead of launched; _flush issues one launch per 20 layers on
 * `stream`.  While deferring, every call must be given its OWN workspace, kept alive until the flush, and the outputs
 * of the finishing stage (gradients, updated weights) are not valid before the flush. */
int mi_wgrad_defer_begin(void);
int mi_wgrad_defer_flush(mi_stream_t stream);

/* ------------------------------------------------------------------ pointwise / resampling
 * avg/max pool 2x2 s2  : sepconv/model.py:197-209, voxel_flow.py:243, superslomo/model.py:69, rrin/unet.py:139
 * bilinear x2 upsample : sepconv/model.py:191 (align_corners=True); voxel_flow.py:400, superslomo/model.py:139,
 *                        rrin/unet.py:184 (align_corners=False) */
/* round_tf32 (here and below): store the result rounded to the TF32 grid because a tensor-core conv reads it next
 * (TF32 operand convention above); 0 keeps the exact fp32 value. */
int mi_avgpool2_fwd(const float* x, int ldx, float* y, int ldy, int n, int h, int wd, int c, int round_tf32, mi_stream_t stream);
int mi_avgpool2_bwd(const float* dy, int lddy, float* dx, int lddx, int accumulate, int n, int h, int wd, int c, mi_stream_t stream);
int mi_maxpool2_fwd(const float* x, int ldx, float* y, int ldy, int n, int h, int wd, int c, mi_stream_t stream);
int mi_maxpool2_bwd(const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx, int accumulate,
                    int n, int h, int wd, int c, mi_stream_t stream);
int mi_upsample2_fwd(const float* x, int ldx, float* y, int ldy, int n, int h, int wd, int c, int align_corners, int round_tf32, mi_stream_t stream);
int mi_upsample2_bwd(const float* dy, int lddy, float* dx, int lddx, int accumulate, int n, int h, int wd, int c,
                     int align_corners, int round_tf32, mi_stream_t stream);
/* Region-of-interest forms.  SepConv crops its prediction to the frame (modulePaddingOutput, sepconv/model.py:264-266,
 * 350), so the four filter Subnets (:313-347) are only needed on the part of the padded canvas whose receptive field
 * reaches the surviving window; the Subnet chain is evaluated on that crop.
 *   mi_upsample2_window_*: x holds rows [ly0,ly0+h) x cols [lx0,lx0+wd) of a full_h x full_w grid, y holds rows
 *   [hy0,hy0+oh) x cols [hx0,hx0+ow) of its x2 bilinear upsampling; weights are those of the FULL grid
 *   (nn.Upsample(scale_factor=2, align_corners=True), sepconv/model.py:213-234), so results equal the full
 *   evaluation wherever the sources lie inside x.
 *   mi_window_copy: dst[dy0:dy0+h, dx0:dx0+wd] (+)= src[sy0:sy0+h, sx0:sx0+wd] (crop, and its adjoint). */
int mi_upsample2_window_fwd(const float* x, int ldx, float* y, int ldy, int n, int h, int wd, int c, int align_corners,
                            int full_h, int full_w, int ly0, int lx0, int oh, int ow, int hy0, int hx0,
                            int round_tf32, mi_stream_t stream);
/* mask_y (optional, shaped like dx): the post-activation tensor that was upsampled; dx is then the gradient w.r.t.
 * its pre-activation (dx *= act'(mask_y)), saving the separate mi_act_bwd pass. */
int mi_upsample2_window_bwd(const float* dy, int lddy, float* dx, int lddx, int accumulate, int n, int h, int wd, int c,
                            int align_corners, int full_h, int full_w, int ly0, int lx0, int oh, int ow, int hy0,
                            int hx0, const float* mask_y, int ldmask, int mask_act, float mask_slope, int round_tf32,
                            mi_stream_t stream);
int mi_window_copy(const float* src, int lds, int sh, int sw, int sy0, int sx0, float* dst, int ldd, int dh, int dw,
                   int dy0, int dx0, int n, int h, int wd, int c, int accumulate, mi_stream_t stream);
/* y = a + b  (skip adds, sepconv/model.py:294-309) ; a,b,y may alias */
int mi_add(const float* a, int lda, const float* b, int ldb, float* y, int ldy, size_t pixels, int c, int round_tf32, mi_stream_t stream);
/* dst (+)= src over a channel slice; used for concat/split and gradient fan-in */
int mi_copy(const float* src, int lds, float* dst, int ldd, int accumulate, size_t pixels, int c, mi_stream_t stream);
/* in place: dy *= act'(y) with y the post-activation tensor */
/* round_tf32 != 0: the masked gradient is also rounded to the TF32 grid (it is the next dgrad / wgrad operand) */
int mi_act_bwd(float* dy, int lddy, const float* y, int ldy, int act, float slope, size_t pixels, int c,
               int round_tf32, mi_stream_t stream);
int mi_fill(float* p, float v, size_t count, mi_stream_t stream);

/* ------------------------------------------------------------------ glue of the flow-based backbones
 * Frozen batch norm + activation (nn.BatchNorm2d in eval mode, voxel_flow.py:241-263,352-355):
 *   y = act((x - mean[c]) * rsqrt(var[c] + eps) * gamma[c] + beta[c]).
 * Backward: dx (+)= dz * gamma * inv_std with dz = dy * act'(y); dgamma = sum dz * xhat, dbeta = sum dz
 * (mode MI_WG_STORE or MI_WG_ACCUM with `scale`); dx / dgamma / dbeta may be NULL. */
int mi_bn_eval_fwd(const float* x, int ldx, float* y, int ldy, const float* gamma, const float* beta,
                   const float* mean, const float* var, float eps, int act, float slope, size_t pixels, int c,
                   mi_stream_t stream);
size_t mi_bn_eval_bwd_workspace(size_t pixels, int c);
int mi_bn_eval_bwd(const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx, float* dx, int lddx,
                   int accumulate_dx, const float* gamma, const float* mean, const float* var, float eps, int act,
                   float slope, float* dgamma, float* dbeta, int mode, float scale, void* workspace,
                   size_t workspace_bytes, size_t pixels, int c, mi_stream_t stream);
/* y = a (op) b, op 0 add / 1 sub / 2 mul / 3 div; b has c channels or 1 (broadcast over channels).
 * Replaces the blend / flow arithmetic of voxel_flow.py:480-507, superslomo/model.py:600-640, rrin/model.py:88-120.
 * Backward writes (acc=0) or accumulates (acc=1) ga / gb; either may be NULL. */
int mi_binary_fwd(int op, const float* a, int lda, const float* b, int ldb, int cb, float* y, int ldy, size_t pixels,
                  int c, mi_stream_t stream);
int mi_binary_bwd(int op, const float* a, int lda, const float* b, int ldb, int cb, const float* go, int ldgo,
                  float* ga, int ldga, int acc_a, float* gb, int ldgb, int acc_b, size_t pixels, int c,
                  mi_stream_t stream);
/* y (+)= alpha * x + beta over [pixels][c] */
int mi_affine(const float* x, int ldx, float* y, int ldy, float alpha, float beta, int accumulate, size_t pixels,
              int c, mi_stream_t stream);

/* ------------------------------------------------------------------ heads of the flow / attention backbones
 * y = act(x) standalone (superslomo/model.py:617 sigmoid of one channel, voxel_flow.py:449 tanh) */
int mi_act_fwd(const float* x, int ldx, float* y, int ldy, int act, float slope, size_t pixels, int c, mi_stream_t stream);
/* y = clamp(x, lo, hi) and its gradient (passes where lo <= x <= hi), rrin/model.py:119 */
int mi_clamp_fwd(const float* x, int ldx, float* y, int ldy, float lo, float hi, size_t pixels, int c, mi_stream_t stream);
int mi_clamp_bwd(const float* dy, int lddy, const float* x, int ldx, float* dx, int lddx, int accumulate, float lo,
                 float hi, size_t pixels, int c, mi_stream_t stream);
/* visibility-weighted blend of two warped frames a, b (c channels) with one-channel maps m0, m1:
 *   mode 0: (w0*m0*a + w1*m1*b) / (w0*m0 + w1*m1 + eps)     rrin/model.py:102-103
 *   mode 1: the same with m1 = 1 - m0 (m1 ignored)           superslomo/model.py:621-630
 *   mode 2: m0*a + (1-m0)*b                                  voxel_flow.py:505-507
 * Backward writes / accumulates ga, gb, gm0, gm1 (any may be NULL). */
int mi_blend_fwd(const float* a, int lda, const float* b, int ldb, const float* m0, int ldm0, const float* m1,
                 int ldm1, float* out, int ldo, float w0, float w1, float eps, int mode, size_t pixels, int c,
                 mi_stream_t stream);
int mi_blend_bwd(const float* a, int lda, const float* b, int ldb, const float* m0, int ldm0, const float* m1,
                 int ldm1, const float* go, int ldgo, float* ga, int ldga, float* gb, int ldgb, float* gm0, int ldgm0,
                 float* gm1, int ldgm1, int accumulate, float w0, float w1, float eps, int mode, size_t pixels, int c,
                 mi_stream_t stream);
/* Reflection-padded convolution (MetaConvNorm, model_utils.py:821-849) on the zero-padding conv engine: the
 * activation lives in the interior of an h x wd buffer whose one-pixel ring mi_ring_fix fills in place with zeros
 * (mode 0) or the reflection of the interior (mode 1); mi_ring_fold is the transpose (ring gradient folded onto the
 * mirrored interior pixel for mode 1, ring cleared in both modes). */
int mi_ring_fix(float* x, int ld, int n, int h, int wd, int c, int mode, mi_stream_t stream);
int mi_ring_fold(float* g, int ld, int n, int h, int wd, int c, int mode, mi_stream_t stream);
/* CAIN input/output (cain/model.py:70-94, model_utils.py:11-28,202-217): per-plane mean of an NCHW tensor;
 * space-to-depth by r of the two reflect-padded, mean-shifted frames into one ringed NHWC buffer
 * [n, oh+2, ow+2, 6*r*r] (ring = 0); depth-to-space of a ringed NHWC buffer [n, ih+2, iw+2, 3*r*r] into the cropped
 * NCHW image plus (mean0+mean1)/2, and its transpose. */
int mi_channel_mean_nchw(const float* f, float* out, int planes, int hw, mi_stream_t stream);
int mi_space_to_depth(const float* f0, const float* f1, const float* mean0, const float* mean1, float* out, int ldo,
                      int n, int h, int wd, int pad_top, int pad_left, int oh, int ow, int r, mi_stream_t stream);
int mi_depth_to_space(const float* in, int ldi, const float* mean0, const float* mean1, float* out, int n, int h,
                      int wd, int pad_top, int pad_left, int ih, int iw, int r, mi_stream_t stream);
int mi_depth_to_space_bwd(const float* gout, float* gin, int ldi, int n, int h, int wd, int pad_top, int pad_left,
                          int ih, int iw, int r, mi_stream_t stream);
/* Channel attention (MetaCALayer, model_utils.py:931-955) on a buffer with `ring` border pixels excluded:
 * out[n,c] = scale * sum_interior x (* mul);  out = o*s[n,c] + res;  dx (+)= g*s[n,c];  dx[interior] += dy[n,c]*scale */
int mi_interior_reduce(const float* x, int ldx, const float* mul, int ldm, float* out, int n, int h, int wd, int c,
                       int ring, float scale, mi_stream_t stream);
int mi_scale_add(const float* o, int ldo, const float* s, const float* res, int ldr, float* out, int ldout, int n,
                 size_t pixels_per_image, int c, mi_stream_t stream);
int mi_scale_bwd(const float* g, int ldg, const float* s, float* dx, int lddx, int accumulate, int n,
                 size_t pixels_per_image, int c, mi_stream_t stream);
int mi_interior_bcast_add(const float* dy, float* dx, int lddx, int n, int h, int wd, int c, int ring, float scale,
                          mi_stream_t stream);

/* ------------------------------------------------------------------ frames in / prediction out
 * Builds the NHWC network input from two NCHW frames with the reference's
 * padding folded in: canvas[n,y,x,0:3]=f0, [3:6]=f1 sampled at
 * (clamp|reflect)(y - pad_top, x - pad_left).  mode 0 = replicate
 * (sepconv/model.py:254-269), 1 = reflect (model_utils.py:17-28, voxel_flow.py:360-368). */
int mi_frames_to_canvas(const float* f0, const float* f1, float* canvas, int ldc,
                        int n, int h, int wd, int ch, int cw, int pad_top, int pad_left, int mode, int round_tf32,
                        mi_stream_t stream);
/* NHWC window -> NCHW [n,c,h,w] (the crop of sepconv/model.py:349) and back (gradient of the crop, zero elsewhere is the caller's fill) */
int mi_nhwc_window_to_nchw(const float* src, int lds, float* dst, int n, int hs, int ws, int y0, int x0, int h, int wd, int c, mi_stream_t stream);
int mi_nchw_to_nhwc_window(const float* src, float* dst, int ldd, int n, int hs, int ws, int y0, int x0, int h, int wd, int c, mi_stream_t stream);

/* ------------------------------------------------------------------ adaptive separable convolution
 * Replaces FunctionSepconv (sepconv/sepconv_op/sepconv.py:247-380; kernels :5-30, :138-190).
 *   frame : NCHW [n,c,fh,fw] fp32 (c <= 4)
 *   vert, horiz : NHWC [n,gh,gw,ldf] with F taps per pixel (reference layout is [n,F,gh,gw])
 *   out   : NCHW [n,c,oh,ow] -- output pixel (i,j) uses vert/horiz at grid (gy0+i, gx0+j) and the FxF window whose
 *           top-left input sample is frame[clamp(i+iy0+fy), clamp(j+ix0+fx)] (replicate border).
 * With gy0=gx0=0, iy0=ix0=0, fh=oh+F-1 this is exactly the reference op on a pre-padded input; with
 * gy0=gx0=25, iy0=ix0=-25 on the raw frame it is the op fused with modulePaddingInput, modulePad and
 * modulePaddingOutput (SURVEY Appx A.1).  No gradient w.r.t. frame is ever needed (sepconv.py:319). */
/* Tap-planar workspace (optional, F = 51, c = 3): the kernels that keep four pixels per thread read the filters
 * as [image][tap][oh][ow] (one cache line per tap per 32 pixels instead of 32 lines).  `planar` of
 * mi_sepconv_planar_bytes(n, oh, ow, taps) bytes is filled by the forward (one transposing pass) and can be handed to
 * the backward of the same call (planar_valid = 1), which also needs `planar_grad` of the same size as scratch for
 * the planar gradients before they are returned to NHWC.  NULL selects the kernels that read NHWC filters directly. */
size_t mi_sepconv_planar_bytes(int n, int oh, int ow, int taps);
int mi_sepconv_fwd(const float* frame, const float* vert, const float* horiz, int ldf, float* out,
                   int n, int c, int fh, int fw, int gh, int gw, int oh, int ow,
                   int gy0, int gx0, int iy0, int ix0, int taps, float* planar, mi_stream_t stream);
/* g_vert/g_horiz get the gradient inside the window; the caller zero-fills the rest of the grid -- or sets
 * MI_SEPCONV_ZERO_OUTSIDE in `round_tf32` (bit 0 of which is the rounding request) and the call defines every pixel of
 * the [n,gh,gw] grids itself: zero outside the window (and in the pad lanes ldg - taps of its rows), in the launch that
 * writes the window. */
#define MI_SEPCONV_ZERO_OUTSIDE 2
int mi_sepconv_bwd(const float* frame, const float* vert, const float* horiz, int ldf, const float* grad_out,
                   float* g_vert, float* g_horiz, int ldg,
                   int n, int c, int fh, int fw, int gh, int gw, int oh, int ow,
                   int gy0, int gx0, int iy0, int ix0, int taps, int round_tf32,
                   float* planar, int planar_valid, float* planar_grad, mi_stream_t stream);

/* ------------------------------------------------------------------ bilinear backward warp (grid_sample)
 * variant 0: superslomo backWarp / rrin warp (superslomo/model.py:292-302, rrin/model.py:8-21):
 *            sample at (x+u-0.5, y+v-0.5), zeros outside (SURVEY Appx E2)
 * variant 1: voxelflow (voxel_flow.py:471-503): sample at clip(x + sx*u*(W-1)/2... see DESIGN.md), border clamp.
 *   img : NHWC [n,h,w,c] (data, no gradient), flow : NHWC [n,h,w,2] (u,v), out : NHWC [n,h,w,c] */
int mi_warp_fwd(const float* img, int ldi, const float* flow, int ldfl, float* out, int ldo,
                int n, int h, int wd, int c, int variant, float flow_scale_x, float flow_scale_y, mi_stream_t stream);
int mi_warp_bwd(const float* img, int ldi, const float* flow, int ldfl, const float* grad_out, int ldgo,
                float* grad_flow, int ldgf, float* grad_img, int ldgi, int accumulate,
                int n, int h, int wd, int c, int variant, float flow_scale_x, float flow_scale_y, mi_stream_t stream);

/* ------------------------------------------------------------------ losses and metrics
 * L1 / MSE mean (loss.py:287-290): loss_out[0] += weight * mean(f(pred-target)); grad = weight * f'(.)/count.
 * pred, target NCHW contiguous of `count` elements. kind 0 = L1, 1 = MSE. grad may be NULL. */
int mi_loss_fwd_bwd(const float* pred, const float* target, float* grad, float* loss_out, size_t count,
                    int kind, float weight, mi_stream_t stream);
/* utils.py:171-204: sum over elements of ((q(p)-q(t))/255)^2 into sq_out[0] (double); psnr = -10 log10(sq/count + 1e-8) on host */
int mi_psnr_accumulate(const float* pred, const float* target, double* sq_out, size_t count, mi_stream_t stream);
/* utils.py:195-204 -> pytorch_msssim/__init__.py:19-75 (ssim, size_average, val_range): pred/target NCHW [c,h,w] in
 * [0,1], quantised to 8 bits in the kernel; `window_host` = the normalised 1-D Gaussian of `win` (<= 11) taps (HOST
 * pointer, read at call time); sum_out[0] += sum of the SSIM map over c x (h-win+1) x (w-win+1); mean on the host. */
int mi_ssim_accumulate(const float* pred, const float* target, double* sum_out, int c, int h, int w,
                       const float* window_host, int win, float val_range, mi_stream_t stream);

/* ------------------------------------------------------------------ inner-loop rules on flat arenas
 * (inner_loop_optimizers.py:150-244, 335-426).  `seg` maps each 1024-float chunk of the arena to a tensor id;
 * lr is indexed [seg*lr_stride + num_step] (LSLR) or per element (Meta-SGD, lr_per_element=1).
 * rule 0 = SGD, 1 = Adam, 2 = Adamax(LSLR quirk: exp_avg persists, exp_inf stateless),
 * 3 = Adamax(Meta-SGD quirk: fully stateless).  skip[seg]!=0 leaves w_out = w_in (None gradient). */
int mi_inner_update(const float* w_in, const float* g, float* w_out, float* exp_avg, float* exp_avg_sq,
                    const float* lr, int lr_per_element, int lr_stride, int num_step,
                    const int32_t* seg, const uint8_t* skip, size_t count, int rule, int step_count, mi_stream_t stream);
/* outer optimizer (meta_learning_system.py:132-143): kind 0 = SGD, 1 = Adam(amsgrad off), 2 = Adamax */
int mi_outer_step(float* p, const float* g, float* m, float* v, size_t count, int kind, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, mi_stream_t stream);
/* y[i] = a*x[i] + b*y[i] */
int mi_axpby(const float* x, float a, float* y, float b, size_t count, mi_stream_t stream);
/* y[i] += a * x1[i] * x2[i]  (Meta-SGD alpha gradient  -gsum (.) G, SURVEY Appx E4) */
int mi_addcmul(float* y, float a, const float* x1, const float* x2, size_t count, mi_stream_t stream);
/* out[t] = sum over tensor t of a[i]*b[i]  (lr / gamma outer gradients, SURVEY Appx E4); segs as above; out zeroed by
 * caller.  b == NULL sums a alone (per-tensor mean of the support gradient = the L2F task embedding,
 * meta_learning_system.py:231-255). */
int mi_segment_dot(const float* a, const float* b, const int32_t* seg, float* out, size_t count, mi_stream_t stream);
/* y[i] (+)= alpha * s(t) * x[i] with s(t) = scale[t] where mask[t] != 0 (mask == NULL: everywhere), else 1.
 * L2F attenuation theta_i <- gamma_i * theta_i (meta_learning_system.py:258-272) and its outer gradient
 * dL/dtheta_i += gamma_i * G_i on the flat arena. */
int mi_segment_scale(const float* x, const float* scale, const int32_t* seg, const float* mask, float* y, float alpha,
                     int accumulate, size_t count, mi_stream_t stream);

/* ------------------------------------------------------------------ input staging
 * data/vimeo_septuplet.py:50-78 in one launch: per-task random crop (:56-61), temporal flip (:64-66), BGR->RGB (:69),
 * HWC uint8 -> CHW float (/255 unless voxelflow, :72-75), per-channel (x - mean) / std (:77-78).
 *   src      : device [tasks][frames][src_h][src_w][3] uint8, as cv2.imread decodes (bgr=1) or RGB (bgr=0)
 *   dst      : device [frames][tasks][3][h][w] float -- frame f of the meta-batch is dst[f] ([tasks,3,h,w], NCHW)
 *   y0, x0   : device int32 [tasks] crop origins (y0+h <= src_h, x0+w <= src_w: the caller checks, it drew them)
 *   reversed : device uint8 [tasks], non-zero = frames in reverse temporal order
 *   mean3, std3 : HOST float[3] or both NULL (no normalisation) */
int mi_septuplet_prepare(const uint8_t* src, float* dst, const int32_t* y0, const int32_t* x0,
                         const uint8_t* reversed, int tasks, int frames, int src_h, int src_w, int h, int w,
                         int bgr, int div255, const float* mean3, const float* std3, mi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MI_B200_H_ */
