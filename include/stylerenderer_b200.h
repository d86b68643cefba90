/* stylerenderer_b200 -- C ABI of the B200-native (sm_100a) StyleRenderer hot path.
 *
 * This is the drop-in boundary: every entry point replaces one C/C++ symbol that sits under the
 * reference's pybind shims.  Plain pointers and sizes only, no torch types.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers owned by the caller (outputs pre-allocated, inputs
 *     never modified); `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function is asynchronous w.r.t. the host and returns 0 on success, a negative
 *     SR_ERR_* code for argument errors, or a positive cudaError_t for launch errors;
 *     sr_last_error() returns a thread-local human readable message for the last failure;
 *   - functions are stateless and re-entrant (the reference ops are called from the forward
 *     thread and from autograd's backward thread).
 */
#ifndef STYLERENDERER_B200_H_
#define STYLERENDERER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SR_OK 0
#define SR_ERR_INVALID_ARGUMENT (-1)
#define SR_ERR_UNSUPPORTED (-2)
#define SR_ERR_DRIVER (-3)

/* ABI version (bumped on any signature change) and last error text. */
int sr_abi_version(void);   /* currently 3 */
const char *sr_last_error(void);
/* Number of kernel launches issued by this library in this process (bench.py's gpu_launches). */
int64_t sr_launch_count(void);

/* ------------------------------------------------------------------ fused bias + leaky-ReLU ----
 * Replaces `bool fused_bias_act_op(float*,const float*,const float*,const float*,int,int,float,
 * float,int,int,int,int,int,int)` (reference op/fused_bias_act.cpp:3-4, kernel
 * op/fused_bias_act_kernel.cu:14-42).
 *   t = x[i] + (bias ? bias[(i / step_b) % size_b] : 0);  r = ref ? ref[i] : 0
 *   act*10+grad: 30 -> t>0 ? t : alpha*t;  31 -> r>0 ? t : alpha*t;  32 -> 0;  10/11 -> t;  12 -> 0
 *   y[i] = that * scale
 * bias / ref may be NULL (the reference's "empty tensor").  NCHW: step_b = H*W, size_b = C.
 * channels-last or [B,C]: step_b = 1, size_b = C.  size_x is 64-bit (the reference caps at 2^31). */
int sr_fused_bias_act_f32(float *y, const float *x, const float *bias, const float *ref,
                          int act, int grad, float alpha, float scale,
                          int64_t size_x, int64_t step_b, int64_t size_b, void *stream);

/* Backward of the leaky-ReLU with the bias gradient fused into the same pass (the reference makes a
 * second full pass: op/fused_act.py:27-38):
 *   dx[i] = scale * (y[i] > 0 ? gy[i] : alpha*gy[i]);   dbias[c] = sum_{i in channel c} dx[i]
 * dbias (size_b floats) is zeroed by the call; pass dbias = NULL to skip the reduction. */
int sr_fused_lrelu_backward_f32(float *dx, float *dbias, const float *gy, const float *y,
                                float alpha, float scale,
                                int64_t size_x, int64_t step_b, int64_t size_b, void *stream);

/* ------------------------------------------------------------------ upfirdn2d ------------------
 * Replaces `bool upfirdn2d_op(float*, const float*, const float*, UpFirDn2DKernelParams&, int, int)`
 * (reference op/upfirdn2d.cpp:2-23, kernels op/upfirdn2d_kernel.cu:79-257): zero-stuff by `up`, pad
 * (negative pads crop), TRUE convolution with taps[kh][kw], keep every `down`-th sample.
 *   x:   [major, in_h, in_w, minor]  (NCHW planes: major = N*C, minor = 1)
 *   out: [major, out_h, out_w, minor], out = (in*up + pad0 + pad1 - k) / down + 1 (floor)
 * Any up/down >= 1 and any tap count are accepted (specialised kernels cover the model's
 * 4x4 up1/down1, up2, down2 cases; everything else takes the generic kernel). */
int sr_upfirdn2d_f32(float *out, const float *x, const float *taps,
                     int64_t major, int64_t in_h, int64_t in_w, int64_t minor,
                     int kernel_h, int kernel_w, int up_x, int up_y, int down_x, int down_y,
                     int pad_x0, int pad_x1, int pad_y0, int pad_y1, void *stream);

/* ------------------------------------------------------------------ rasterizer -----------------
 * Replace `rasterize_gpu<scalar,int64_t>` / `rasterize_gpu_backward<scalar,int64_t>` (reference
 * op/rasterize.cpp:14-19, kernels op/rasterize.cu:40-138, math op/rasterize.h:9-228).
 *
 * Forward: z-buffer rasterisation of `nf` triangles over `nv` vertices into h x w (h == w required:
 * the reference swaps the two).
 *   verts [b,nv,3] (or [nv,3] when shared_v), tris int64 [nf,3] when shared_f else [b,nf,3]
 *   ids   int64 [b,h,w,3]  = winning triangle's vertex ids (+ nv*batch unless shared_v), 0 = background
 *   bary  [b,h,w,3]        = normalised barycentric coefficients, 0 = background
 *   keys  workspace of sr_rasterize_workspace_bytes() bytes (16-byte aligned), contents undefined on return:
 *         uint64 [b,h,w] depth/triangle keys, then (float path) the products of the per-vertex / per-triangle
 *         pre-pass -- pixel-space vertices [b,nv,3] and the triangle list narrowed to int32 [b,nf,3]
 * Deterministic: among equal depths the first triangle in list order wins, like the reference's
 * CPU loop (the reference CUDA kernel is racy).  Bit-exact with that CPU loop for ids and bary.
 * Optional fused attribute interpolation (reference op/rasterize.py:29-37): when tex != NULL,
 *   out[b,y,x,:] = sum_k bary_k * tex[ids_k, :],  tex [b*nv, c] (row index = ids), out [b,h,w,c]. */
int sr_rasterize_forward_f32(int64_t b, int64_t nv, int64_t nf, int64_t h, int64_t w,
                             int shared_v, int shared_f, int perspective,
                             const float *verts, const int64_t *tris,
                             int64_t *ids, float *bary, uint64_t *keys, float eps,
                             const float *tex, int64_t c, float *out, void *stream);
int sr_rasterize_forward_f64(int64_t b, int64_t nv, int64_t nf, int64_t h, int64_t w,
                             int shared_v, int shared_f, int perspective,
                             const double *verts, const int64_t *tris,
                             int64_t *ids, double *bary, uint64_t *keys, double eps,
                             const double *tex, int64_t c, double *out, void *stream);
/* size in bytes of the `keys` workspace for a given problem (nv, nf: vertices per image, triangles per list) */
int64_t sr_rasterize_workspace_bytes(int64_t b, int64_t nv, int64_t nf, int64_t h, int64_t w, int is_f64);

/* Reference-shaped backward (`rasterize.backward`, op/rasterize.cpp:179-241): dcoeff [b,h,w,3,9]
 * = d bary_i / d vertex_k.{x,y,z}; pixels whose three ids are not distinct are left untouched
 * (the caller zero-fills, like the reference). n = vertices per image. */
int sr_rasterize_dcoeff_f32(int64_t b, int64_t n, int64_t h, int64_t w, int perspective,
                            const float *verts, const int64_t *ids, float *dcoeff, float eps, void *stream);
int sr_rasterize_dcoeff_f64(int64_t b, int64_t n, int64_t h, int64_t w, int perspective,
                            const double *verts, const int64_t *ids, double *dcoeff, double eps, void *stream);

/* Fused backward of the whole `Rasterize` autograd function (reference op/rasterize.py:39-80, which
 * materialises dcoeff and scatter-adds through a host-built sparse matrix):
 *   grad_verts[ids_k,:] += sum_i (sum_c gout_c * tex[ids_i,c]) * dcoeff[i][k][:]
 *   grad_tex[ids_k,c]   += gout_c * bary_k
 * by atomic accumulation; grad_verts [b*n,3] / grad_tex [b*n,c] must be zero-filled by the caller;
 * either may be NULL.  Summation order is not deterministic (float atomics). */
int sr_rasterize_backward_f32(int64_t b, int64_t n, int64_t h, int64_t w, int64_t c, int perspective,
                              const float *verts, const float *tex, const int64_t *ids, const float *bary,
                              const float *gout, float *grad_verts, float *grad_tex, float eps, void *stream);
int sr_rasterize_backward_f64(int64_t b, int64_t n, int64_t h, int64_t w, int64_t c, int perspective,
                              const double *verts, const double *tex, const int64_t *ids, const double *bary,
                              const double *gout, double *grad_verts, double *grad_tex, double eps, void *stream);

/* Resolution pyramid: the SAME mesh rasterised at several square sizes in one triangle pass + one resolve pass
 * (forward) and one scatter pass (backward).  Replaces the seven independent `rasterize(..., h)` calls of
 * GeneratorWithMap.forward (reference model.py:260-270: sizes 4, 8, ..., 256) -- each triangle's indices and vertices are
 * loaded once, the per-size set-up arithmetic is the reference's, so every level is bit-identical to a single-size
 * `sr_rasterize_forward_f32` call (and to `rasterize_cpu`).  `levels` is a HOST array; pointers inside are device
 * pointers.  keys: workspace of `sr_rasterize_pyramid_workspace_bytes` bytes.
 * Backward: levels with gout == NULL take no part; all others scatter-add into the same zero-filled
 * grad_verts [b*n,3] / grad_tex [b*n,c] (either may be NULL). */
#define SR_RASTER_MAX_LEVELS 8
typedef struct sr_raster_level {
    int64_t size;          /* h == w of this level */
    int64_t *ids;          /* [b,size,size,3] */
    void *bary;            /* float [b,size,size,3] */
    void *out;             /* float [b,size,size,c] (forward, required when tex != NULL) */
    const void *gout;      /* float [b,size,size,c] (backward: gradient of out) or NULL */
} sr_raster_level;
int64_t sr_rasterize_pyramid_workspace_bytes(int64_t b, int64_t nv, int64_t nf, int n_levels, const int64_t *sizes);
int sr_rasterize_pyramid_forward_f32(int64_t b, int64_t nv, int64_t nf, int n_levels, const sr_raster_level *levels,
                                     int shared_v, int shared_f, int perspective,
                                     const float *verts, const int64_t *tris, uint64_t *keys, float eps,
                                     const float *tex, int64_t c, void *stream);
/* Forward-only form: only the interpolated maps are produced.  levels[i].ids / .bary may be NULL (then not written);
 * planar != 0 writes each map as [b, c, size, size] planes (NCHW, what the convolution stack consumes) instead of
 * [b, size, size, c].  Values are those of sr_rasterize_pyramid_forward_f32. */
int sr_rasterize_pyramid_maps_f32(int64_t b, int64_t nv, int64_t nf, int n_levels, const sr_raster_level *levels,
                                  int shared_v, int shared_f, int perspective, const float *verts, const int64_t *tris,
                                  uint64_t *keys, float eps, const float *tex, int64_t c, int planar, void *stream);
int sr_rasterize_pyramid_backward_f32(int64_t b, int64_t n, int n_levels, const sr_raster_level *levels, int64_t c,
                                      int perspective, const float *verts, const float *tex,
                                      float *grad_verts, float *grad_tex, float eps, void *stream);

/* ------------------------------------------------------------------ modulated convolution ------
 * Replaces the dense contraction inside ModulatedConv2d.forward (reference layers.py:293-323: cuDNN grouped
 * conv2d / conv_transpose2d over B per-sample weight copies) with ONE shared-weight implicit GEMM on the
 * tcgen05 tensor cores (tf32 x tf32 -> fp32):
 *   D[n, gy, gx, co] = sum_{t < num_taps} sum_{ci} in[n, gy*in_stride + tap_dy[t], gx*in_stride + tap_dx[t], ci]
 *                                                   * weight[co, tap_w[t], ci]            (zero outside the input)
 * for every point (gy, gx) of a grid_h x grid_w lattice; D is written to
 *   out[n, out_y0 + gy*out_stride, out_x0 + gx*out_stride, co]      (NHWC, [batch, out_h, out_w, cout]).
 * Layouts: in [batch, in_h, in_w, cin] NHWC fp32; weight [cout][taps_total][cin] fp32 (see
 * sr_conv_weight_prep_tf32).  cin % 32 == 0, cout % 128 == 0, all tensors 16-byte aligned.
 * Epilogues:
 *   0: out = D * rowscale[n, co]                                  (rowscale may be NULL)
 *   1: t = D * rowscale[n,co] (* stylemap[n,0,y,x] + stylemap[n,1,y,x]) + noise_weight[0]*noise[n,y,x] + bias[co]
 *      out = (t > 0 ? t : alpha*t) * gain          -- the StyledConv / StyledMapConv tail (reference model.py:26-32,48-55)
 *   2: as 1, but `out` receives D * rowscale[n,co] -- the demodulated conv output BEFORE the map affine / noise / bias /
 *      activation, which is what the StyledMapConv backward needs (sr_styled_bwd_prologue3_f32 rebuilds the activated
 *      value from it; a map that is exactly 0 would make it unrecoverable from the activated one); out2 and the fused
 *      ToRGB still see the activated value y.
 *   in every case, if out2 != NULL: out2 = tf32_round(y * scale2[n, co]) (next layer's modulated input; y = out for 0/1).
 * noise is planar [*, out_h, out_w] with batch stride noise_batch_stride (0 broadcasts one plane);
 * stylemap is planar with two planes of out_h*out_w per image and batch stride stylemap_batch_stride. */
typedef struct sr_conv_args {
    const float *in;
    int64_t batch, in_h, in_w, cin;
    const float *weight;
    int64_t cout, taps_total;
    int32_t num_taps;
    int32_t tap_dy[9], tap_dx[9], tap_w[9];
    int32_t in_stride;
    int64_t grid_h, grid_w, out_h, out_w;
    int32_t out_stride, out_y0, out_x0;
    float *out;
    float *out2;
    int32_t epilogue;
    const float *rowscale, *scale2, *bias, *noise, *noise_weight, *stylemap;
    int64_t noise_batch_stride, stylemap_batch_stride;
    float alpha, gain;
    /* fused ToRGB (reference model.py:56-69): rgb_out[n,y,x,k] = sum_co out[n,y,x,co] * rgb_weight[n,k,co], k < 3;
     * rgb_weight [batch,3,cout] are the per-sample modulated 1x1 weights, rgb_out [batch,out_h,out_w,3] is zeroed by
     * the call.  Both NULL = off. */
    const float *rgb_weight;
    float *rgb_out;
} sr_conv_args;
int sr_conv_igemm_tf32(const sr_conv_args *args, void *stream);
/* `count` (<= 4) such contractions that share every tensor, stride and epilogue and differ only in tap list,
 * lattice size and lattice origin -- the four output-parity classes of the stride-2 transposed convolution
 * (reference layers.py:301-309) -- executed as ONE persistent launch. */
int sr_conv_igemm_multi_tf32(const sr_conv_args *args, int count, void *stream);

/* Weight gradient of the same contraction on the tensor cores (replaces cuDNN's grouped wgrad under
 * ModulatedConv2d's backward): for every tap t < num_taps
 *   dw[co, tap_out[t], ci] (+)= sum_{n, gy, gx} g[n, gy*g_stride + g_dy[t], gx*g_stride + g_dx[t], co]
 *                                            * x[n, gy*x_stride + x_dy[t], gx*x_stride + x_dx[t], ci]
 * over a grid_h x grid_w lattice (zero outside either tensor).  g [batch, g_h, g_w, cout] and
 * x [batch, x_h, x_w, cin] are NHWC fp32 (pre-rounded to tf32), dw is [cout][taps_total][cin] fp32.
 * cin % 128 == 0, cout % 128 == 0.  Split-K with fp32 atomics: zero_init != 0 clears dw first. */
typedef struct sr_wgrad_args {
    const float *g;
    int64_t batch, g_h, g_w, cout;
    const float *x;
    int64_t x_h, x_w, cin;
    int64_t grid_h, grid_w;
    int32_t num_taps;
    int32_t g_dy[9], g_dx[9], x_dy[9], x_dx[9], tap_out[9];
    int32_t g_stride, x_stride;
    float *dw;
    int64_t taps_total;
    int32_t zero_init;
} sr_wgrad_args;
int sr_conv_wgrad_tf32(const sr_wgrad_args *args, void *stream);

/* xs[n,p,c] = tf32_round(x[n,p,c] * style[n,c]) for an NHWC tensor (style == NULL: rounding only). */
int sr_modulate_tf32(float *xs, const float *x, const float *style, int64_t batch, int64_t pixels, int64_t channels,
                     void *stream);

/* Re-layout + scale + tf32-round a reference-layout weight [cout, cin, kh, kw] into a GEMM B operand
 * [rows][kh*kw][cols]:  transpose 0/3: rows = cout, cols = cin, tap = ky*kw+kx;  1: rows = cin, cols = cout,
 * taps flipped (dgrad of the plain conv);  2: rows = cin, cols = cout, taps as is (transposed-conv forward). */
int sr_conv_weight_prep_tf32(float *dst, const float *w, float scale, int64_t cout, int64_t cin, int kh, int kw,
                             int transpose, void *stream);

/* ------------------------------------------------------------------ fused passes of a StyledConv block (NHWC) ---
 * 4x4 FIR (up = down = 1, same pad on both axes) over an NHWC tensor with the StyledConv tail fused in
 * (reference layers.py:310 Blur -> model.py:28-31 NoiseInjection + FusedLeakyReLU):
 *   out[n,y,x,c] = lrelu( fir(x)[n,y,x,c] + noise_weight[0]*noise[n,y,x] + bias[c] ) * gain
 * noise planar [*, out_h, out_w] with batch stride noise_batch_stride (0 = broadcast), may be NULL. */
int sr_blur_nhwc_styled_f32(float *out, const float *x, const float *taps, int64_t batch, int64_t in_h, int64_t in_w,
                            int64_t channels, int pad0, int pad1, const float *noise, int64_t noise_batch_stride,
                            const float *noise_weight, const float *bias, float alpha, float gain, void *stream);
/* same, plus a second output out2 = tf32_round(out * scale2[n,c]) (the next layer's modulated GEMM operand). */
int sr_blur_nhwc_styled2_f32(float *out, float *out2, const float *scale2, const float *x, const float *taps, int64_t batch,
                             int64_t in_h, int64_t in_w, int64_t channels, int pad0, int pad1, const float *noise,
                             int64_t noise_batch_stride, const float *noise_weight, const float *bias, float alpha,
                             float gain, void *stream);

/* The transpose-side FIR of the up-sampling block's backward with its tail fused:
 *   f = fir(x) (4x4, up = down = 1, pad (pad0,pad1));  out = tf32_round(f * scale[n,c]);  dot[n,c] = sum_p f * other[n,p,c]
 * (other has the output's shape; dot is zeroed by the call). */
int sr_blur_nhwc_scaledot_f32(float *out, float *dot, const float *x, const float *taps, const float *scale,
                              const float *other, int64_t batch, int64_t in_h, int64_t in_w, int64_t channels, int pad0,
                              int pad1, void *stream);

/* Backward prologue of a StyledConv block, one pass over (gy, y) [batch, pixels, channels]:
 *   g_pre = gain * (y > 0 ? gy : alpha*gy)                       (reference op/fused_act.py:27-31)
 *   ga    = d ? tf32_round(g_pre * d[n,c]) : g_pre               (operand of the dgrad / wgrad GEMMs)
 *   g_bias[c]   = sum g_pre            g_noise_w[0] = sum g_pre * noise[n,p]
 *   e[n,c]      = sum_p g_pre * (t - noise_w*noise - bias[c]),  t = pre-activation recovered from y  (NULL: skip)
 * g_bias / g_noise_w / e are zeroed by the call. channels = 4k with k | 256. */
int sr_styled_bwd_prologue_f32(float *ga, float *g_bias, float *g_noise_w, float *e, const float *gy, const float *y,
                               const float *noise, int64_t noise_batch_stride, const float *noise_weight,
                               const float *bias, const float *d, int64_t batch, int64_t pixels, int64_t channels,
                               float alpha, float gain, void *stream);
/* Chained form: the incoming gradient is assembled on the fly from up to three sources,
 *   g_total = gy (or 0) + gxs * s_next[n,c] + sum_k g_rgb[n,p,k] * rgb_weight[n,k,c]
 * (gxs = gradient w.r.t. the NEXT layer's modulated input, g_rgb = gradient of the fused ToRGB output), and the
 * reductions those sources need are produced in the same pass:
 *   ds_next[n,c] = sum_p gxs * y,   d_rgb_weight[n,k,c] = sum_p g_rgb[n,p,k] * y[n,p,c].
 * Everything else as sr_styled_bwd_prologue_f32.  Absent sources are NULL. */
int sr_styled_bwd_prologue2_f32(float *ga, float *g_bias, float *g_noise_w, float *e, float *ds_next, float *d_rgb_weight,
                                const float *gy, const float *gxs, const float *s_next, const float *g_rgb,
                                const float *rgb_weight, const float *y, const float *noise, int64_t noise_batch_stride,
                                const float *noise_weight, const float *bias, const float *d, int64_t batch,
                                int64_t pixels, int64_t channels, float alpha, float gain, void *stream);

/* StyledMapConv variants (reference model.py:33-55: `out = out * stylemap[:, :1] + stylemap[:, 1:2]` between the conv and
 * the noise): stylemap = [batch, 2, h, w] planes with batch stride `stylemap_batch_stride` floats (plane stride h*w), may
 * be NULL.  The FIR tail computes y = lrelu(fir * map0 + map1 + noise_w*noise + bias) * gain.  WITH a stylemap `out`
 * receives fir (the filtered conv output BEFORE the map affine) and only out2 = tf32_round(y * scale2) sees y; the
 * backward prologue then takes that tensor in place of `y`, rebuilds y from it, hands gp * map0 (* d) to the GEMMs and
 * accumulates g_stylemap [batch, 2, pixels] (zeroed by the call):
 *   g_map1 = sum_c gp,  g_map0 = sum_c gp * conv_d      (no division by map0: the map may be exactly 0). */
int sr_blur_nhwc_styled3_f32(float *out, float *out2, const float *scale2, const float *x, const float *taps,
                             int64_t batch, int64_t in_h, int64_t in_w, int64_t channels, int pad0, int pad1,
                             const float *noise, int64_t noise_batch_stride, const float *noise_weight,
                             const float *bias, float alpha, float gain, const float *stylemap,
                             int64_t stylemap_batch_stride, void *stream);
int sr_styled_bwd_prologue3_f32(float *ga, float *g_bias, float *g_noise_w, float *e, float *ds_next, float *d_rgb_weight,
                                const float *gy, const float *gxs, const float *s_next, const float *g_rgb,
                                const float *rgb_weight, const float *y, const float *noise, int64_t noise_batch_stride,
                                const float *noise_weight, const float *bias, const float *d, int64_t batch,
                                int64_t pixels, int64_t channels, float alpha, float gain, const float *stylemap,
                                int64_t stylemap_batch_stride, float *g_stylemap, void *stream);

/* out[n,p,c] = a[n,p,c] * scale[n,c] (optionally rounded to tf32), dot[n,c] = sum_p a[n,p,c] * other[n,p,c];
 * out or dot may be NULL; dot is zeroed by the call. */
int sr_scale_dot_nhwc_f32(float *out, float *dot, const float *a, const float *other, const float *scale,
                          int64_t batch, int64_t pixels, int64_t channels, int round_out_tf32, void *stream);

/* ------------------------------------------------------------------ bf16 operand mode ------------------------
 * BASELINE.json configs[3] (the GAR train step) asks for bf16 convolutions.  In this mode the GEMM OPERANDS -- the
 * modulated activations, the re-laid-out weights and the gradient operand of the dgrad / wgrad GEMMs -- are bfloat16
 * tensors (tcgen05 kind::f16, twice the tensor-core rate and half the operand bytes of the tf32 form); accumulation,
 * the saved activations, all epilogue vectors, reductions and parameter gradients stay fp32.  Each function below is
 * its *_f32 / *_tf32 namesake with the operand pointers retyped (`void *` = bfloat16 data of the same logical shape):
 *   sr_conv_igemm_multi_bf16   in, weight, out2 are bf16 (cin % 64 == 0); out fp32
 *   sr_conv_wgrad_bf16         g, x are bf16; dw fp32
 *   sr_modulate_bf16           xs bf16
 *   sr_conv_weight_prep_dual_bf16   fwd, tr bf16; wsq fp32
 *   sr_blur_nhwc_styled3_bf16  out2 bf16 (out fp32)
 *   sr_blur_nhwc_scaledot_bf16 out bf16
 *   sr_styled_bwd_prologue3_bf16    ga bf16 when d != NULL (fp32 g_pre otherwise: it then feeds the backward FIR) */
int sr_conv_igemm_multi_bf16(const sr_conv_args *args, int count, void *stream);
int sr_conv_wgrad_bf16(const sr_wgrad_args *args, void *stream);
int sr_modulate_bf16(void *xs, const float *x, const float *style, int64_t batch, int64_t pixels, int64_t channels,
                     void *stream);
int sr_conv_weight_prep_dual_bf16(void *fwd, void *tr, float *wsq, const float *w, float scale, int64_t cout, int64_t cin,
                                  int taps, int flip_transposed, void *stream);
int sr_blur_nhwc_styled3_bf16(float *out, void *out2, const float *scale2, const float *x, const float *taps,
                              int64_t batch, int64_t in_h, int64_t in_w, int64_t channels, int pad0, int pad1,
                              const float *noise, int64_t noise_batch_stride, const float *noise_weight,
                              const float *bias, float alpha, float gain, const float *stylemap,
                              int64_t stylemap_batch_stride, void *stream);
int sr_blur_nhwc_scaledot_bf16(void *out, float *dot, const float *x, const float *taps, const float *scale,
                               const float *other, int64_t batch, int64_t in_h, int64_t in_w, int64_t channels, int pad0,
                               int pad1, void *stream);
int sr_styled_bwd_prologue3_bf16(void *ga, float *g_bias, float *g_noise_w, float *e, float *ds_next, float *d_rgb_weight,
                                 const float *gy, const float *gxs, const float *s_next, const float *g_rgb,
                                 const float *rgb_weight, const float *y, const float *noise, int64_t noise_batch_stride,
                                 const float *noise_weight, const float *bias, const float *d, int64_t batch,
                                 int64_t pixels, int64_t channels, float alpha, float gain, const float *stylemap,
                                 int64_t stylemap_batch_stride, float *g_stylemap, void *stream);

/* ------------------------------------------------------------------ style path (all layers per launch) ----
 * Replaces, for every modulated convolution of a network at once, the reference's per-layer
 *   style = self.modulation(style)                              (EqualLinear, reference layers.py:232-239, 295)
 *   demod = rsqrt(weight.pow(2).sum([2,3,4]) + 1e-8)            (reference layers.py:297-299)
 * in the activation-scaling form of this package (the style scales the activations, the demodulation the outputs):
 *   s[b,i] = mod_scale * sum_k latent[b, latent_index, k] * mod_weight[i,k] + lr_mul * mod_bias[i]
 *   d[b,o] = rsqrt( sum_i s[b,i]^2 * wsq[o,i] + eps ),   wsq[o,i] = scale^2 * sum_taps W[o,i,:]^2  (sr_weight_sq_f32)
 * `layers` is a HOST array of n_layers <= 32 descriptors with DEVICE pointers; every layer is one slice of the same
 * launches (2 forward, 4 backward).  latent: [batch, n_latent, style_dim].
 * Backward: given g_s / g_d (either may be NULL) it writes g_mod_weight, g_mod_bias, g_wsq (zeros when the layer has
 * no demodulation gradient) and accumulates g_latent [batch, n_latent, style_dim] (zeroed by the call);
 * gs_total [batch, cin] and du [batch, cout] are caller-provided workspaces. */
typedef struct sr_style_layer {
    const float *mod_weight;   /* [cin, style_dim] */
    const float *mod_bias;     /* [cin] */
    const float *wsq;          /* [cout, cin] or NULL (no demodulation) */
    float *s;                  /* [batch, cin]  (forward: out, backward: in) */
    float *d;                  /* [batch, cout] (forward: out, backward: in), NULL without demodulation */
    int32_t cin, cout, latent_index, reserved;
    const float *g_s;          /* backward inputs */
    const float *g_d;
    float *g_mod_weight;       /* backward outputs */
    float *g_mod_bias;
    float *g_wsq;
    float *gs_total;           /* backward workspaces */
    float *du;
} sr_style_layer;
int sr_style_scales_forward_f32(const sr_style_layer *layers, int n_layers, const float *latent, int64_t batch,
                                int64_t n_latent, int64_t style_dim, float mod_scale, float lr_mul, float eps, void *stream);
int sr_style_scales_backward_f32(const sr_style_layer *layers, int n_layers, const float *latent, float *g_latent,
                                 int64_t batch, int64_t n_latent, int64_t style_dim, float mod_scale, float lr_mul,
                                 void *stream);
/* wsq[o,i] = scale^2 * sum_t w[o,i,t]^2 (w: [cout, cin, taps], the reference weight layout) and its gradient
 * gw[o,i,t] = 2 scale^2 w[o,i,t] g_wsq[o,i]. */
int sr_weight_sq_f32(float *wsq, const float *w, float scale, int64_t cout, int64_t cin, int taps, void *stream);
int sr_weight_sq_backward_f32(float *gw, const float *w, const float *g_wsq, float scale, int64_t cout, int64_t cin,
                              int taps, void *stream);
/* One pass over a [cout, cin, taps] weight: fwd[co][t][ci] and tr[ci][t'][co] (t' = taps-1-t when flip_transposed, the
 * dgrad operand of the plain conv; t' = t for the transposed conv), both tf32(scale*w), and wsq[co][ci] = scale^2 sum_t w^2.
 * Any output may be NULL.  Same results as sr_conv_weight_prep_tf32 (modes 0 / 1 / 2) + sr_weight_sq_f32. */
int sr_conv_weight_prep_dual_tf32(float *fwd, float *tr, float *wsq, const float *w, float scale, int64_t cout,
                                  int64_t cin, int taps, int flip_transposed, void *stream);
/* gw[o,i,t] = scale * dwk[o,t,i]: result of sr_conv_wgrad_tf32 ([cout][taps][cin]) -> reference weight layout. */
int sr_weight_grad_layout_f32(float *gw, const float *dwk, float scale, int64_t cout, int64_t cin, int taps, void *stream);

/* ------------------------------------------------------------------ mesh front-end -------------
 * Area-weighted vertex normals, replaces `mesh_point_normal` (reference utils_3d.py:379-404: three host-built sparse
 * matrix products + Normalize, reference layers.py:13-30):
 *   normals[b,v,:] = normalize( sum over faces f containing v of (v_b - v_a) x (v_c - v_a) ),  |.| clamped at eps.
 * verts [batch,nv,3], tris int64 [nf,3] (shared_f) or [batch,nf,3]; faces with an index outside [0,nv) are skipped.
 * Accumulation uses float atomics (summation order not deterministic). */
int sr_mesh_vertex_normals_f32(float *normals, const float *verts, const int64_t *tris, int64_t batch, int64_t nv,
                               int64_t nf, int shared_f, float eps, void *stream);
/* Rigid pose + scale, replaces the batched matmul of `random_apply_pose3D` (reference utils_3d.py:374-376):
 *   out[b,v,:] = verts[b,v,:3] . R[b] + t[b],   pose[b] = the 3x4 matrix [R | t] row-major ([batch,12]).
 * verts rows are `vert_stride` floats apart (>= 3: extra per-vertex columns are ignored, like v[..., :3]). */
int sr_mesh_pose_apply_f32(float *out, const float *verts, const float *pose, int64_t batch, int64_t nv,
                           int64_t vert_stride, void *stream);
/* The whole mesh front-end of GeneratorWithMap in one call (reference face_model.py:71-72 -> utils_3d.py:360-404 ->
 * model.py:260-270): (optional) pose -> area-weighted vertex normals -> the normal map at every resolution of the
 * generator as [batch, 3, size, size] planes (levels[i].out; .ids / .bary may be NULL and are then not written).
 * verts_out [batch,nv,3] receives the posed vertices (required with a pose), normals [batch,nv,3] the vertex normals;
 * tris int64 [nf,3] shared by the batch; keys = workspace of sr_rasterize_pyramid_workspace_bytes. Forward only. */
int sr_mesh_normal_pyramid_f32(int64_t batch, int64_t nv, int64_t nf, const float *verts_in, int64_t vert_stride,
                               const float *pose, const int64_t *tris, float *verts_out, float *normals, int n_levels,
                               const sr_raster_level *levels, uint64_t *keys, float eps, void *stream);

/* sr_conv_weight_prep_dual_* / sr_weight_sq_backward_f32 for up to SR_WEIGHT_PREP_MAX conv weights in ONE launch each (every
 * ModulatedConv2d of a generator, reference layers.py:296-299): fwd / tr are tf32-rounded fp32 (bfloat16 for _bf16) GEMM
 * operands or NULL, wsq fp32 or NULL; the backward reads w and g_wsq [cout, cin] and writes gw [cout, cin, taps]. */
#define SR_WEIGHT_PREP_MAX 16
typedef struct sr_weight_prep_item {
    void *fwd, *tr;
    float *wsq;
    const float *w;
    float *gw;
    const float *g_wsq;
    float scale;
    int32_t cout, cin, taps, flip_transposed, reserved;
} sr_weight_prep_item;
int sr_conv_weight_prep_multi_tf32(const sr_weight_prep_item *items, int n, void *stream);
int sr_conv_weight_prep_multi_bf16(const sr_weight_prep_item *items, int n, void *stream);
int sr_weight_sq_backward_multi_f32(const sr_weight_prep_item *items, int n, void *stream);
/* out = (a + b) * scale over n fp32 elements (n % 4 == 0) and, when `operand` is not NULL, the GEMM-operand copy of out
 * (tf32-rounded fp32 / bfloat16) in the same pass: the residual sum `(out + skip) / sqrt(2)` of a Discriminator ResBlock
 * (reference layers.py:386-391) together with the operand conversion of the next block's first convolution. */
int sr_residual_combine_tf32(float *out, float *operand, const float *a, const float *b, float scale, int64_t n, void *stream);
int sr_residual_combine_bf16(float *out, void *operand, const float *a, const float *b, float scale, int64_t n, void *stream);
/* The style-map network of GeneratorWithMap -- ResBlock(3 -> cout, downsample = False), cout = 2 or 4 (reference
 * model.py:194-216 builds them, model.py:262,271-275 runs them on the rasterised normal map of every resolution; the block
 * is reference layers.py:379-391 over the ConvLayers of layers.py:341-378) -- as ONE pass over [batch, 3, h, w] planes:
 *   y1  = lrelu(conv3x3(x,  w1 / sqrt(27)) + b1_conv + b1_act) * gain
 *   y2  = lrelu(conv3x3(y1, w2 / sqrt(27)) + b2_conv + b2_act) * gain
 *   out = (y2 + conv1x1(x, w_skip / sqrt(3))) / sqrt(2)                       out: [batch, cout, h, w] planes
 * Weights are the raw parameters (w1 [3,3,3,3], w2 [cout,3,3,3], w_skip [cout,3,1,1]); bias pointers may be NULL.
 * Replaces three cuDNN convolutions and five elementwise passes per block (and their autograd graph). */
int sr_stylemap_resblock_forward_f32(float *out, const float *x, const float *w1, const float *b1_conv, const float *b1_act,
                                     const float *w2, const float *b2_conv, const float *b2_act, const float *w_skip,
                                     int64_t batch, int cin, int cout, int64_t h, int64_t w, float alpha, float gain,
                                     void *stream);
/* Parameter gradients of the block above from grad_out [batch, cout, h, w] and x (y1 and both activation masks are
 * recomputed): grads = [d w1 (81) | d b1 (3) | d w2 (27 cout) | d b2 (cout) | d w_skip (3 cout)] floats, overwritten.
 * d b1 / d b2 are the gradient of BOTH biases of the layer (conv bias and activation bias enter as a sum).
 * The gradient with respect to x is not produced. */
int sr_stylemap_resblock_backward_f32(float *grads, const float *grad_out, const float *x, const float *w1,
                                      const float *b1_conv, const float *b1_act, const float *w2, const float *b2_conv,
                                      const float *b2_act, const float *w_skip, int64_t batch, int cin, int cout,
                                      int64_t h, int64_t w, float alpha, float gain, void *stream);

/* Small-channel convolution pair (1..8 channels, 1x1 or 3x3 with zero padding k/2, stride 1, [batch, c, h, w] planes):
 * y = conv(x, w), w [cout, cin, k, k] (cross-correlation like F.conv2d, no bias), and its weight gradient
 * dw[o,i,ky,kx] = sum gy[n,o,p] x[n,i,p + (ky,kx) - k/2] (cin * cout * k * k <= 256).  The two maps are each other's
 * derivatives, so they carry the style-map nets through the regulariser iterations that differentiate twice (reference
 * train.py:335-354 with the normal maps among the inputs), where torch's double backward of the cuDNN convolutions computes
 * weight gradients as convolutions with image-sized kernels. */
int sr_small_conv_f32(float *y, const float *x, const float *w, int64_t batch, int cin, int cout, int ksize, int64_t h,
                      int64_t wd, void *stream);
int sr_small_conv_wgrad_f32(float *dw, const float *gy, const float *x, int64_t batch, int cin, int cout, int ksize,
                            int64_t h, int64_t wd, void *stream);
/* The Discriminator's stem, ConvLayer(3, cout, 1) = EqualConv2d 1x1 + bias + FusedLeakyReLU (reference model.py:303,
 * layers.py:341-378), as one bandwidth pass per direction.  x: [batch, 3, h, w] planes or (x_channels_last) [batch, h, w, 3];
 * y / gy: channels-last [batch, h, w, cout]; w [cout, 3] raw (scaled by 1/sqrt(3) inside); biases may be NULL.
 *   forward : y = lrelu(conv1x1(x, w / sqrt(3)) + b_conv + b_act) * gain
 *   backward: grads = [d w (3 cout) | d b (cout)] (d b is the gradient of both biases), dx (layout of x) optional;
 *             the activation mask is recomputed from x, so nothing but gy is read at full size. */
int sr_stem_conv_forward_f32(float *y, const float *x, const float *w, const float *b_conv, const float *b_act,
                             int64_t batch, int cin, int64_t cout, int64_t h, int64_t wd, int x_channels_last, float alpha,
                             float gain, void *stream);
int sr_stem_conv_backward_f32(float *grads, float *dx, const float *gy, const float *x, const float *w, const float *b_conv,
                              const float *b_act, int64_t batch, int cin, int64_t cout, int64_t h, int64_t wd,
                              int x_channels_last, float alpha, float gain, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* STYLERENDERER_B200_H_ */
