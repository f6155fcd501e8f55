/*
 * sg2b200.h -- C ABI of libsg2b200.so: the B200 (sm_100a) kernels behind the
 * StyleGAN2 G+D training step of STomoya/animeface.
 *
 * Every entry point is what the reference's pybind11 plugin layer (or the ATen
 * call it makes) for this path would bind.  The reference interface each entry
 * replaces is cited as  <file>:<line>  relative to the reference checkout.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless
 *     stated otherwise; no torch types.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on
 *     that stream and never synchronises.
 *   - return value 0 = success, negative = error (SG2_E*); the message for the
 *     calling thread is available through sg2_last_error().
 *   - strides are in ELEMENTS, order (n, c, h, w), so NCHW-contiguous and
 *     channels_last (NHWC) tensors are both described without copies -- the
 *     same contract as upfirdn2d.cpp:32 / bias_act.cpp:41 (output keeps the
 *     input's memory format).
 *   - dtype codes: SG2_F32 = 0, SG2_F16 = 1, SG2_F64 = 2 (internal math is fp32,
 *     fp64 for SG2_F64 -- thirdparty/stylegan3_ops/ops/upfirdn2d.cu:9-12).
 */
#ifndef SG2B200_H
#define SG2B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG2_OK        0
#define SG2_EINVAL   -1   /* bad argument (what TORCH_CHECK raises in the reference) */
#define SG2_ELAUNCH  -2   /* CUDA launch / runtime error */
#define SG2_ENOTSUP  -3   /* valid request, no kernel for it */

#define SG2_F32 0
#define SG2_F16 1
#define SG2_F64 2

typedef void* sg2_stream_t;

/* library ------------------------------------------------------------------ */
int         sg2_version(void);
const char* sg2_last_error(void);
/* number of kernels this library has launched in the calling process. */
int64_t     sg2_launch_count(void);
/* profiling aid (no reference counterpart): when `device_buf` is non-null the persistent tcgen05 convolution kernels
 * add, per CTA, the cycles each warp role spent waiting on each pipeline barrier into device_buf[cta * 16 + slot]
 * (int64, caller-zeroed, >= 16 * grid entries).  NULL switches it off (default). */
int         sg2_debug_trace(void* device_buf);

/* upfirdn2d ---------------------------------------------------------------- *
 * replaces: thirdparty/stylegan3_ops/ops/upfirdn2d.cpp:10  (pybind `upfirdn2d`)
 *           + kernels thirdparty/stylegan3_ops/ops/upfirdn2d.cu:23,92
 * Pad -> zero-insert upsample -> FIR -> decimate, one launch.
 * f: [fh, fw] float32, contiguous.  out_h/out_w must equal
 *   (in*up + pad0 + pad1 - f + down) / down      (upfirdn2d.cpp:29-30).      */
int sg2_upfirdn2d(const void* x, const float* f, void* y, int dtype,
                  int n, int c, int in_h, int in_w, const int64_t x_strides[4],
                  int out_h, int out_w, const int64_t y_strides[4],
                  int fh, int fw, int upx, int upy, int downx, int downy,
                  int padx0, int padx1, int pady0, int pady1,
                  int flip, float gain, sg2_stream_t stream);

/* StyleGAN2 resampling (channels_last fp32) ---------------------------------- *
 * replaces: implementations/StyleGAN2/model.py:56-58 (Upsample2x 'bilinear',
 *           align_corners=False) followed by model.py:138-149 (Blur2d,
 *           [1,2,1]x[1,2,1]/16, zero pad 1) -- ONE pass, replicate-then-zero
 *           border.  blur=0 gives the bare bilinear x2 of ToImage
 *           (model.py:243,248-249).
 * layout: x [n,h,w,c] / y [n,2h,2w,c] dense NHWC when nhwc=1, else dense NCHW.
 * scale: optional [n,c] per-(sample,channel) factor applied to the output.
 * The *_adj entry is the exact adjoint (backward of fwd; and fwd is the
 * backward of adj, so the pair is closed under differentiation).             */
int sg2_up2x_fwd(const float* x, float* y, const float* scale,
                 int n, int c, int h, int w, int nhwc, int blur, sg2_stream_t stream);
int sg2_up2x_adj(const float* gy, float* gx, const float* scale,
                 int n, int c, int h, int w, int nhwc, int blur, sg2_stream_t stream);

/* replaces: implementations/StyleGAN2/model.py:61-63 (AvgPool2d(2)) and the
 * residual merge (x + t)/sqrt(2) of DBlock.forward model.py:209-212.
 * y = alpha * (avg2x2(x) + (t ? avg2x2(t) : 0)).   x,t [n,h,w,c] NHWC dense, h,w even.
 * t_pooled = 1: t is [n,h/2,w/2,c] and is added as it is, y = alpha * (avg2x2(x) + t) -- the skip branch of the block
 *   computed at the pooled resolution (a 1x1 convolution commutes with the average pooling that follows it).
 * signs (may be NULL): [n,h/2,w/2,c/4] uint16, the signs (x > 0) of each 2x2 window x 4 channels (bit 4*pixel + channel) --
 *   sg2_bwd_prep_planes reads them instead of x when it applies the leaky-ReLU gradient of the layer that produced x.
 * adj: gx = alpha * 0.25 * gy broadcast over each 2x2 window.                 */
int sg2_avgpool2_fwd(const float* x, const float* t, float* y, void* signs, float alpha,
                     int n, int c, int h, int w, int t_pooled, sg2_stream_t stream);
int sg2_avgpool2_adj(const float* gy, float* gx, float alpha,
                     int n, int c, int h, int w, sg2_stream_t stream);

/* bias_act ----------------------------------------------------------------- *
 * replaces: thirdparty/stylegan3_ops/ops/bias_act.cpp:26 (pybind `bias_act`)
 *           + kernel thirdparty/stylegan3_ops/ops/bias_act.cu:17-141
 * Works on the flat dense buffer: b index = (i / step_b) % size_b
 * (bias_act.cpp:47-48), so contiguous and channels_last need no copy.
 * Null pointer = "absent" (the reference's empty tensor, bias_act.py:31).
 * grad: 0 forward, 1 first-order, 2 second-order.  act: 1 linear 2 relu
 * 3 lrelu 4 tanh 5 sigmoid 6 elu 7 selu 8 softplus 9 swish (bias_act.py:16-26).
 * clamp < 0 disables clamping.                                               */
int sg2_bias_act(const void* x, const void* b, const void* xref, const void* yref,
                 const void* dy, void* y, int dtype, int64_t numel,
                 int size_b, int64_t step_b, int grad, int act,
                 float alpha, float gain, float clamp, sg2_stream_t stream);

/* minibatch stddev --------------------------------------------------------- *
 * replaces: implementations/StyleGAN2/model.py:215-236 (MiniBatchStdDev.forward;
 *           = nnutils/module/layers.py:40-52).
 * x [n,c,h,w] (strides given) -> y [n,c+1,h,w] (strides given): y[:, :c] = x,
 * y[:, c] = mean_{c,h,w} sqrt(var_g(x) + eps) of the sample's group column.
 * `groups` = G as resolved by the caller (group_size if n % group_size == 0
 * else n, model.py:234-236); sample i belongs to column m = i % (n/G).
 * stat: [n/G] float32 workspace/output (the per-column statistic).
 * bwd: gx = gy[:, :c] + d(stat)/dx * sum_{g,h,w} gy[g*M+m, c, h, w].          */
int sg2_mbstd_fwd(const float* x, const int64_t x_strides[4], float* y,
                  const int64_t y_strides[4], float* stat,
                  int n, int c, int h, int w, int groups, float eps, sg2_stream_t stream);
int sg2_mbstd_bwd(const float* x, const int64_t x_strides[4], const float* gy,
                  const int64_t gy_strides[4], float* gx, const int64_t gx_strides[4],
                  int n, int c, int h, int w, int groups, float eps, sg2_stream_t stream);

/* dense / modulated convolution ------------------------------------------- *
 * replaces: the ATen/cuDNN calls of the path --
 *   F.conv2d(groups=B) of ModulatedConv2d.forward implementations/StyleGAN2/model.py:106-132,
 *   nn.Conv2d inside ELR model.py:29-37,50-53 (DBlock model.py:186-212),
 *   and their autograd (convolution_backward: dgrad + wgrad), composed exactly as
 *   thirdparty/stylegan3_ops/ops/conv2d_gradfix.py:99-187 composes them.
 * Stride-1 "same" cross-correlation, odd square kernel k (1 or 3), zero padding.
 * x  [n,h,w,ci] NHWC dense fp32;  y [n,h,w,co] with explicit strides (n,c,h,w).
 *
 * Weights are passed PACKED: sg2_conv2d_pack_weight turns the reference layout
 * w[co][ci][k][k] into the operand layout of the selected kernel, folding the
 * ELR coefficient (`coef`, model.py:32,105) in.  transpose=1 packs the weight of
 * the data-gradient conv (ci<->co swapped, taps flipped): dgrad is then the SAME
 * forward kernel applied to gy.
 *
 * Fused prologue/epilogue of sg2_conv2d_fwd (any pointer may be NULL):
 *   in_scale  [n,ci]   style modulation s[b,i] applied to x      (model.py:115)
 *   out_scale [n,co]   demodulation d[b,o]                         (model.py:118-120)
 *   bias      [co]                                                  (model.py:132)
 *   noise     [n,h,w]  InjectNoise, added unscaled                  (model.py:85-88)
 *   act: 1 linear, 3 lrelu(alpha); then * gain                      (model.py:164)
 *   y = gain * act( out_scale * conv(x * in_scale, w) + bias + noise )
 * impl: 0 = auto, 1 = fp32 kernels (SIMT implicit GEMM; one-pass "thin" kernels when one side has <= 4 channels),
 *       4, 5 = tcgen05 "halo" kernels: the (8+2)x(16+2) patch of a tile is loaded once and the k*k taps are shifted descriptor
 *       windows of it (input channels % 32 == 0, output channels % 32 == 0; images that tile by 8x16, ragged images >= 32x32
 *       with masked edge tiles, and small images whose tiles fit the machine in one wave):
 *   4 = "bf16x3": fp32 operands split into bf16 hi/lo pairs, the three products hi*hi + lo*hi + hi*lo accumulated in fp32
 *       in TMEM (~5e-6 relative) -- data and weight gradients, which are linear in their operands;
 *   5 = "fp32-class": operands split into fp16 big/small pairs carrying 22 mantissa bits (the residual scaled by 2^11), at
 *       the full 16-bit MMA rate; the big*big accumulator is promoted into fp32 registers every few taps (the tensor core
 *       truncates its accumulator per MMA): fp32-class results (~5e-7).  FORWARD convs use it: a relative input error eps
 *       flips ~0.8*eps of the leaky-ReLU signs against the fp32 reference.
 *   (2 and 3 were round 1's per-tap kernels; they no longer exist and are refused.)
 *   sg2_conv2d_select_impl resolves `impl` for a shape (0 = auto: halo, else fp32; `precise` picks 5 over 4) and returns
 *   1, 4 or 5 (or SG2_ENOTSUP); pass the returned code to BOTH pack_weight and fwd.   */
int sg2_conv2d_select_impl(int n, int h, int w, int ci, int co, int k, int impl, int precise);
int64_t sg2_conv2d_packed_size(int co, int ci, int k, int impl);   /* bytes */
int sg2_conv2d_pack_weight(const float* w, void* packed, int co, int ci, int k,
                           float coef, int transpose, int impl, sg2_stream_t stream);
int sg2_conv2d_fwd(const float* x, const void* packed_w, float* y, const int64_t y_strides[4],
                   int n, int h, int w, int ci, int co, int k,
                   const float* in_scale, const float* out_scale, const float* bias,
                   const float* noise, int act, float alpha, float gain,
                   int impl, sg2_stream_t stream);
/* weight gradient: dw[co][ci][k][k] (reference layout) = coef * sum_{n,h,w} gy (x) x.
 * x [n,h,w,ci], gy [n,h,w,co] NHWC dense.  in_scale [n,ci] / out_scale [n,co]
 * optional: x is taken as x*in_scale and gy as gy*out_scale (modulated layers).
 * accumulate=0 overwrites dw, 1 adds into it.
 * workspace: sg2_conv2d_wgrad_workspace(...) bytes, or NULL.  With a workspace every pixel split stores its partial dw and
 * a second kernel adds the splits in a fixed order: run-to-run identical results (what CUDA-graph replay == eager needs).
 * NULL keeps the one-pass fp32-atomic reduction (order, and so the last bits, vary from run to run).                   */
int64_t sg2_conv2d_wgrad_workspace(int n, int h, int w, int ci, int co, int k, int impl);
int sg2_conv2d_wgrad(const float* x, const float* gy, float* dw,
                     int n, int h, int w, int ci, int co, int k, float coef,
                     const float* in_scale, const float* out_scale,
                     int accumulate, int impl, void* workspace, sg2_stream_t stream);

/* per-(sample,channel) reductions used by the modulated-conv backward ------- *
 * replaces: the autograd graph of ModulatedConv2d.forward
 *           (implementations/StyleGAN2/model.py:106-132).
 * a, bm, a_out: [n,hw,c] NHWC dense, c % 4 == 0.
 *   out[b,c]   = sum_hw a * (bm ? bm : 1)         (d s[b,i] = sum_hw x * g_xs)
 *   a_out      = a * scale[b,c]   when a_out != NULL  (g_x = g_xs * s, same pass)  */
/* workspace (all three calls below): sg2_reduce_hw_workspace(n, hw, c) bytes for a deterministic two-pass reduction over the
 * hw slices, or NULL for one pass with fp32 atomics.                                                                      */
int64_t sg2_reduce_hw_workspace(int n, int hw, int c);
int sg2_reduce_hw(const float* a, const float* bm, float* out,
                  int n, int hw, int c, void* workspace, sg2_stream_t stream);
int sg2_scale_reduce_hw(const float* a, const float* bm, const float* scale,
                        float* a_out, float* out, int n, int hw, int c, void* workspace, sg2_stream_t stream);
/* backward prologue of y = lrelu_alpha(d * acc + bias + noise) in ONE pass over (gy, y):
 *   gu = gy * (y > 0 ? 1 : alpha);  g_acc = gu * d;  gb_part[b,o] = sum_hw gu;
 *   gd[b,o] = sum_hw gu * acc,  acc recovered from y as (u - bias - noise)/d.
 * alpha = 1 means "no activation".  d, gd, bias, noise may be NULL.           */
int sg2_modconv_bwd_prep(const float* gy, const float* y, const float* noise, const float* bias,
                         const float* d, float* g_acc, float* gb_part, float* gd,
                         int n, int hw, int c, float alpha, void* workspace, sg2_stream_t stream);

/* bf16 pair planes: the first-order backward fast path ----------------------- *
 * replaces: the same ATen convolution_backward calls as sg2_conv2d_fwd(transposed pack) / sg2_conv2d_wgrad above
 *           (implementations/StyleGAN2/model.py:129, 44-53), for operands that an elementwise pass has already
 *           written as "pair planes": hi = bf16(v), lo = bf16(v - hi), stored back to back as [2][n][hw][c] bf16
 *           (4 bytes per element, like the fp32 tensor they stand for).  The tcgen05 kernels then take their
 *           operands by TMA straight into swizzled tiles -- no in-kernel fp32 -> bf16 conversion -- and the k*k taps
 *           are descriptor offsets into ONE tile.  bf16x3 arithmetic (~5e-6 relative), as impl 4.
 * sg2_split_planes:    planes = split(x * scale[b,c])   (scale may be NULL)                        c % 4 == 0
 * sg2_bwd_prep_planes: sg2_modconv_bwd_prep with g_acc written as planes and DETERMINISTIC per-(sample, channel) sums:
 *                      gb[c] = sum_{n,hw} gu, gd[n,c] = sum_hw gu * acc (NULL = skip); y NULL = no activation (gu = gy);
 *                      workspace: sg2_bwd_prep_planes_workspace(n, hw, c) bytes.  pool_w > 0: gy is the gradient of the 2x2
 *                      average pooling that follows the layer (implementations/StyleGAN2/model.py:209-212), [n, hw/4, c] with
 *                      full-resolution width pool_w; its adjoint (broadcast * gscale) is applied on the fly.  gscale also
 *                      scales an ordinary gy.  signs (pooled form, y = NULL): the window signs sg2_avgpool2_fwd wrote, read
 *                      instead of y (2 bytes instead of 64 per window).  sg2_conv2d_fwd_planes(accumulate = 1) adds into y instead of overwriting it.
 * sg2_conv2d_fwd_planes: y = gain * act(out_scale * conv(x, w) + bias) with x given as planes [2][n,h,w,ci]
 *                      (ci % 64 == 0); packed_w from sg2_conv2d_pack_weight(impl = 4) -- with transpose = 1 this is the
 *                      data gradient.  sg2_conv2d_planes_supported(.., wgrad = 0) tells whether the shape is taken.
 * sg2_conv2d_wgrad_planes: dw[co][ci][k][k] (+)= coef * sum gy (x) x with both operands as planes (ci, co % 64 == 0,
 *                      image width >= 8); run-to-run deterministic (two-pass split-K through `workspace`,
 *                      sg2_conv2d_wgrad_planes_workspace bytes).                                                    */
int sg2_split_planes(const float* x, const float* scale, void* planes, int n, int hw, int c, sg2_stream_t stream);
int64_t sg2_bwd_prep_planes_workspace(int n, int hw, int c);
int sg2_bwd_prep_planes(const float* gy, const float* y, const float* noise, const float* bias, const float* d,
                        void* planes, float* gb, float* gd, void* workspace, const void* signs,
                        int n, int hw, int c, float alpha, int pool_w, float gscale, sg2_stream_t stream);
int sg2_conv2d_planes_supported(int n, int h, int w, int ci, int co, int k, int wgrad);
int sg2_conv2d_fwd_planes(const void* x_planes, const void* packed_w, float* y, const int64_t y_strides[4],
                          int n, int h, int w, int ci, int co, int k,
                          const float* out_scale, const float* bias, int act, float alpha, float gain,
                          int accumulate, sg2_stream_t stream);
int64_t sg2_conv2d_wgrad_planes_workspace(int n, int h, int w, int ci, int co, int k);
int sg2_conv2d_wgrad_planes(const void* x_planes, const void* gy_planes, float* dw, void* workspace,
                            int n, int h, int w, int ci, int co, int k, float coef, int accumulate,
                            sg2_stream_t stream);

/* backward of a thin-input 1x1 convolution + bias + leaky ReLU ---------------- *
 * replaces: convolution_backward + LeakyReLU backward + the bias reduction behind Discriminator.from_rgb
 *           (implementations/StyleGAN2/model.py:383-384: Conv2d('elr', 3, 32, 1) + LeakyReLU) in ONE pass over (gy, y):
 *             gu = gy * gain * (y > 0 ? 1 : slope);  gw[o][c] = coef * sum gu x;  gb[o] = sum gu;  gx[pix][c] = coef * sum_o gu w[o][c]
 * gy, y [n,hw,co] NHWC dense (co in {4,8,16,32,64}); x [n,hw,cin] (cin <= 4); w [co][cin] (the reference layout, k = 1).
 * gx, gw, gb may be NULL.  workspace: sg2_thin_in_bwd_workspace bytes.  Deterministic (block partials added in order).         */
int64_t sg2_thin_in_bwd_workspace(int n, int hw, int cin, int co);
int sg2_thin_in_bwd(const float* gy, const float* y, const float* x, const float* w, float* gx, float* gw, float* gb,
                    void* workspace, int n, int hw, int cin, int co, float slope, float gain, float coef, sg2_stream_t stream);

/* demodulation coefficient ----------------------------------------------------- *
 * replaces: the tensor expression of implementations/StyleGAN2/model.py:115-120 reduced to the [B,Co] coefficient
 *           d[b,o] = rsqrt(coef^2 * sum_i s[b,i]^2 * sum_k w[o,i,k]^2 + eps)   (pow, reduce, sgemm, mul, add, rsqrt)
 *           and its autograd backward (two sgemms + elementwise): one launch forward, two backward.
 * w [co][ci][kk], s [B][ci], wsq [co][ci] (written by fwd, read by bwd), d / gd [B][co]; gw [co][ci][kk] and gs [B][ci] may be
 * NULL (skipped).  ci, co <= 2048, B <= 256.  No atomics.                                                                   */
int sg2_demod_fwd(const float* w, const float* s, float* wsq, float* d, int B, int co, int ci, int kk, float coef, float eps,
                  sg2_stream_t stream);
int sg2_demod_bwd(const float* w, const float* s, const float* wsq, const float* d, const float* gd, float* gw, float* gs,
                  int B, int co, int ci, int kk, float coef, sg2_stream_t stream);

/* filtered_lrelu ------------------------------------------------------------- *
 * replaces: thirdparty/stylegan3_ops/ops/filtered_lrelu.py:50-268 (plugin entry filtered_lrelu.cpp:17, kernels
 *           filtered_lrelu.cu:133-1093), called from implementations/StyleGAN3/model.py:186-190:
 *             z = up^2 * FIR_fu(zero-insert(x + b[c], up), padded);  a = clamp(lrelu_slope(z) * gain);  y = decimate_down(FIR_fd(a))
 *           as ONE kernel: a CTA keeps the input tile, the up-sampled intermediates and the activation in shared memory.
 * x [planes = N*C][in_h][in_w] -> y [planes][out_h][out_w], dense NCHW fp32.  The "z grid" (zh x zw) is the up-sampled,
 * fu-filtered signal: zw = in_w*up + padx0 + padx1 - (fu_n - 1).  Filters are given ORIENTED FOR CORRELATION (the caller
 * flips them as upfirdn2d's flip_filter says): z[u] = up_gain * sum_t fu[t] xup[u + t] with xup[q] = x[(q - pad0)/up] where
 * divisible, y[o] = sum_s fd[s] a[o*down + s + doff].  fu: fu_n taps applied along both axes, or fu_n x fu_n (fu_2d = 1); fd:
 * fd_n taps (separable) or fd_n x fd_n (fd_2d = 1).
 * mode 0: activation.  mode 1: activation, and mask[planes][zh][zw] (bytes) receives 0 (z <= 0), 1 (z > 0) or 2 (clamped).
 * mode 2: a = z * gain * {slope, 1, 0}[mask] instead of the activation -- with x = dy, the flipped filters in swapped roles,
 *   up <-> down, pad0 = fd_n - 1, doff = pad0_fwd - (fu_n - 1), up_gain = 1 and gain = gain_fwd * up_fwd^2 this is the
 *   gradient w.r.t. x on the SAME z grid (the reference re-pads and offsets its sign tensor instead, filtered_lrelu.py:236-247). */
int sg2_filtered_lrelu(const float* x, const float* b, float* y, void* mask, const float* fu, const float* fd, int fu_2d, int fd_2d,
                       int planes, int channels, int in_h, int in_w, int up, int pad0x, int pad0y, int zh, int zw,
                       int fu_n, int fd_n, int down, int doffx, int doffy, int out_h, int out_w,
                       float up_gain, float gain, float slope, float clamp, int mode, sg2_stream_t stream);

/* fully connected layers ------------------------------------------------------ *
 * replaces: nn.Linear inside ELR (implementations/StyleGAN2/model.py:29-37, 44-47) = ATen addmm (cuBLAS) -- the 8
 *           MapLinear + LeakyReLU of Mapping (:71-78, 263-282), ModulatedConv2d.affine (:102, 110), the discriminator
 *           epilogue Linear(8192,512) -> LeakyReLU -> Linear(512,1) (:392-396) -- and PixelNorm (:253-256).
 * x [B,K], w [N,K] (nn.Linear layout), bias [N] or NULL, y [B,N]; all dense fp32.  slope = 1 means "no activation".
 *   fwd:        y  = lrelu_slope( gain * (coef * x W^T + bias) )
 *   bwd_data:   gx = coef * gu W,      gu = gy * gain * (y > 0 ? 1 : slope); y NULL -> gu = gy * gain
 *   bwd_weight: gw = coef * gu^T x,    gb = sum_b gu (gb may be NULL)
 * With bias = NULL, gain = 1, slope = 1, y = NULL the three calls are F(x,W) = x W^T, Dx(g,W) = g W, Dw(g,x) = g^T x:
 * a family closed under differentiation (what R1's double backward through the epilogue uses).  No atomics.
 *   pixelnorm:  y = x / (sqrt(mean_k x^2) + eps)
 * fwd splits a large K (the 8192-wide discriminator layer) over CTAs and adds the partial sums in a fixed order in a second
 * pass: `workspace` = sg2_linear_fwd_workspace(B, K, N) bytes of device memory (0 bytes / NULL when K fits one slice).           */
long long sg2_linear_fwd_workspace(int B, int K, int N);
int sg2_linear_fwd(const float* x, const float* w, const float* bias, float* y, int B, int K, int N,
                   float coef, float gain, float slope, void* workspace, sg2_stream_t stream);
int sg2_linear_bwd_data(const float* gy, const float* y, const float* w, float* gx, int B, int K, int N,
                        float coef, float gain, float slope, sg2_stream_t stream);
int sg2_linear_bwd_weight(const float* gy, const float* y, const float* x, float* gw, float* gb, int B, int K, int N,
                          float coef, float gain, float slope, sg2_stream_t stream);
int sg2_pixelnorm(const float* x, float* y, int B, int K, float eps, sg2_stream_t stream);

/* DiffAugment ------------------------------------------------------------------ *
 * replaces: thirdparty/diffaugment/DiffAugment.py:10-77 with policy a subset of 'color,translation,cutout' in that order
 *           (the training step uses 'color,translation': implementations/StyleGAN2/utils.py:63-68,92) -- one pass instead of
 *           ~10.  x, y [B,C,H,W] dense NCHW fp32, C <= 8.  rb/rs/rc: [B] raw U[0,1) draws (DiffAugment.py:24,30,36) or NULL;
 *           ty/tx: [B] int64 shifts (:42-43) or NULL; cy/cx: [B] int64 cut-out offsets (:58-59) or NULL with the window
 *           size cut_h x cut_w.  backward = 1 applies the transpose (x = gy, y = gx); linear_only = 1 drops the brightness
 *           constant (the derivative of the map).  workspace: sg2_diffaugment_workspace bytes.  Deterministic.        */
int64_t sg2_diffaugment_workspace(int B, int H, int W);
int sg2_diffaugment(const float* x, float* y, const float* rb, const float* rs, const float* rc,
                    const int64_t* ty, const int64_t* tx, const int64_t* cy, const int64_t* cx, int cut_h, int cut_w,
                    int B, int C, int H, int W, int backward, int linear_only, void* workspace, sg2_stream_t stream);

/* ADA augmentation pipeline -------------------------------------- *
 * replaces: the ATen calls of thirdparty/ada/augment.py:115-427 that the StyleGAN2 step does not have --
 *   torch.nn.functional.pad(mode='reflect') (:284), affine_grid + grid_sample_gradfix.grid_sample (:293-295; bilinear,
 *   zeros padding, align_corners=False; thirdparty/stylegan3_ops/ops/grid_sample_gradfix.py:1-77), and the per-sample
 *   colour matrix product (:352-361).  Dense NCHW fp32.  adjoint / transpose = 1 applies the exact adjoint (x = gy, y = gx);
 *   every op is linear in the image, so forward and adjoint are each other's derivative.
 * reflect_pad:   x [planes, h, w] -> y [planes, h + py0 + py1, w + px0 + px1]
 * affine_sample: theta [n, 2, 3] maps normalised OUTPUT coordinates to normalised INPUT coordinates (F.affine_grid);
 *                x [n,c,ih,iw] -> y [n,c,oh,ow]; the sampling grid is never materialised.  The adjoint scatters with fp32
 *                atomics (as ATen's grid_sampler_2d_backward does).
 * color_affine:  cmat [n, 4, 4] homogeneous colour transforms; x, y [n, 3, hw]: y = C[:3,:3] x + C[:3,3].            */
int sg2_reflect_pad(const float* x, float* y, int64_t planes, int h, int w, int px0, int px1, int py0, int py1,
                    int adjoint, sg2_stream_t stream);
int sg2_affine_sample(const float* x, float* y, const float* theta, int n, int c, int ih, int iw, int oh, int ow,
                      int adjoint, sg2_stream_t stream);
int sg2_color_affine(const float* x, float* y, const float* cmat, int n, int64_t hw, int transpose, sg2_stream_t stream);

/* optimizer ---------------------------------------------------------------- *
 * replaces: torch.optim.Adam.step (implementations/StyleGAN2/utils.py:220-221,
 *           85-86,112-113; ~125 per-tensor launches) over ONE flat fp32 buffer,
 *           and update_ema (nnutils/training.py:23-40; 81 lerps) likewise.
 * p,g,m,v [numel]; ema may be NULL (else ema = decay*ema + (1-decay)*p_new).
 * grad_scale multiplies g first (1/world after the all-reduce).
 * step: 1-based step count of this segment, held on the DEVICE (int64) so the
 * call replays inside a CUDA graph; the caller increments it.                 */
int sg2_adam_ema(float* p, const float* g, float* m, float* v, float* ema,
                 int64_t numel, const int64_t* step, float lr, float beta1, float beta2,
                 float eps, float grad_scale, float ema_decay, sg2_stream_t stream);
int sg2_ema_update(float* ema, const float* p, int64_t numel, float decay, sg2_stream_t stream);
/* Multi-tensor form: ONE launch over the flat buffer with exact per-tensor step counts.
 * seg_off: int64[nseg+1] element offsets of the tensors inside the flat buffer (16-byte aligned segments);
 * steps: int64[nseg] per-tensor step counts (device; incremented here for present tensors);
 * present: int32[nseg], 0 = the tensor had no gradient this step -> skipped exactly like torch.optim.Adam skips
 * grad=None (its moments and step count do not move); coef_ws: float2[nseg] workspace.                        */
int sg2_adam_multi(float* p, const float* g, float* m, float* v, int64_t numel, const int64_t* seg_off,
                   int64_t* steps, const int* present, void* coef_ws, int nseg,
                   float lr, float beta1, float beta2, float eps, float grad_scale, sg2_stream_t stream);
/* counters[first .. first+count) += delta  (the per-tensor Adam step counts; a kernel so that it is captured
 * in CUDA graphs without any host->device copy). */
int sg2_counter_add(void* counters, int first, int count, int delta, int is64, sg2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SG2B200_H */
