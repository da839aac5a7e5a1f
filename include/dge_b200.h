/*
 * dge_b200.h -- C ABI of the B200-native GAN-inversion hot path (E forward + frozen G forward).
 *
 * The reference (disanda/Deep-GAN-Encoders) has NO FFI / plugin layer: its hot path is a chain of
 * ATen calls issued from Python nn.Modules.  Each entry point below replaces one such chain; the
 * reference call site it stands in for is cited as `file:line` (paths relative to the reference
 * repository root).  Host code (the mirrored nn.Modules under deep-gan-encoders_b200/model) binds
 * these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator in the
 *     shipped host code); the library never allocates, frees or retains device memory;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = OK, negative = error; dge_last_error() returns a thread-local message;
 *   - nothing throws across the boundary; sm_100a only (dge_device_ok() reports it).
 *
 * Device tensor layouts (all produced/consumed by this library only)
 *   NCHW   fp32 [N][C][H][W]                      -- the reference's layout, used at the module boundary
 *   ACT    bf16 [N][C/8][planes][H][W][8]         -- conv operand; planes=2 stores x = hi + lo (bf16x3
 *                                                    split precision, ~fp32-equivalent), planes=1 is plain bf16
 *   F32B   fp32 [N][C/8][H][W][8]                 -- channel-blocked fp32 (encoder residual stream, IN inputs)
 *   WPK    bf16 [taps][Cin/8][planes][Cout][8]    -- packed conv weights (K-major, 16-byte K chunks)
 *   C must be a multiple of 16 for ACT/F32B/WPK tensors.
 */
#ifndef DGE_B200_H_
#define DGE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGE_OK 0
#define DGE_ERR_BAD_ARG (-1)
#define DGE_ERR_CUDA (-2)
#define DGE_ERR_UNSUPPORTED (-3)

/* ---- library ------------------------------------------------------------------------------- */
const char* dge_last_error(void);
int dge_version(void);
/* 0 if the current device is sm_100 (B200); DGE_ERR_UNSUPPORTED otherwise. */
int dge_device_ok(void);
/* number of kernels launched by this library in this process since the last reset (bench.py gpu_launches); process-wide:
   torch.autograd runs the backward nodes on its own worker thread */
int64_t dge_launch_count(void);
void dge_launch_count_reset(void);

/* ---- conv: tcgen05 implicit-GEMM with fused epilogue ------------------------------------------ */
enum {
  DGE_CONV_3X3 = 0,    /* 3x3, stride 1, pad 1  (stylegan2_generator.py:897-904; E.py:59,72; lreq.py:126-156) */
  DGE_CONV_1X1 = 1,    /* 1x1                  (E.py:81-82 conv_3; net.py:231-240 FromRGB when Cin%16==0) */
  DGE_CONV_UP3X3 = 2,  /* 3x3 transposed, stride 2, pad 0 -> raw (2H+1)x(2W+1) map, before the FIR
                          (stylegan2_generator.py:879-895) */
  DGE_CONV_DOWN4X4S2 = 3 /* 4x4, stride 2, pad 1 (the `transform_kernel` strided conv of model/E/E_Blur.py:32-33,72;
                            lreq.py:144-156).  x is the SPACE-TO-DEPTH input: ACT [n][4*C/8][planes][h][w][8] at the
                            OUTPUT resolution, channel block 2*py+px holding xin[2y+py][2x+px]
                            (dge_instance_norm_blur writes it); cin = 4*C; wpk has 16 taps of C channels. */
};
enum {
  DGE_CONV_FLAG_CHECKER = 1  /* run the slow CUDA-core checker kernel instead of tcgen05 (tests only) */
};

typedef struct dge_conv_args {
  int32_t kind;            /* DGE_CONV_* */
  int32_t flags;
  int32_t n, h, w;         /* input batch / height / width */
  int32_t cin, cout;
  int32_t planes;          /* 1 or 2 (see ACT) -- must match x and wpk */
  const void* x;           /* ACT  [n][cin/8][planes][h][w][8] */
  const void* wpk;         /* WPK  [taps][cin/8][planes][cout][8] */
  /* pointwise epilogue, applied in this order (NULL / 0 = skipped):
       v  = acc * demod[n][co]
       v += noise[n*noise_bstride + y*W + x] * (noise_w ? noise_w[co] : noise_scalar)
       v += bias[co]
       v += preact_add[n][co][y][x]                   (F32B residual added BEFORE the activation, E_PG.py:100)
       v  = (v < 0 ? v*slope : v) * gain
       v  = blend_a * S(blend_src) + blend_b * v      (S = 2x2 mean of a double-resolution F32B tensor
                                                        if blend_pool, else the same-resolution value) */
  const float* demod;      /* [n][cout] */
  const float* noise;      /* fp32, [h][w] (bstride 0) or [n][h][w] */
  int64_t noise_bstride;
  const float* noise_w;    /* [cout] */
  float noise_scalar;
  const float* bias;       /* [cout] */
  float slope, gain;
  const float* blend_src;  /* F32B */
  int32_t blend_pool;
  float blend_a, blend_b;
  /* outputs (any subset) */
  void* out_act;           /* ACT [n][cout/8][out_planes][h][w][8], multiplied by out_scale[n][co] if given */
  int32_t out_planes;
  const float* out_scale;  /* [n][cout] -- the NEXT layer's style (modulation folded into the producer) */
  float* out_f32b;         /* F32B */
  float* out_nchw;         /* NCHW fp32 */
  const float* rgb_w;      /* [n][3][cout]: fused ToRGB weights (style and wscale folded in) */
  float* rgb_out;          /* NCHW [n][3][h][w]; contributions are atomically ADDED (pre-initialise with
                              dge_rgb_init) -- stylegan2_generator.py:515-522 */
  float* out_raw_up;       /* DGE_CONV_UP3X3 only: F32B-like [n][cout/8][2h+1][2w+1][8] raw transposed conv */
  const float* preact_add; /* F32B [n][preact_c/8][h/preact_up][w/preact_up][8] or NULL */
  int32_t preact_c;        /* channels of the residual tensor (0 = cout); only the first cout are used (BigGAN
                              GenBlock channel drop, biggan_generator.py:195-197) */
  int32_t preact_up;       /* 2 = the residual is half resolution, read with nearest x2 (:198-199); else 1 */
  float* splitk_ws;        /* optional split-K scratch of dge_conv_splitk_ws_bytes(args) bytes (NULL: never split K).
                              Small maps (<= 16x16 at batch 8) divide the K loop over several CTAs per tile and sum the
                              parts here before the epilogue; the call zeroes it. */
  int32_t out_pool;        /* 1: out_f32b is [n][cout/8][h/2][w/2][8] and receives the 2x2 MEAN of the epilogue value
                              (avg_pool2d(2,2) fused into the producer: model/E/E.py:78-84 only ever reads conv_2's
                              output pooled).  Needs even h, w; out_f32b must be the only output. */
  int32_t in_h, in_w;      /* 0 = h, w.  Otherwise the stored extent of x when it is LARGER than the output domain h x w
                              (rows / columns beyond the domain are read by the halo instead of zero padding): the
                              stride-2 data gradient of the x2 layer reads an (h+1) x (w+1) space-to-depth map
                              (dge_up_fir_bwd_s2d) and produces h x w.  in_h >= h, in_w >= w. */
} dge_conv_args;

int dge_conv_forward(const dge_conv_args* a, void* stream);
/* bytes of scratch dge_conv_forward would use for this problem through args->splitk_ws (0: split-K not applicable) */
size_t dge_conv_splitk_ws_bytes(const dge_conv_args* a);

/* ---- weight preparation ---------------------------------------------------------------------- */
/* OIHW fp32 [cout][cin][k][k] -> WPK.  flip=1 packs the spatially flipped kernel (transposed conv,
   stylegan2_generator.py:880).  scale multiplies every weight (wscale, :858). */
int dge_pack_conv_weight(const float* w_oihw, void* wpk, int cout, int cin, int ksize, int flip,
                         float scale, int planes, void* stream);
/* Data-gradient operand of the same layer: OIHW fp32 [cout][cin][k][k] -> WPK [taps][cout/8][planes][cin][8] holding
   W'[i][o][ky][kx] = W[o][i][k-1-ky][k-1-kx]*scale.  dge_conv_forward(dy as ACT with `cout` channels, this operand,
   cout := cin) then computes dL/dx of y = conv2d(x, W, pad (k-1)/2) (torch.nn.functional.conv2d backward w.r.t. input;
   lreq.py:126-156, stylegan2_generator.py:897-904) -- the first building block of the training step. */
int dge_pack_conv_weight_dgrad(const float* w_oihw, void* wpk, int cout, int cin, int ksize, float scale, int planes,
                               void* stream);
/* Weight gradient of y = conv2d(x, W, padding k/2), stride 1, k in {1, 3} (autograd of lreq.py:126-156 as driven by
   E_align_s2.py:205-233): dw[o][i][ky][kx] (+)= sum_{n,y,x} dy[n][o][y][x] * x[n][i][y+ky-k/2][x+kx-k/2].
   dy_act / x_act: ACT tensors ([n][c/8][planes][h][w][8] bf16) with cout / cin channels and the same `planes`
   (2 = split precision hi+lo, ~fp32 result; 1 = plain bf16).  dw: fp32 [cout][cin][k][k] (the layout of the
   parameter's .grad).  accumulate=0 zeroes dw first; 1 adds to it.  tcgen05 kernel, contraction over pixels with both
   operands read as MN-major tiles straight from the ACT layout; partial sums merge with fp32 atomics (run-to-run
   differences at the 1e-7 relative level).  cout, cin multiples of 8. */
int dge_conv_wgrad(const void* dy_act, const void* x_act, float* dw, int n, int cout, int cin, int h, int w, int ksize,
                   int planes, int accumulate, void* stream);
/* W2[o][i] = sum_k (w[o][i][k]*scale)^2 -- the demodulation Gram diagonal (stylegan2_generator.py:867-870) */
int dge_weight_sqsum(const float* w_oihw, float* w2, int cout, int cin, int ksize, float scale, void* stream);
/* d[n][o] = rsqrt(sum_i W2[o][i]*s[n][i]^2 + eps) */
int dge_demod(const float* w2, const float* style, float* d, int n, int cout, int cin, float eps, void* stream);
/* rgb_w[n][ch][c] = w[ch][c]*scale*style[n][c]   (ToRGB 1x1 modulated conv without demod, :462-474) */
int dge_rgb_weights(const float* w, const float* style, float* rgb_w, int n, int nch, int cin, float scale,
                    void* stream);

/* ---- all per-layer scalars of one StyleGAN2 synthesis pass in ONE launch ---------------------- */
/* For every layer: style = Dense(wp[:, wp_index]) (stylegan2_generator.py:872-877, 990-996), then either the
   demodulation coefficients d[n][o] = rsqrt(sum_i W2[o][i]*style[n][i]^2 + eps) (:867-870) or the ToRGB weights
   rgbw[n][ch][c] = w[ch][c]*scale*style[n][c] (:462-474).  Replaces ~53 tiny launches (dge_dense / dge_demod /
   dge_rgb_weights per layer) of a 1024^2 pass.  Outputs go to one caller-provided arena at the given float offsets
   (a negative offset = output not wanted).  `items` is a DEVICE array. */
typedef struct dge_sg2_prep_item {
  const float* st_w;       /* style weight [cin][wdim] */
  const float* st_b;       /* style bias [cin] or NULL */
  const float* w2;         /* [cout][cin] from dge_weight_sqsum, or NULL */
  const float* rgb_w;      /* ToRGB weight [nch][cin], or NULL */
  int64_t style_off;       /* -> arena: style [n][cin] */
  int64_t demod_off;       /* -> arena: demod [n][cout] */
  int64_t rgbw_off;        /* -> arena: rgb weights [n][nch][cin] */
  int32_t wp_index;
  int32_t cin, cout, nch;
  float st_wscale, st_bscale, st_add_bias, rgb_scale, eps;
  int32_t pad_;
} dge_sg2_prep_item;
int dge_sg2_prep(const dge_sg2_prep_item* items, int n_items, const float* wp, float* arena, int n, int num_layers,
                 int wdim, void* stream);

/* Transpose of dge_sg2_prep for the training step (the frozen generator's gradient w.r.t. wp, E_align_s2.py:160-205): the
   per-layer reductions dge_sg2_layer_bwd left in `sums` -> d_wp [n][num_layers][wdim] (zeroed by the call), one launch.
   `items`, `arena`: what dge_sg2_prep was given / filled (styles, demods).  gsrc: DEVICE array [n_items][4] of float offsets
   into `sums`, (s_off, s_stride, d_off, t_off), -1 = absent:
     layer item : S[c] = sums[s_off + (n*cin + c)*s_stride]          (d style),  D[o] = sums[d_off + (n*cout + o)*5] (demod * d demod)
                  ds[c] = S[c] - style[n][c] * sum_o D[o] * demod[n][o]^2 * w2[o][c]           (:867-877)
     ToRGB item : T[c][j] = sums[t_off + (n*cin + c)*5 + j],  ds[c] = rgb_scale * sum_j T[c][j] * rgb_w[j][c]   (:462-474)
     d_wp[n][wp_index][k] += st_wscale * sum_c ds[c] * st_w[c][k]                                   (:990-996 transposed) */
int dge_sg2_prep_bwd(const dge_sg2_prep_item* items, int n_items, const float* arena, const int64_t* gsrc,
                     const float* sums, float* d_wp, int n, int num_layers, int wdim, void* stream);

/* ---- dense (DenseBlock.forward stylegan2_generator.py:990-996; ln.Linear lreq.py:68-75) ------ */
/* y[n][m] = act((sum_k x[n][k]*w[m][k])*wscale + b[m]*bscale + add_bias) * gain ; act = lrelu(slope) */
int dge_dense(const float* x, const float* w, const float* b, float* y, int n, int k, int m, float wscale,
              float bscale, float add_bias, float slope, float gain, void* stream);
/* PixelNormLayer (stylegan2_generator.py:550-553) on [n][k] */
int dge_pixel_norm(const float* x, float* y, int n, int k, float eps, void* stream);

/* ---- layout / elementwise --------------------------------------------------------------------- */
/* NCHW fp32 -> ACT, optionally times scale[n][c]; x_bstride = 0 broadcasts one sample (InputBlock :630-632) */
int dge_nchw_to_act(const float* x, int64_t x_bstride, const float* scale, void* act, int n, int c, int h, int w,
                    int planes, void* stream);
int dge_nchw_to_f32b(const float* x, float* out, int n, int c, int h, int w, void* stream);
int dge_f32b_to_nchw(const float* x, float* out, int n, int c, int h, int w, void* stream);
int dge_act_to_nchw(const void* act, float* out, int n, int c, int h, int w, int planes, void* stream);
int dge_f32b_to_act(const float* x, void* act, int n, int c, int h, int w, int planes, void* stream);
/* standalone ToRGB on NCHW input: out[n][ch] = bias[ch] + sum_c x[n][c]*rgb_w[n][ch][c]
   (ModulateConvBlock k=1, demodulate=False, stylegan2_generator.py:462-474) */
int dge_to_rgb_nchw(const float* x, const float* rgb_w, const float* bias, float* out, int n, int c, int nch, int h,
                    int w, void* stream);

/* 4x4 FIR ([1,3,3,1]x[1,3,3,1]/16, pad 1) over the raw up-conv map + demod + noise + bias + lrelu*gain,
   output ACT times out_scale (stylegan2_generator.py:603-615 with filter padding (1,1,1,1), :907-921) */
int dge_up_fir_epilogue(const float* raw_up, const float* demod, const float* noise, int64_t noise_bstride,
                        float noise_scalar, const float* bias, float slope, float gain, const float* out_scale,
                        void* out_act, float* out_nchw, int n, int c, int h_out, int w_out, int planes,
                        void* stream);
/* img_out[n][ch][2h][2w] = bias[ch] + up2_fir(img_in)  (UpsamplingLayer(scale 2), :603-615, skip sum :519-522);
   img_in == NULL: img_out = bias only (first resolution, h_out x w_out given directly) */
int dge_rgb_init(const float* img_in, const float* bias, float* img_out, int n, int nch, int h_out, int w_out,
                 void* stream);

/* ---- encoder pieces (model/E/E.py:50-85) ------------------------------------------------------- */
/* FromRGB: 1x1 conv (cin=3) + bias + lrelu(0.2): NCHW image -> F32B  (model/utils/net.py:231-240) */
int dge_from_rgb(const float* img, const float* w, const float* b, float* out_f32b, int n, int cimg, int c, int h,
                 int wd, float slope, void* stream);
/* per-(n,c) mean and biased std over H*W of an F32B tensor: style[n][0:c]=mean, style[n][c:2c]=std
   (E.py:51-53), mean_rstd[n][c] = (mean, 1/sqrt(var+eps)) for the instance norm (E.py:23,58).
   `scratch` = 2*n*c doubles, zeroed by the call. */
int dge_instance_stats(const float* x_f32b, double* scratch, float* style, float* mean_rstd, int n, int c, int h,
                       int w, float eps, void* stream);
/* instance norm apply: F32B -> ACT (conv operand) and/or F32B */
int dge_instance_norm(const float* x_f32b, const float* mean_rstd, void* out_act, float* out_f32b, int n, int c,
                      int h, int w, int planes, void* stream);
/* as dge_instance_norm with the affine InstanceNorm2d(affine=True) weight/bias (E_PG.py:59,99): gamma/beta [c] or NULL */
int dge_instance_norm_affine(const float* x_f32b, const float* mean_rstd, const float* gamma, const float* beta,
                             void* out_act, float* out_f32b, int n, int c, int h, int w, int planes, void* stream);
/* instance norm, then the depthwise 3x3 Blur ([1,2,1]^2/16, zero padding) of model/E/E_Blur.py:71 -> ACT.
   s2d = 0: ACT [n][c/8][planes][h][w][8];  s2d = 1: space-to-depth ACT [n][4c/8][planes][h/2][w/2][8] for DGE_CONV_DOWN4X4S2 */
int dge_instance_norm_blur(const float* x_f32b, const float* mean_rstd, void* out_act, int s2d, int n, int c, int h,
                           int w, int planes, void* stream);
/* instance norm (as dge_instance_norm, F32B -> ACT) fused with the 2x2 average pool of the RAW input -> ACT at
   (h/2, w/2): one pass over x feeds conv_1's operand (E.py:58) and the residual branch (E.py:78) */
int dge_instance_norm_pool(const float* x_f32b, const float* mean_rstd, void* out_act, void* out_pool_act, int n, int c,
                           int h, int w, int planes, void* stream);
/* dge_from_rgb that also produces the instance statistics of its output (style = mean||std, mean_rstd as
   dge_instance_stats; `scratch` = 2*n*c doubles): net.py:231-240 + E.py:51-53,58.  c must be 16 or 32. */
int dge_from_rgb_stats(const float* img, const float* w, const float* b, float* out_f32b, double* scratch, float* style,
                       float* mean_rstd, int n, int cimg, int c, int h, int wd, float slope, float eps, void* stream);
/* 2x2 average pool F32B -> ACT (residual branch, E.py:78) */
int dge_avgpool_to_act(const float* x_f32b, void* out_act, int n, int c, int h, int w, int planes, void* stream);
/* out = a*A' + b*B' where X' = 2x2 mean of a double-resolution tensor if its pool bit is set (bit 0: A, bit 1: B),
   else X; all F32B (E.py:76-84 when Cin==Cout: pool = 3; E_Blur.py fused blocks: pool = 2) */
int dge_blend(const float* a_src, const float* b_src, float* out, float a, float b, int pool, int n, int c,
              int h_out, int w_out, void* stream);

/* ---- StyleGAN1 pieces (model/stylegan1/net.py:32-58, 141-169, 244-253) ------------------------------ */
/* filter + noise + bias + lrelu between a conv and the next instance norm.
   mode 0: src F32B [n][c/8][h][w][8] -> 3x3 Blur ([1,2,1]^2/16, zero pad)               (Blur, :48-58)
   mode 1: src raw transposed-conv map [n][c/8][h+1][w+1][8] (dge_conv_forward UP3X3) -> 2x2 box sum (the
           `transform_kernel` 4-shift sum with stride 2 / padding 1, lreq.py:127-131) -> the same Blur
   mode 2: no filtering (first block, x = const)
   then v = v + noise_w[c]*noise[n][y][x] + bias[c]; lrelu(slope)                           (:148-152) */
int dge_sg1_post(const float* src, int mode, const float* noise, const float* noise_w, const float* bias, float slope,
                 float* out_f32b, int n, int c, int h_out, int w_out, void* stream);
/* instance norm + style_mod: y = (x-mean)*rstd*(style[n][c]+1) + style[n][C+c]  (:32-34,154-156); in_n == 1 broadcasts a
   single input sample over the batch (x = const); up = 2 writes the nearest-upsampled result (upscale2d :37-43) */
int dge_instance_norm_style(const float* x_f32b, int in_n, const float* mean_rstd, const float* style, int up,
                            void* out_act, float* out_f32b, int n, int c, int h, int w, int planes, void* stream);
/* ToRGB: 1x1 conv F32B -> NCHW [n][nch][h][w]  (:244-253) */
int dge_to_rgb_f32b(const float* x_f32b, const float* w, const float* bias, float* out_nchw, int n, int c, int nch,
                    int h, int wd, void* stream);

/* ---- BigGAN pieces (model/biggan_generator.py:58-150, 175-256; model/E/E_BIG.py:33-82) ------------------ */
/* conditional-BN coefficients: A[n][c] = (1+scale[n][c])*rsqrt(var[c]+eps), B[n][c] = offset[n][c] - mean[c]*A[n][c]
   (:143-148).  scale/offset NULL => plain BN with weight/bias [c] (:150): A = weight*rsqrt(var+eps), B = bias - mean*A. */
int dge_cbn_coeffs(const float* scale, const float* offset, const float* weight, const float* bias, const float* mean,
                   const float* var, float eps, float* a_out, float* b_out, int n, int c, void* stream);
/* y = act(x*A[n][c] + B[n][c]), act = relu if relu != 0; optional nearest x2; F32B -> ACT and/or F32B (:178-190) */
int dge_affine_act(const float* x_f32b, const float* a, const float* b, int relu, int up, void* out_act,
                   float* out_f32b, int n, int c, int h, int w, int planes, void* stream);
/* nn.MaxPool2d(2, 2) on F32B (SelfAttn, :82,91) */
int dge_maxpool2_f32b(const float* x, float* out, int n, int c, int h_out, int w_out, void* stream);
/* softmax over the CHANNEL axis of an F32B tensor -> ACT (attention rows: channels = keys, :85-86) */
int dge_channel_softmax_to_act(const float* x_f32b, void* out_act, int n, int c, int h, int w, int planes, void* stream);
/* out[n][0:nch] = tanh(x[n][0:nch]) from NCHW [n][c][h][w] (:250-253) */
int dge_tanh_slice_nchw(const float* x, float* out, int n, int c, int nch, int hw, void* stream);

/* ---- PGGAN pieces (model/pggan/pggan_generator.py:214-216, 230-233, 319-339) ------------------- */
/* PixelNormLayer over channels + optional nearest x2 upsample: F32B [n][c/8][h][w][8] -> ACT at (h*up, w*up) */
int dge_pixelnorm_to_act(const float* x_f32b, void* out_act, int n, int c, int h, int w, int up, float eps, int planes,
                         void* stream);
/* output ConvBlock: pixel-norm + 1x1 conv (w [nch][c], already times wscale) + bias -> NCHW [n][nch][h][w] */
int dge_pixelnorm_to_rgb(const float* x_f32b, const float* w, const float* bias, float* out_nchw, int n, int c, int nch,
                         int h, int wd, float eps, void* stream);
/* F.interpolate(scale_factor=2, mode='nearest') on NCHW */
int dge_upsample_nearest_nchw(const float* x, float* out, int64_t planes, int h, int w, void* stream);
/* out = a*x + b*y over n floats */
int dge_axpby(const float* x, const float* y, float* out, float a, float b, int64_t n, void* stream);

/* ---- losses (training_utils.py:54-99 space_loss; metric/pytorch_ssim.py:18-38) ------------------ */
/* out6 (zeroed by the call) = sum a, sum b, sum a^2, sum b^2, sum a*b, sum (a-b)^2 over n elements:
   MSE, mean/std MSE terms and the whole-batch cosine all derive from these (training_utils.py:63-75). */
int dge_pair_moments(const float* a, const float* b, int64_t n, double* out6, void* stream);
/* sum over all elements of softmax(a)*(log softmax(a) - log softmax(b)), softmax over the axis of size d with
   stride `inner` (torch's implicit-dim softmax + KLDivLoss, training_utils.py:68-69) */
int dge_softmax_kl_sum(const float* a, const float* b, int64_t outer, int d, int64_t inner, double* out1, void* stream);
/* factor x factor mean pooling of [planes][h_out*factor][w_out*factor] (avg_pool2d(2,2) loop, :81-84) */
int dge_avgpool_nchw(const float* x, float* out, int64_t planes, int h_out, int w_out, int factor, void* stream);
/* sum of the SSIM map (11x11 gaussian sigma 1.5, zero padding, C1=1e-4, C2=9e-4) over [planes][h][w] */
int dge_ssim_sum(const float* a, const float* b, int64_t planes, int h, int w, double* out1, void* stream);

/* Gradient of mean(SSIM map) w.r.t. the SECOND image (the one the generator produced; pytorch_ssim.py:18-38 under
   loss.backward(), training_utils.py:87-88): db[planes][h][w] = go[0] * d mean(ssim(a, b)) / d b.  go: one device float (the
   upstream gradient); scratch3: 3 * planes*h*w floats. */
int dge_ssim_grad(const float* a, const float* b, const float* go, float* scratch3, float* db, int64_t planes, int h, int w,
                  void* stream);

/* ---- Grad-CAM (metric/grad_cam.py:101-194) ---------------------------------------------------- */
/* idx_out[n] = np.argmax(logits, axis=1) (first maximum); *mode_out = np.argmax(np.bincount(idx)) (:164-165). Bit-exact. */
int dge_argmax_mode(const float* logits, int n, int k, int64_t* idx_out, int64_t* mode_out, void* stream);
/* CAM maps from the hooked layer: feature / gradient NCHW fp32 [n][c][h][w] -> out float64 [n][h_out][w_out].
   plus = 1: Grad-CAM++ as coded (:169-190, float64): w_c = sum(relu(g_c) * 1/sum(relu(g_c))), cam = sum_c f_c*w_c,
             cam -= min, cam /= max, cv2.resize bilinear to (w_out, h_out);
   plus = 0: Grad-CAM (:113-126, float32): w_c = mean(g_c), relu(cam), same normalisation / resize.
   w_scratch: n*c doubles, cam_scratch: n*h*w doubles. */
int dge_gradcam(const float* feature, const float* gradient, int plus, double* w_scratch, double* cam_scratch,
                double* out, int n, int c, int h, int w, int h_out, int w_out, void* stream);

/* mask2cam (:234-251): heat = JET(uint8(255*mask)) / 255 (RGB, `lut_rgb` = the 256x3 colour table), cam = heat + img,
   then per image i: cam[i] -= min(cam) (whole array, as upstream), cam[i] /= max(cam[i]).  mask float64 [n][h][w],
   img / heat / cam fp32 NCHW [n][3][h][w]; scratch4 = 4 floats. */
int dge_mask2cam(const double* mask, const float* img, const float* lut_rgb, float* heat, float* cam, float* scratch4,
                 int n, int h, int w, void* stream);

/* ---- training step: point-wise / reduction backward kernels ----------------------- */
/* These replace what torch.autograd runs for `loss.backward()` in E_align_s2.py:205,220 between the convolutions of
   model/E/E.py:50-85.  Gradients arrive as F32B maps (what a data-gradient dge_conv_forward wrote) and leave as the ACT
   operand of the next contraction; per-channel parameter gradients are reduced in the same pass.  Every `sums` output
   is zeroed by the call and accumulated with atomics (run-to-run differences at the 1e-7 relative level). */

/* Tail of an encoder block, backward (E.py:72-84: out = ga*avg_pool2d(lrelu(conv_2 + nw2*noise + b2)) + gb*residual).
   d_out F32B [n][co/8][h/2][w/2][8]; y2 F32B [n][co/8][h][w][8] = the ACTIVATED conv_2 output (its sign is the
   pre-activation's); noise [n][h][w] or NULL.
   dy2_act  ACT [n][co/8][planes][h][w][8]     = ga/4 * d_out[y/2][x/2] * (y2 > 0 ? 1 : slope)   -> dgrad / wgrad of conv_2
   dres_act ACT [n][co/8][planes][h/2][w/2][8] = gb * d_out  (NULL: not wanted)                   -> dgrad / wgrad of conv_3
   sums fp32 [3][co]: sum dy2 (bias_2.grad), sum dy2*noise (noise_weight_2.grad), sum gb*d_out (conv_3.bias.grad) */
int dge_be_head_bwd(const float* d_out, const float* y2, const float* noise, float ga, float gb, float slope,
                    void* dy2_act, void* dres_act, float* sums, int n, int co, int h, int w, int planes, void* stream);
/* Instance-norm backward, reduction pass (nn.InstanceNorm2d, E.py:58,69): sums fp64 [n][c][2] = (sum g, sum g*xn) over
   h*w with xn = (x - mean)*rstd; g, x F32B; mean_rstd as written by dge_instance_stats. */
int dge_in_bwd_stats(const float* g, const float* x, const float* mean_rstd, double* sums, int n, int c, int h, int w,
                     void* stream);
/* Instance-norm backward, apply pass.  dx = rstd*(g - mean(g) - xn*mean(g*xn))  [sums from dge_in_bwd_stats]
                                           + dstyle[n][c]/HW + dstyle[n][C+c]*(x - mean)/(HW*std)
   where style = mean || std [n][2c] is the block's second output (E.py:51-54, 64-67) and dstyle its gradient (NULL: none).
   mode 0: dx += rscale * res  (res F32B at the same resolution, or half resolution read through the 2x2-pool broadcast when
           res_pool; NULL: none) -> out_f32b                                            (block input; E.py:78-84 residual)
   mode 1: dx *= (x > 0 ? 1 : slope) -> out_act [n][c/8][planes][h][w][8]; sums2 fp32 [2][c] = sum dx, sum dx*noise
           (x = lrelu(conv_1 + nw1*noise + b1), E.py:60-62: bias_1.grad and noise_weight_1.grad)
   gscale [n][c] (NULL = 1): the incoming gradient is gscale * g -- the `x*(s0+1)+s1` style modulation that follows the
   instance norm in the StyleGAN1 generator (stylegan1/net.py:32-34, 154-156), sums being those of the un-scaled g; mode 1
   may also (or instead) write the result as F32B (out_f32b). */
int dge_in_bwd_apply(const float* g, const float* x, const float* mean_rstd, const float* style, const float* dstyle,
                     const double* sums, const float* gscale, int mode, const float* res, float rscale, int res_pool,
                     const float* noise, float slope, float* out_f32b, void* out_act, float* sums2, int n, int c, int h,
                     int w, int planes, void* stream);
/* Conditional batch norm + (leaky) ReLU, backward: the frozen BigGAN generator under loss.backward()
   (biggan_generator.py:138-150 with :178-190 -- t = relu(a*x + b) [-> nearest x2] -> conv; a, b fp32 [n][c] come from the
   condition vector, so their gradients are what the encoder's z receives through every block norm).
   g F32B [n][c/8][h*up][w*up][8]: gradient of the conv's input; x F32B [n][c/8][h][w][8]: the norm's input.
     d = (a*x + b > 0 ? 1 : slope) * sum_{up x up} g;   sums fp32 [n][c][2] = (sum d*x, sum d) = (d a, d b)   (zeroed by the call)
     dx = a*d, plus for channels < skip_c the sum over skip_up x skip_up of skip F32B [n][skip_c/8][h*skip_up][w*skip_up][8]
          (GenBlock's identity branch :192-203: channel drop + nearest up-sampling; NULL: none)
   -> out_f32b [n][c/8][h][w][8] and / or out_act [n][c/8][planes][h][w][8] (operand of the next data-gradient conv). */
int dge_affine_relu_bwd(const float* g, const float* x, const float* a, const float* b, float slope, int up,
                        const float* skip, int skip_c, int skip_up, float* out_f32b, void* out_act, float* sums, int n,
                        int c, int h, int w, int planes, void* stream);
/* FromRGB backward (net.py:231-240, f = lrelu(conv1x1(img, W) + b)): sums fp32 [c][4] = (dW[c][0..2], db[c]) with
   d_pre = d_f * (f > 0 ? 1 : slope); d_f, f F32B [n][c/8][h][w][8]; img NCHW [n][cimg<=3][h][w].
   d_img (optional, NCHW like img; zeroed by the call): the image gradient sum_c wgt[c][i] * d_pre[c] -- the inversion loop
   of embedding_img.py:86-88 back-propagates through E(imgs2) into the generator; wgt = the 1x1 weight [c][cimg]. */
int dge_from_rgb_bwd(const float* d_f, const float* f, const float* img, const float* wgt, float slope, float* sums,
                     float* d_img, int n, int cimg, int c, int h, int w, void* stream);

/* Backward of everything that follows the contraction in a StyleGAN2 synthesis layer (stylegan2_generator.py:907-921:
   y = lrelu(conv*dm + noise*ns + b)*gain), plus the two consumers of y: the next layer (y * s_next, :877) and ToRGB
   (sum_c rgbw[n][ch][c]*y[c], :462-474, 515-522).  The forward keeps y only as the next layer's operand
   ya = y * ya_scale (ACT [n][c/8][planes][h][w][8]; ya_scale NULL: ya = y), so y = ya / ya_scale.
     dy = dxs * ya_scale + sum_ch rgbw[n][ch][c] * dimg[n][ch]     dxs  F32B [n][c/8][h][w][8] (gradient w.r.t. the next
                                                                    layer's operand; NULL: none), dimg NCHW [n][3][h][w]
     d_conv = dy * gain * lrelu'(y) * demod[n][c]  -> out_act (ACT, operand of the data-gradient conv) and/or out_f32b
     sums fp32 [n][c][5] = S  = sum dxs*y                   (gradient of ya_scale = the next layer's style)
                           T0..T2 = sum dimg[ch]*y          (gradient of rgbw[n][ch][c])
                           D  = sum d_pre*(pre - noise*ns - b), d_pre = dy*gain*lrelu', pre = the pre-activation
                                                            (demod[n][c] * gradient of demod[n][c])
   noise [h][w] (bstride 0) or [n][h][w]; noise_scalar = the layer's noise strength; bias [c] (already times bscale). */
int dge_sg2_layer_bwd(const void* ya_act, const float* ya_scale, const float* dxs, const float* dimg, const float* rgbw,
                      const float* noise, int64_t noise_bstride, float noise_scalar, const float* bias,
                      const float* demod, float gain, float slope, void* out_act, int out_planes, float* out_f32b,
                      float* sums, int n, int c, int h, int w, int planes, void* stream);
/* Transpose of the x2 layer's 4x4 FIR (dge_up_fir_epilogue; :603-615) written as the space-to-depth operand of the
   stride-2 data-gradient conv: dconv F32B [n][c/8][2h][2w][8] -> ACT [n][4c/8][planes][h+1][w+1][8] with channel block
   (2py+px)*c/8 + g holding dt[2Y+py][2X+px], dt[u][v] = sum_{a,b<4} f[a]f[b] dconv[u-a+1][v-b+1] (f = [1,3,3,1]/4), zeros
   beyond the (2h+1) x (2w+1) raw map.  dge_conv_forward(DGE_CONV_DOWN4X4S2, in_h = h+1, in_w = w+1) with the 3x3 kernel
   placed in rows / columns 1..3 of the 4x4 taps then is the data gradient of the transposed conv (:879-895).
   box != 0: the filter is the 2x2 box sum of the StyleGAN1 `transform_kernel` layers (stylegan1/lreq.py:127-131; forward:
   dge_sg1_post mode 1) instead of the FIR: dt[u][v] = sum_{a,b<2} dconv[u-a][v-b]. */
int dge_up_fir_bwd_s2d(const float* dconv, void* out_act, int n, int c, int h, int w, int planes, int box, void* stream);
/* Transpose of dge_rgb_init's x2 up-sampling of the skip image (:519-522): d_in [planes][h_in][w_in] from
   d_out [planes][2h_in][2w_in]; per axis d_in[m] = (d[2m-1] + 3d[2m] + 3d[2m+1] + d[2m+2])/4. */
int dge_rgb_up_bwd(const float* d_out, float* d_in, int64_t planes, int h_in, int w_in, void* stream);

/* ---- LPIPS-VGG16 perceptual distance (third-party `lpips.LPIPS(net='vgg')`, E_align_s2.py:98; consumed by
        training_utils.py:93) -- the pieces between the VGG convolutions; the convs are dge_conv_forward with the bias and the
        ReLU (slope 0) in the epilogue, forward and data gradient.  PARITY UNPINNED (no reference vectors offline). ---- */
/* ScalingLayer + channel padding: NCHW [n][3][h][w] -> ACT [n][16/8][planes][h][w][8] = (x - shift[c]) / scale[c], channels
   3..15 zero (the first VGG conv then runs on the tensor cores with a zero-padded [64][16][3][3] weight). */
int dge_lpips_input(const float* x, void* out_act, float shift0, float shift1, float shift2, float scale0, float scale1,
                    float scale2, int n, int h, int w, int planes, void* stream);
/* nn.MaxPool2d(2, 2), floor mode: F32B [n][c/8][h][w][8] -> ACT [n][c/8][planes][h/2][w/2][8] (the next conv's operand). */
int dge_maxpool_to_act(const float* x_f32b, void* out_act, int n, int c, int h, int w, int planes, void* stream);
/* Backward of [ReLU -> MaxPool2d(2,2)] into the ACT operand of the conv's data gradient:
     out = (y > 0) * (g_same + (pixel is the first maximum of its 2x2 window ? g_pool[y/2][x/2] : 0))
   y = the activated conv output: F32B (y_f32b) or ACT (y_act, y_planes: hi (+ lo) planes; an ACT holds ~17 bits, so window
   maxima closer than that may route differently from the fp32 map -- the LPIPS node keeps its pooled taps in F32B) -- exactly one;
   g_same F32B [n][c/8][h][w][8] or NULL; g_pool F32B [n][c/8][h/2][w/2][8] or NULL (at least one).
   clamp_pos != 0: out = max(out, 0) -- the ReLU backward hook of GuidedBackPropagation (metric/grad_cam.py:209-211). */
int dge_relu_pool_bwd(const float* y_f32b, const void* y_act, int y_planes, const float* g_same, const float* g_pool,
                      void* out_act, int n, int c, int h, int w, int planes, int clamp_pos, void* stream);
/* One LPIPS tap: f F32B [2*nb][c/8][h][w][8] holds the features of the two image batches (a = f[0:nb], b = f[nb:2nb]).
   forward  (ga == gb == NULL): out[n] += mean_p sum_c lin_w[c] * (a_c/(|a|+eps) - b_c/(|b|+eps))^2      (out: caller-zeroed)
   backward (go = upstream gradient [nb]): gb / ga (F32B [nb][c/8][h][w][8], either may be NULL) = d out / d b, d out / d a. */
int dge_lpips_dist(const float* f, const float* lin_w, float* out, const float* go, float* ga, float* gb, int nb, int c,
                   int h, int w, float eps, void* stream);

/* ---- optimiser (model/utils/custom_adam.py:24-76, LREQAdam.step) ------------------------------ */
/* One multi-tensor launch:  v = beta2*v + (1-beta2)*g*g ;  p -= step[t]*g/(sqrt(v)+eps)   (beta1 == 0).
   params/grads/vs: DEVICE arrays of n_tensors device pointers; numel/step: per-tensor DEVICE arrays
   (step[t] = lr*sqrt(1-beta2^t_t)*lr_equalization_coef_t); block b updates tensor blk_tensor[b], elements
   [blk_off[b], blk_off[b]+chunk). */
int dge_lreq_adam_step(void* const* params, const void* const* grads, void* const* vs, const int64_t* numel,
                       const float* step, const int32_t* blk_tensor, const int64_t* blk_off, int n_blocks, int chunk,
                       float beta2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGE_B200_H_ */
