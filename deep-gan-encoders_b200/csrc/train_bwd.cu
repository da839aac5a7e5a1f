// train_bwd.cu -- the HBM-bound backward kernels of the training step: everything that sits between the
// tensor-core contractions (dge_conv_forward as data gradient, dge_conv_wgrad) when E_align_s2.py:205,220 calls
// loss.backward().  Each kernel is ONE pass over its tensors: it consumes the fp32 gradient a data-gradient conv wrote
// (F32B), applies the local derivative (leaky-ReLU mask, instance-norm Jacobian, pooling / blend weights, demodulation)
// and writes the bf16 hi/lo ACT operand of the next contraction directly, while the per-channel parameter gradients
// (bias, noise weight, style, demodulation) are reduced in the same pass.
//
//   encoder block  model/E/E.py:50-85        k_be_head_bwd, k_in_bwd_stats, k_in_bwd_apply, k_from_rgb_bwd
//   generator      stylegan2_generator.py:855-922, 515-522, 603-615   k_sg2_layer_bwd, k_up_fir_bwd_s2d, k_rgb_up_bwd
//
// Thread mapping: grid = (splits, n * C/8); a block owns ONE (sample, 8-channel group) and grid-strides over its pixels
// with 32-byte (F32B) / 16-byte (ACT) vector accesses, so per-channel sums reduce inside the block (registers -> warp
// shuffle -> shared memory) and leave it as one atomicAdd per channel and block.
#include "dge_common.cuh"

namespace dge {

constexpr int TB_THREADS = 256;

static int g_tb_sms = 0;
static int tb_sms() {
  if (g_tb_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_tb_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_tb_sms <= 0) g_tb_sms = 148;
  }
  return g_tb_sms;
}
// blocks per (sample, channel group): enough to fill the chip ~12 blocks per SM, never more than the pixels allow
static int tb_splits(long long pixels, long long groups) {
  long long want = ((long long)tb_sms() * 12 + groups - 1) / groups;
  const long long cap = (pixels + TB_THREADS - 1) / TB_THREADS;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return (int)want;
}

// sum NV per-thread values over the block; thread i < NV then holds the total of value i in the return slot
template <int NV>
__device__ __forceinline__ float block_sums(float (&v)[NV], float (*red)[NV]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp][i] = v[i];
  }
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < NV) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
  }
  return t;
}

__device__ __forceinline__ void store8_act_at(void* base, size_t idx16_hi, size_t plane_stride16, int planes,
                                              const float* v) {
  uint4 hi, lo;
  split8(v, hi, lo);
  uint4* p = reinterpret_cast<uint4*>(base);
  p[idx16_hi] = hi;
  if (planes == 2) p[idx16_hi + plane_stride16] = lo;
}

// ---------------------------------------------------------------------------------------------
// encoder block tail, backward (E.py:72-84): out = ga * avgpool2(lrelu(conv_2 + nw2*noise + b2)) + gb * residual
//   d_out F32B [n][co/8][ho][wo][8]  ->  dy2  ACT [n][co/8][planes][2ho][2wo][8] = ga/4 * d_out(pooled idx) * lrelu'(y2)
//                                        dres ACT [n][co/8][planes][ho][wo][8]   = gb * d_out        (optional)
//   sums[0][c] += sum dy2, sums[1][c] += sum dy2 * noise, sums[2][c] += sum dres     (bias_2, noise_weight_2, conv_3.bias)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TB_THREADS, 3)
k_be_head_bwd(const float* __restrict__ d_out, const float* __restrict__ y2, const float* __restrict__ noise,
              float ga4, float gb, float slope, void* __restrict__ dy2, void* __restrict__ dres,
              float* __restrict__ sums, int co, int ho, int wo, int planes) {
  __shared__ float red[TB_THREADS / 32][24];
  const int C8 = co >> 3, ng = blockIdx.y, nidx = ng / C8, g = ng - nidx * C8;
  const int H = 2 * ho, W = 2 * wo;
  const size_t hwp = (size_t)ho * wo, hw = (size_t)H * W;
  float acc[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] = 0.f;
  // (32-bit pixel indices and per-block base pointers: the 64-bit index arithmetic was a fifth of the issued instructions
  //  in these latency-bound loops -- ncu source page of k_sg2_layer_bwd)
  const float* d_out_b = d_out + (size_t)ng * hwp * 8;
  const float* y2_b = y2 + (size_t)ng * hw * 8;
  const float* noise_b = noise ? noise + (size_t)nidx * hw : nullptr;
  const size_t dres_b = (size_t)ng * planes * hwp, dy2_b = (size_t)ng * planes * hw;
  const unsigned uwo = (unsigned)wo, stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)hwp; i += stride) {
    const unsigned py = i / uwo, px = i - py * uwo;
    float d[8], r[8];
    load8_f32b(d_out_b, i, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      r[k] = gb * d[k];
      acc[16 + k] += r[k];
      d[k] *= ga4;
    }
    if (dres) store8_act_at(dres, dres_b + i, hwp, planes, r);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const unsigned pix = (2 * py + dy) * (unsigned)W + (2 * px + dx);
        float y[8], v[8];
        load8_f32b(y2_b, pix, y);
        const float nz = noise_b ? __ldg(noise_b + pix) : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = y[k] > 0.f ? d[k] : d[k] * slope;
          acc[k] += v[k];
          acc[8 + k] = fmaf(v[k], nz, acc[8 + k]);
        }
        store8_act_at(dy2, dy2_b + pix, hw, planes, v);
      }
  }
  const float t = block_sums<24>(acc, red);
  if (threadIdx.x < 24) atomicAdd(sums + (size_t)(threadIdx.x >> 3) * co + g * 8 + (threadIdx.x & 7), t);
}

// ---------------------------------------------------------------------------------------------
// instance-norm backward, pass 1: sums[n][c] = (sum g, sum g * xn), xn = (x - mean) * rstd      (fp64 across threads)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TB_THREADS)
k_in_bwd_stats(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mr,
               double* __restrict__ sums, int c, int hw) {
  __shared__ double red[TB_THREADS / 32][16];
  const int C8 = c >> 3, ng = blockIdx.y, nidx = ng / C8, grp = ng - nidx * C8;
  float m[8], r[8], s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const size_t o = ((size_t)nidx * c + grp * 8 + k) * 2;
    m[k] = __ldg(mr + o);
    r[k] = __ldg(mr + o + 1);
    s1[k] = s2[k] = 0.f;
  }
  const float* g_b = g + (size_t)ng * hw * 8;
  const float* x_b = x + (size_t)ng * hw * 8;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)hw; i += stride) {
    float gv[8], xv[8];
    load8_f32b(g_b, i, gv);
    load8_f32b(x_b, i, xv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s1[k] += gv[k];
      s2[k] = fmaf(gv[k], (xv[k] - m[k]) * r[k], s2[k]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], off);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], off);
    }
    if (lane == 0) {
      red[warp][2 * k] = (double)s1[k];
      red[warp][2 * k + 1] = (double)s2[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0.0;
    for (int w = 0; w < TB_THREADS / 32; ++w) t += red[w][threadIdx.x];
    atomicAdd(&sums[((size_t)nidx * c + grp * 8 + (threadIdx.x >> 1)) * 2 + (threadIdx.x & 1)], t);
  }
}

// ---------------------------------------------------------------------------------------------
// instance-norm backward, pass 2 (E.py:51-58 / 64-69 reversed).  With xn = (x - mean) * rstd, A = mean(g), B = mean(g*xn):
//   dx = rstd * (g - A - xn * B)                     the normalisation itself
//      + dmean / HW + dstd * (x - mean) / (HW * std) the style = (mean, std) the block also emits (inver_mod inputs)
//   mode 0: dx += rscale * res (same resolution, or the 2x2-pool broadcast of a half-resolution map) -> F32B
//           (the residual branch's gradient: E.py:78-84)
//   mode 1: dx *= lrelu'(x) (x is the activated conv output, so its sign is the pre-activation's) -> ACT, and
//           sums2[0][c] += sum dx (bias), sums2[1][c] += sum dx * noise (noise weight)            (E.py:60-62)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TB_THREADS, 3)
k_in_bwd_apply(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mr,
               const float* __restrict__ style, const float* __restrict__ dstyle, const double* __restrict__ sums,
               const float* __restrict__ gscale, int mode, const float* __restrict__ res, float rscale, int res_pool,
               const float* __restrict__ noise, float slope, float* __restrict__ out_f32b, void* __restrict__ out_act,
               float* __restrict__ sums2, int c, int h, int w, int planes) {
  __shared__ float red[TB_THREADS / 32][16];
  const int C8 = c >> 3, ng = blockIdx.y, nidx = ng / C8, grp = ng - nidx * C8;
  const size_t hw = (size_t)h * w;
  const float inv_hw = 1.f / (float)hw;
  float m[8], r[8], a[8], b[8], cm[8], cs[8], gs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = grp * 8 + k;
    const size_t o = ((size_t)nidx * c + ch) * 2;
    m[k] = __ldg(mr + o);
    r[k] = __ldg(mr + o + 1);
    a[k] = (float)(sums[o] / (double)hw);
    b[k] = (float)(sums[o + 1] / (double)hw);
    gs[k] = gscale ? __ldg(gscale + (size_t)nidx * c + ch) : 1.f;
    cm[k] = cs[k] = 0.f;
    if (dstyle) {
      const float sd = __ldg(style + (size_t)nidx * 2 * c + c + ch);
      cm[k] = __ldg(dstyle + (size_t)nidx * 2 * c + ch) * inv_hw;
      cs[k] = sd > 0.f ? __ldg(dstyle + (size_t)nidx * 2 * c + c + ch) * inv_hw / sd : 0.f;
    }
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  const size_t base = (size_t)ng * hw;
  const unsigned wr = res_pool ? (unsigned)(w >> 1) : (unsigned)w, uw = (unsigned)w;
  const float* g_b = g + base * 8;
  const float* x_b = x + base * 8;
  const float* res_b = res ? res + (size_t)ng * (res_pool ? (hw >> 2) : hw) * 8 : nullptr;
  const float* noise_b = noise ? noise + (size_t)nidx * hw : nullptr;
  float* of_b = out_f32b ? out_f32b + base * 8 : nullptr;
  const size_t act_b = (size_t)ng * planes * hw;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)hw; i += stride) {
    float gv[8], xv[8], v[8];
    load8_f32b(g_b, i, gv);
    load8_f32b(x_b, i, xv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xc = xv[k] - m[k];
      v[k] = r[k] * gs[k] * (gv[k] - a[k] - xc * r[k] * b[k]) + cm[k] + cs[k] * xc;
    }
    if (mode == 0) {
      if (res_b) {
        unsigned ri = i;
        if (res_pool) {
          const unsigned y = i / uw, xx = i - y * uw;
          ri = (y >> 1) * wr + (xx >> 1);
        }
        float rv[8];
        load8_f32b(res_b, ri, rv);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaf(rscale, rv[k], v[k]);
      }
      store8_f32b(of_b, i, v);
    } else {
      const float nz = noise_b ? __ldg(noise_b + i) : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = xv[k] > 0.f ? v[k] : v[k] * slope;
        acc[k] += v[k];
        acc[8 + k] = fmaf(v[k], nz, acc[8 + k]);
      }
      if (out_act) store8_act_at(out_act, act_b + i, hw, planes, v);
      if (of_b) store8_f32b(of_b, i, v);
    }
  }
  if (mode == 1) {
    const float t = block_sums<16>(acc, red);
    if (threadIdx.x < 16) atomicAdd(sums2 + (size_t)(threadIdx.x >> 3) * c + grp * 8 + (threadIdx.x & 7), t);
  }
}

// ---------------------------------------------------------------------------------------------
// FromRGB backward (net.py:231-240: f = lrelu(conv1x1(img) + b)): sums[c][0..2] += sum_pix d_pre * img[i], sums[c][3] += sum d_pre
//   with d_pre = d_f * lrelu'(f); img NCHW [n][cimg <= 3][h][w]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TB_THREADS)
k_from_rgb_bwd(const float* __restrict__ d_f, const float* __restrict__ f, const float* __restrict__ img,
               const float* __restrict__ wgt, float slope, float* __restrict__ sums, float* __restrict__ d_img, int cimg,
               int c, int hw) {
  __shared__ float red[TB_THREADS / 32][32];
  const int C8 = c >> 3, ng = blockIdx.y, nidx = ng / C8, grp = ng - nidx * C8;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  float wk[8][3];   // this group's rows of the 1x1 weight [c][cimg] (image gradient only)
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) wk[k][ci] = (d_img && ci < cimg) ? __ldg(wgt + (size_t)(grp * 8 + k) * cimg + ci) : 0.f;
  const size_t base = (size_t)ng * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)hw; i += (size_t)gridDim.x * blockDim.x) {
    float dv[8], fv[8], px[3] = {0.f, 0.f, 0.f}, gi[3] = {0.f, 0.f, 0.f};
    load8_f32b(d_f, base + i, dv);
    load8_f32b(f, base + i, fv);
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
      if (ci < cimg) px[ci] = __ldg(img + ((size_t)nidx * cimg + ci) * hw + i);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = fv[k] > 0.f ? dv[k] : dv[k] * slope;
      acc[4 * k] = fmaf(d, px[0], acc[4 * k]);
      acc[4 * k + 1] = fmaf(d, px[1], acc[4 * k + 1]);
      acc[4 * k + 2] = fmaf(d, px[2], acc[4 * k + 2]);
      acc[4 * k + 3] += d;
      gi[0] = fmaf(d, wk[k][0], gi[0]);
      gi[1] = fmaf(d, wk[k][1], gi[1]);
      gi[2] = fmaf(d, wk[k][2], gi[2]);
    }
    if (d_img) {   // d img[n][ci] = sum_c W[c][ci] * d_pre[c]: the channel groups of one pixel meet in an atomic add
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        if (ci < cimg) atomicAdd(d_img + ((size_t)nidx * cimg + ci) * hw + i, gi[ci]);
    }
  }
  const float t = block_sums<32>(acc, red);
  if (threadIdx.x < 32) atomicAdd(sums + (size_t)(grp * 8 + (threadIdx.x >> 2)) * 4 + (threadIdx.x & 3), t);
}

// ---------------------------------------------------------------------------------------------
// StyleGAN2 synthesis layer, backward of everything after the contraction (stylegan2_generator.py:907-921, 515-522):
//   y = lrelu(conv * dm + noise * ns + b) * gain ;  next layer reads y * s_next ;  ToRGB reads sum_c rgbw[ch][c] * y[c]
// The forward kept y only as the NEXT layer's operand ya = y * s_next (ACT), so y = ya / s_next.
//   dy     = dxs * s_next + sum_ch rgbw[n][ch][c] * dimg[n][ch]
//   d_pre  = dy * gain * lrelu'(y) ;  d_conv = d_pre * dm   -> ACT (plain conv) or F32B (x2 layer: input of the FIR transpose)
//   sums[n][c] = ( S = sum dxs * y            -> d s_next[n][c]
//                  T0..T2 = sum dimg[ch] * y  -> d rgbw[n][ch][c]
//                  D = sum d_pre * (pre - noise*ns - b), pre = pre-activation (recovered from y)  -> d dm[n][c] * dm[n][c] )
// ---------------------------------------------------------------------------------------------
struct Sg2BwdParams {
  const void* ya;
  const float* ya_scale;
  const float* dxs;
  const float* dimg;
  const float* rgbw;
  const float* noise;
  long long noise_bstride;
  float noise_scalar;
  const float* bias;
  const float* demod;
  float gain, slope;
  void* out_act;
  float* out_f32b;
  float* sums;
  int c, hw, planes, out_planes;
};

// Thread = (pixel, channel HALF): 4 channels per thread keep the 5 x 4 accumulators and the per-channel constants in ~70
// registers (three blocks per SM instead of one with 8 channels per thread: the kernel is latency-bound on its loads);
// the two halves of a 16-byte ACT chunk / 32-byte F32B sector are read by neighbouring lanes, so accesses stay coalesced.
// RGB: the layer feeds a ToRGB (d_image term + the three T sums); without it 12 accumulators and 12 constants fewer.
// DXS: a next layer exists (its data gradient dxs and the S sum); the last layer of the network has none.
template <bool RGB, bool DXS>
__global__ void __launch_bounds__(TB_THREADS, RGB ? 3 : 4)
k_sg2_layer_bwd(const __grid_constant__ Sg2BwdParams p) {
  __shared__ float red[TB_THREADS / 32][2][20];
  const int c = p.c, C8 = c >> 3, ng = blockIdx.y, nidx = ng / C8, grp = ng - nidx * C8;
  const size_t hw = (size_t)p.hw;
  const int half = threadIdx.x & 1;
  // the 3 x 8 ToRGB weights of this (sample, channel group) live in shared memory (12 registers the loop cannot spare)
  __shared__ __align__(16) float srw[2][3][4];
  if (RGB && threadIdx.x < 24) {
    const int hf = threadIdx.x / 12, q = (threadIdx.x % 12) >> 2, k = threadIdx.x & 3;
    srw[hf][q][k] = __ldg(p.rgbw + ((size_t)nidx * 3 + q) * c + grp * 8 + 4 * hf + k);
  }
  if (RGB) __syncthreads();
  float sn[4], isn[4], dm[4], bs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch = grp * 8 + 4 * half + k;
    sn[k] = p.ya_scale ? __ldg(p.ya_scale + (size_t)nidx * c + ch) : 1.f;
    isn[k] = sn[k] != 0.f ? 1.f / sn[k] : 0.f;
    dm[k] = p.demod ? __ldg(p.demod + (size_t)nidx * c + ch) : 1.f;
    bs[k] = p.bias ? __ldg(p.bias + ch) : 0.f;
  }
  const float g_pos = p.gain, g_neg = p.gain * p.slope;
  const float ig_pos = 1.f / p.gain, ig_neg = p.slope != 0.f ? 1.f / (p.gain * p.slope) : 0.f;
  float acc[20];
#pragma unroll
  for (int i = 0; i < 20; ++i) acc[i] = 0.f;
  // every base pointer of the loop is formed once (the 64-bit index products were ~20 % of the instructions issued)
  const uint2* ya = reinterpret_cast<const uint2*>(p.ya) + ((size_t)ng * p.planes * hw) * 2 + half;
  const float4* dxs = DXS ? reinterpret_cast<const float4*>(p.dxs) + ((size_t)ng * hw) * 2 + half : nullptr;
  const float* di0 = RGB ? p.dimg + (size_t)nidx * 3 * hw : nullptr;
  const float* nzp = p.noise ? p.noise + (size_t)nidx * p.noise_bstride : nullptr;
  uint2* oact = p.out_act ? reinterpret_cast<uint2*>(p.out_act) + ((size_t)ng * p.out_planes * hw) * 2 + half : nullptr;
  const unsigned uhw = (unsigned)hw;
  float4* of32 = p.out_f32b ? reinterpret_cast<float4*>(p.out_f32b) + ((size_t)ng * hw) * 2 + half : nullptr;
  const bool two_in = p.planes == 2, two_out = p.out_planes == 2;
  const float nsc = p.noise_scalar;
  const unsigned stride = (gridDim.x * blockDim.x) >> 1;
  for (unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 1; i < (unsigned)hw; i += stride) {
    float y[4], dx[4], v[4];
    {
      const uint2 q = __ldg(ya + 2 * i);
      y[0] = __uint_as_float(q.x << 16); y[1] = __uint_as_float(q.x & 0xffff0000u);
      y[2] = __uint_as_float(q.y << 16); y[3] = __uint_as_float(q.y & 0xffff0000u);
      if (two_in) {
        const uint2 l = __ldg(ya + 2 * (size_t)(i + uhw));
        y[0] += __uint_as_float(l.x << 16); y[1] += __uint_as_float(l.x & 0xffff0000u);
        y[2] += __uint_as_float(l.y << 16); y[3] += __uint_as_float(l.y & 0xffff0000u);
      }
    }
    if (DXS) {
      const float4 t = __ldg(dxs + 2 * i);
      dx[0] = t.x; dx[1] = t.y; dx[2] = t.z; dx[3] = t.w;
    } else {
      dx[0] = dx[1] = dx[2] = dx[3] = 0.f;
    }
    float di[3] = {0.f, 0.f, 0.f};
    if (RGB) {
      di[0] = __ldg(di0 + i);
      di[1] = __ldg(di0 + (i + uhw));
      di[2] = __ldg(di0 + (size_t)i + 2 * (size_t)uhw);
    }
    const float nzs = nzp ? __ldg(nzp + i) * nsc : 0.f;
    float rw[3][4];
    if (RGB) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(srw[half][q]);
        rw[q][0] = t.x; rw[q][1] = t.y; rw[q][2] = t.z; rw[q][3] = t.w;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float yv = y[k] * isn[k];
      float dy = DXS ? dx[k] * sn[k] : 0.f;
      if (DXS) acc[k] = fmaf(dx[k], yv, acc[k]);
      if (RGB) {
        dy = fmaf(rw[0][k], di[0], fmaf(rw[1][k], di[1], fmaf(rw[2][k], di[2], dy)));
        acc[4 + k] = fmaf(di[0], yv, acc[4 + k]);
        acc[8 + k] = fmaf(di[1], yv, acc[8 + k]);
        acc[12 + k] = fmaf(di[2], yv, acc[12 + k]);
      }
      const bool pos = yv > 0.f;
      const float dpre = dy * (pos ? g_pos : g_neg);
      const float pre = yv * (pos ? ig_pos : ig_neg);
      acc[16 + k] = fmaf(dpre, pre - nzs - bs[k], acc[16 + k]);
      v[k] = dpre * dm[k];
    }
    if (oact) {
      uint32_t hw2[2], lw2[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        hw2[j] = *reinterpret_cast<const uint32_t*>(&hb);
        const __nv_bfloat162 lb = __floats2bfloat162_rn(v[2 * j] - __uint_as_float(hw2[j] << 16),
                                                        v[2 * j + 1] - __uint_as_float(hw2[j] & 0xffff0000u));
        lw2[j] = *reinterpret_cast<const uint32_t*>(&lb);
      }
      oact[2 * i] = make_uint2(hw2[0], hw2[1]);
      if (two_out) oact[2 * (size_t)(i + uhw)] = make_uint2(lw2[0], lw2[1]);
    }
    if (of32) of32[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
  }
  // reduce over the lanes of the same half (xor 16, 8, 4, 2), then over the warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 20; ++i) {
#pragma unroll
    for (int off = 16; off >= 2; off >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
  }
  if (lane < 2) {
#pragma unroll
    for (int i = 0; i < 20; ++i) red[warp][lane][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < 40) {
    const int hf = threadIdx.x / 20, j = threadIdx.x - hf * 20;      // j = which * 4 + k
    float t = 0.f;
    for (int w = 0; w < TB_THREADS / 32; ++w) t += red[w][hf][j];
    atomicAdd(p.sums + ((size_t)nidx * c + grp * 8 + 4 * hf + (j & 3)) * 5 + (j >> 2), t);
  }
}

// ---------------------------------------------------------------------------------------------
// Transpose of the x2 layer's FIR (stylegan2_generator.py:603-615, pad (1,1,1,1), f = [1,3,3,1]/4 per axis), written
// as the SPACE-TO-DEPTH operand of the stride-2 data-gradient conv:
//   dt[u][v] = sum_{a,b<4} f[a] f[b] dconv[u-a+1][v-b+1]          u in [0, 2H], v in [0, 2W]   (dconv = 0 outside)
//   out ACT [n][4*C/8][planes][H+1][W+1][8], channel block (2*py+px)*C/8 + g holds dt[2Y+py][2X+px] (0 beyond the map)
// DGE_CONV_DOWN4X4S2 with a zero first kernel row / column then is the 3x3 stride-2 conv that inverts the transposed conv.
// Thread = one (Y, X, channel group): a 5x5 window of dconv -> the four phases.
// ---------------------------------------------------------------------------------------------
// box != 0: the 2x2 box sum of the StyleGAN1 `transform_kernel` layers instead (dge_sg1_post mode 1, lreq.py:127-131):
//   dt[u][v] = sum_{a,b<2} dconv[u-a][v-b].
// weights of window rows / columns 2Y-2 .. 2Y+2 for the even (0) and odd (1) phase: compile-time constants, so the taps
// with a zero weight cost nothing (BOX: 2 of 5 per phase, FIR: 4 of 5)
template <bool BOX>
__device__ __forceinline__ constexpr float s2d_w(int q, int j) {
  return BOX ? ((j == 1 + q || j == 2 + q) ? 1.f : 0.f)
             : ((j == 0 + q || j == 3 + q) ? 0.25f : ((j == 1 + q || j == 2 + q) ? 0.75f : 0.f));
}

template <bool BOX>
__global__ void __launch_bounds__(TB_THREADS)
k_up_fir_bwd_s2d(const float* __restrict__ dconv, void* __restrict__ out, int n, int c, int H, int W, int planes) {
  const int C8 = c >> 3, Hs = H + 1, Ws = W + 1, Ho = 2 * H, Wo = 2 * W;
  const size_t total = (size_t)n * C8 * Hs * Ws;
  // (32-bit index arithmetic: the host checks total < 2^31; three 64-bit divisions per thread cost more than the filter)
  const unsigned uWs = (unsigned)Ws, uHs = (unsigned)Hs, uC8 = (unsigned)C8, stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)total; i += stride) {
    unsigned t = i / uWs;
    const int X = (int)(i - t * uWs);
    const unsigned t2 = t / uHs;
    const int Y = (int)(t - t2 * uHs);
    const unsigned nq = t2 / uC8;
    const int g = (int)(t2 - nq * uC8), nidx = (int)nq;
    float o[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < 8; ++k) o[q][k] = 0.f;
    const float* src = dconv + ((size_t)nidx * C8 + g) * Ho * (size_t)Wo * 8;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int r = 2 * Y - 2 + j;
      if (r < 0 || r >= Ho) continue;
      float h0[8], h1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) h0[k] = h1[k] = 0.f;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const int cc = 2 * X - 2 + b;
        if (cc < 0 || cc >= Wo) continue;
        float v[8];
        load8_f32b(src, (unsigned)(r * Wo + cc), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (s2d_w<BOX>(0, b) != 0.f) h0[k] = fmaf(s2d_w<BOX>(0, b), v[k], h0[k]);
          if (s2d_w<BOX>(1, b) != 0.f) h1[k] = fmaf(s2d_w<BOX>(1, b), v[k], h1[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (s2d_w<BOX>(0, j) != 0.f) {
          o[0][k] = fmaf(s2d_w<BOX>(0, j), h0[k], o[0][k]);
          o[1][k] = fmaf(s2d_w<BOX>(0, j), h1[k], o[1][k]);
        }
        if (s2d_w<BOX>(1, j) != 0.f) {
          o[2][k] = fmaf(s2d_w<BOX>(1, j), h0[k], o[2][k]);
          o[3][k] = fmaf(s2d_w<BOX>(1, j), h1[k], o[3][k]);
        }
      }
    }
    const size_t hws = (size_t)Hs * Ws, pix = (size_t)Y * Ws + X;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // phases beyond the (2H+1) x (2W+1) map are padding: finite zeros (their weights are zero, 0 * NaN is not)
      const bool inside = (2 * Y + (q >> 1) <= Ho) && (2 * X + (q & 1) <= Wo);
      if (!inside) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[q][k] = 0.f;
      }
      store8_act_at(out, (((size_t)nidx * 4 + q) * C8 + g) * planes * hws + pix, hws, planes, o[q]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Transpose of the skip branch's x2 up-sampling (dge_rgb_init / stylegan2_generator.py:519-522, 603-615):
//   forward per axis: out[2m] = (x[m-1] + 3 x[m]) / 4, out[2m+1] = (3 x[m] + x[m+1]) / 4
//   d_in[m] = (d[2m-1] + 3 d[2m] + 3 d[2m+1] + d[2m+2]) / 4 per axis (d = 0 outside);  planes = n * channels
// ---------------------------------------------------------------------------------------------
__global__ void k_rgb_up_bwd(const float* __restrict__ d_out, float* __restrict__ d_in, size_t planes, int hin, int win) {
  const size_t total = planes * hin * win;
  const int ho = 2 * hin, wo = 2 * win;
  const float f[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int mx = (int)(i % win);
    size_t t = i / win;
    const int my = (int)(t % hin);
    t /= hin;
    const float* src = d_out + t * (size_t)ho * wo;
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int y = 2 * my - 1 + a;
      if (y < 0 || y >= ho) continue;
      float r = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int x = 2 * mx - 1 + b;
        if (x >= 0 && x < wo) r = fmaf(f[b], __ldg(src + (size_t)y * wo + x), r);
      }
      s = fmaf(f[a], r, s);
    }
    d_in[i] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// LPIPS-VGG16 (lpips v0.1 net='vgg'; training_utils.py:93 consumes it): the pieces between the VGG convolutions.
// The convs themselves are dge_conv_forward with bias + ReLU (slope 0) in the epilogue.
// ---------------------------------------------------------------------------------------------
// ScalingLayer + channel padding: NCHW [n][3][h][w] -> ACT [n][16/8][planes][h][w][8] = (x - shift) / scale, channels 3..15 = 0
__global__ void k_lpips_input(const float* __restrict__ x, void* __restrict__ out, int n, int h, int w, int planes,
                              float s0, float s1, float s2, float i0, float i1, float i2) {
  const size_t hw = (size_t)h * w, total = (size_t)n * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / hw, pix = i - b * hw;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    v[0] = (__ldg(x + (b * 3 + 0) * hw + pix) - s0) * i0;
    v[1] = (__ldg(x + (b * 3 + 1) * hw + pix) - s1) * i1;
    v[2] = (__ldg(x + (b * 3 + 2) * hw + pix) - s2) * i2;
    store8_act_at(out, (b * 2 + 0) * planes * hw + pix, hw, planes, v);
    store8_act_at(out, (b * 2 + 1) * planes * hw + pix, hw, planes, z);
  }
}

// nn.MaxPool2d(2, 2) (floor mode) F32B [n][c/8][h][w][8] -> ACT [n][c/8][planes][h/2][w/2][8]
__global__ void k_maxpool_to_act(const float* __restrict__ x, void* __restrict__ out, int n, int c, int h, int w,
                                 int planes) {
  const int C8 = c >> 3, ho = h >> 1, wo = w >> 1;
  const size_t total = (size_t)n * C8 * ho * wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int px = (int)(i % wo);
    size_t t = i / wo;
    const int py = (int)(t % ho);
    const size_t ng = t / ho;
    float m[8];
    load8_f32b(x, (ng * h + 2 * py) * w + 2 * px, m);
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      float v[8];
      load8_f32b(x, (ng * h + 2 * py + (q >> 1)) * w + 2 * px + (q & 1), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
    }
    store8_act_at(out, (ng * planes * ho + py) * (size_t)wo + px, (size_t)ho * wo, planes, m);
  }
}

// Backward of [ReLU -> (tap) -> MaxPool2d(2,2)]: the gradient of a conv's pre-activation as the ACT operand of its
// data-gradient conv.   d_pre = (y > 0) * ( g_same + (this pixel is the arg-max of its 2x2 window ? g_pool[y/2][x/2] : 0) )
//   y: the activated conv output, F32B (y_act == NULL) or ACT (only its hi plane is read: the sign);  g_same: F32B at the
//   same resolution (NULL: none);  g_pool: F32B at (h/2, w/2), routed to the FIRST maximum of each window in row-major order
//   (torch's max_pool2d tie rule; ties at 0 are masked by the ReLU anyway) (NULL: none).  Thread = one 2x2 window.
__global__ void k_relu_pool_bwd(const float* __restrict__ y_f32b, const void* __restrict__ y_act, int y_planes,
                                const float* __restrict__ g_same, const float* __restrict__ g_pool,
                                void* __restrict__ out, int n, int c, int h, int w, int planes, int clamp_pos) {
  const int C8 = c >> 3, hc = (h + 1) >> 1, wc = (w + 1) >> 1, ho = h >> 1, wo = w >> 1;
  const size_t total = (size_t)n * C8 * hc * wc, hw = (size_t)h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int px = (int)(i % wc);
    size_t t = i / wc;
    const int py = (int)(t % hc);
    const size_t ng = t / hc;
    float yv[4][8];
    bool in[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = 2 * py + (q >> 1), xx = 2 * px + (q & 1);
      in[q] = yy < h && xx < w;
      if (in[q]) {
        if (y_act) {
          const uint4* yp = reinterpret_cast<const uint4*>(y_act) + ng * y_planes * hw + (size_t)yy * w + xx;
          unpack8(__ldg(yp), yv[q]);
          if (y_planes == 2 && g_pool) {   // the arg-max needs hi + lo; the ReLU mask alone is decided by the hi plane
            float l[8];
            unpack8(__ldg(yp + hw), l);
#pragma unroll
            for (int k = 0; k < 8; ++k) yv[q][k] += l[k];
          }
        } else {
          load8_f32b(y_f32b, ng * hw + (size_t)yy * w + xx, yv[q]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) yv[q][k] = -1.f;
      }
    }
    float gp[8];
    const bool pooled = g_pool && py < ho && px < wo;
    if (pooled) load8_f32b(g_pool, (ng * ho + py) * (size_t)wo + px, gp);
    int am[8];
    if (pooled) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int a = 0;
        float mv = yv[0][k];
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (yv[q][k] > mv) { mv = yv[q][k]; a = q; }
        am[k] = a;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!in[q]) continue;
      const size_t pix = (size_t)(2 * py + (q >> 1)) * w + 2 * px + (q & 1);
      float g[8];
      if (g_same) {
        load8_f32b(g_same, ng * hw + pix, g);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (pooled && am[k] == q) g[k] += gp[k];
        g[k] = yv[q][k] > 0.f ? g[k] : 0.f;
        if (clamp_pos) g[k] = fmaxf(g[k], 0.f);      // guided back-propagation (grad_cam.py:209-211)
      }
      store8_act_at(out, ng * planes * hw + pix, hw, planes, g);
    }
  }
}

// LPIPS distance of one tap (lpips: normalize_tensor, squared difference, 1x1 `lin`, spatial mean):
//   out[n] += (1/HW) sum_p sum_c w_c (a_c / (|a|+eps) - b_c / (|b|+eps))^2,   a = f[n], b = f[n + nb]   (f: F32B, 2*nb samples)
// Thread = one pixel: four channel sums give everything (also for the backward):
//   Saa = sum a^2, Sbb = sum b^2, Waa = sum w a^2, Wbb = sum w b^2, Wab = sum w a b ;  d = Waa/na^2 + Wbb/nb^2 - 2 Wab/(na nb)
// grad_mode: 0 = forward; 1 = gradient w.r.t. b -> gb (F32B [nb]); 2 = w.r.t. a -> ga; 3 = both.  go[n] = upstream gradient.
// Eight lanes share a pixel, each taking every 8th channel group (the deep taps have 512 channels on 16 x 16 pixels: one
// thread per pixel would walk 64 groups serially with a few thousand threads in flight); the five sums meet in a 3-step
// shuffle.  A warp reads 4 pixels x 8 groups: 128-byte runs.
constexpr int LP_CS = 8;
__global__ void __launch_bounds__(256)
k_lpips_dist(const float* __restrict__ f, const float* __restrict__ lw, float* __restrict__ out,
             const float* __restrict__ go, float* __restrict__ ga, float* __restrict__ gb, int grad_mode, int nb, int c,
             int hw, float eps) {
  __shared__ float red[8];
  const int C8 = c >> 3, n = blockIdx.y;
  const int sub = threadIdx.x & (LP_CS - 1);
  const float inv_hw = 1.f / (float)hw;
  float part = 0.f;
  const size_t stride = ((size_t)gridDim.x * blockDim.x) / LP_CS;
  // the loop bound is WARP-uniform (the shuffles below name all 32 lanes); pixels beyond the map contribute zeros
  const size_t warp_first = (blockIdx.x * (size_t)blockDim.x + (threadIdx.x & ~31u)) / LP_CS;
  for (size_t i0 = warp_first; i0 < (size_t)hw; i0 += stride) {
    const size_t i = i0 + ((threadIdx.x & 31) / LP_CS);
    const bool valid = i < (size_t)hw;
    float saa = 0.f, sbb = 0.f, waa = 0.f, wbb = 0.f, wab = 0.f;
#pragma unroll 2
    for (int g = sub; valid && g < C8; g += LP_CS) {
      float a[8], b[8];
      load8_f32b(f, ((size_t)n * C8 + g) * hw + i, a);
      load8_f32b(f, ((size_t)(n + nb) * C8 + g) * hw + i, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float w = __ldg(lw + g * 8 + k);
        saa = fmaf(a[k], a[k], saa);
        sbb = fmaf(b[k], b[k], sbb);
        waa = fmaf(w * a[k], a[k], waa);
        wbb = fmaf(w * b[k], b[k], wbb);
        wab = fmaf(w * a[k], b[k], wab);
      }
    }
#pragma unroll
    for (int off = LP_CS / 2; off; off >>= 1) {
      saa += __shfl_xor_sync(0xffffffffu, saa, off);
      sbb += __shfl_xor_sync(0xffffffffu, sbb, off);
      waa += __shfl_xor_sync(0xffffffffu, waa, off);
      wbb += __shfl_xor_sync(0xffffffffu, wbb, off);
      wab += __shfl_xor_sync(0xffffffffu, wab, off);
    }
    const float ra = sqrtf(saa), rb = sqrtf(sbb), na = ra + eps, nbn = rb + eps;
    if (grad_mode == 0) {
      if (sub == 0 && valid) part += waa / (na * na) + wbb / (nbn * nbn) - 2.f * wab / (na * nbn);
      continue;
    }
    if (!valid) continue;
    // u = a/na, v = b/nbn;  dD/dv_c = -2 w_c (u_c - v_c);  dv_c/db_k = delta_ck/nbn - b_c b_k/(nbn^2 rb)
    //   dD/db_k = (1/nbn) [ gv_k - b_k/(nbn rb) * sum_c gv_c b_c ],  sum_c gv_c b_c = -2 (wab/na - wbb/nbn)   (a: symmetric)
    const float scale = __ldg(go + n) * inv_hw;
    const float sgb = -2.f * (wab / na - wbb / nbn), sga = -2.f * (wab / nbn - waa / na);
    const float cb = rb > 0.f ? sgb / (nbn * rb) : 0.f, ca = ra > 0.f ? sga / (na * ra) : 0.f;
    for (int g = sub; g < C8; g += LP_CS) {
      float a[8], b[8], da[8], db[8];
      load8_f32b(f, ((size_t)n * C8 + g) * hw + i, a);
      load8_f32b(f, ((size_t)(n + nb) * C8 + g) * hw + i, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float w = __ldg(lw + g * 8 + k);
        const float diff = a[k] / na - b[k] / nbn;
        db[k] = scale * (-2.f * w * diff - b[k] * cb) / nbn;
        da[k] = scale * (2.f * w * diff - a[k] * ca) / na;
      }
      if (grad_mode & 1) store8_f32b(gb, ((size_t)n * C8 + g) * hw + i, db);
      if (grad_mode & 2) store8_f32b(ga, ((size_t)n * C8 + g) * hw + i, da);
    }
  }
  if (grad_mode == 0) {
#pragma unroll
    for (int off = 16; off; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
      atomicAdd(out + n, t * inv_hw);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// conditional batch norm + (leaky) ReLU, backward -- the frozen BigGAN generator under loss.backward()
// (biggan_generator.py:138-150 + 178-190: t = relu(a[n][c] * x + b[n][c]) [-> nearest x2] -> conv).
//   g    F32B [n][c/8][h*up][w*up][8]  gradient of the conv's input (what its data-gradient conv wrote)
//   x    F32B [n][c/8][h][w][8]        the block-norm's input kept by the forward
//   d    = (a*x + b > 0 ? 1 : slope) * sum_{up x up} g
//   sums [n][c][2] += (sum d * x, sum d)                 -> gradients of a and b (the condition vector's path)
//   dx   = a * d (+ for channels < skip_c: sum_{sup x sup} skip, the channel-drop / nearest-up identity branch :192-203)
//        -> out_f32b and / or out_act (the operand of the next data-gradient conv)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TB_THREADS, 3)
k_affine_relu_bwd(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ a,
                  const float* __restrict__ b, float slope, int up, const float* __restrict__ skip, int skip_c, int sup,
                  float* __restrict__ out_f32b, void* __restrict__ out_act, float* __restrict__ sums, int c, int h, int w,
                  int planes) {
  __shared__ float red[TB_THREADS / 32][16];
  const int C8 = c >> 3, ng = blockIdx.y, nidx = ng / C8, grp = ng - nidx * C8;
  const size_t hw = (size_t)h * w;
  float av[8], bv[8], acc[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    av[k] = __ldg(a + (size_t)nidx * c + grp * 8 + k);
    bv[k] = __ldg(b + (size_t)nidx * c + grp * 8 + k);
    acc[k] = acc[8 + k] = 0.f;
  }
  const size_t base = (size_t)ng * hw, gbase = base * (size_t)(up * up);
  const int gw = w * up;
  const bool has_skip = skip && grp * 8 < skip_c;
  const size_t sbase = ((size_t)nidx * (skip_c >> 3) + grp) * hw * (size_t)(sup * sup);
  const int sw = w * sup;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), xx = (int)(i - (size_t)y * w);
    float xv[8], gs[8], v[8];
    load8_f32b(x, base + i, xv);
    if (up == 1) {
      load8_f32b(g, gbase + i, gs);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) gs[k] = 0.f;
      for (int dy = 0; dy < up; ++dy)
        for (int dx = 0; dx < up; ++dx) {
          float t[8];
          load8_f32b(g, gbase + (size_t)(y * up + dy) * gw + (xx * up + dx), t);
#pragma unroll
          for (int k = 0; k < 8; ++k) gs[k] += t[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = fmaf(av[k], xv[k], bv[k]) > 0.f ? gs[k] : gs[k] * slope;
      acc[2 * k] = fmaf(d, xv[k], acc[2 * k]);
      acc[2 * k + 1] += d;
      v[k] = av[k] * d;
    }
    if (has_skip) {
      for (int dy = 0; dy < sup; ++dy)
        for (int dx = 0; dx < sup; ++dx) {
          float t[8];
          load8_f32b(skip, sbase + (size_t)(y * sup + dy) * sw + (xx * sup + dx), t);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] += t[k];
        }
    }
    if (out_f32b) store8_f32b(out_f32b, base + i, v);
    if (out_act) store8_act_at(out_act, (size_t)ng * planes * hw + i, hw, planes, v);
  }
  const float t = block_sums<16>(acc, red);
  if (threadIdx.x < 16) atomicAdd(sums + ((size_t)nidx * c + grp * 8) * 2 + threadIdx.x, t);
}

}  // namespace dge

using namespace dge;

#define TB_STREAM ((cudaStream_t)stream)
#define TB_ZERO(ptr, bytes)                                                           \
  do {                                                                                \
    cudaError_t me_ = cudaMemsetAsync((ptr), 0, (bytes), TB_STREAM);                  \
    if (me_ != cudaSuccess) {                                                         \
      set_error("train_bwd: memset failed: %s", cudaGetErrorString(me_));             \
      return DGE_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

extern "C" int dge_be_head_bwd(const float* d_out, const float* y2, const float* noise, float ga, float gb, float slope,
                               void* dy2_act, void* dres_act, float* sums, int n, int co, int h, int w, int planes,
                               void* stream) {
  DGE_REQUIRE(d_out && y2 && dy2_act && sums, "be_head_bwd: null pointer");
  DGE_REQUIRE(n > 0 && co >= 16 && co % 16 == 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0,
              "be_head_bwd: bad dims n=%d co=%d h=%d w=%d", n, co, h, w);
  DGE_REQUIRE(planes == 1 || planes == 2, "be_head_bwd: planes=%d", planes);
  TB_ZERO(sums, (size_t)3 * co * sizeof(float));
  const long long hwp = (long long)(h / 2) * (w / 2);
  dim3 grid(tb_splits(hwp, (long long)n * (co / 8)), n * (co / 8));
  k_be_head_bwd<<<grid, TB_THREADS, 0, TB_STREAM>>>(d_out, y2, noise, ga * 0.25f, gb, slope, dy2_act, dres_act, sums, co,
                                                     h / 2, w / 2, planes);
  count_launch();
  return check_launch("k_be_head_bwd");
}

extern "C" int dge_in_bwd_stats(const float* g, const float* x, const float* mean_rstd, double* sums, int n, int c,
                                int h, int w, void* stream) {
  DGE_REQUIRE(g && x && mean_rstd && sums, "in_bwd_stats: null pointer");
  DGE_REQUIRE(n > 0 && c >= 8 && c % 8 == 0 && h > 0 && w > 0, "in_bwd_stats: bad dims n=%d c=%d h=%d w=%d", n, c, h, w);
  TB_ZERO(sums, (size_t)2 * n * c * sizeof(double));
  dim3 grid(tb_splits((long long)h * w, (long long)n * (c / 8)), n * (c / 8));
  k_in_bwd_stats<<<grid, TB_THREADS, 0, TB_STREAM>>>(g, x, mean_rstd, sums, c, h * w);
  count_launch();
  return check_launch("k_in_bwd_stats");
}

extern "C" int dge_in_bwd_apply(const float* g, const float* x, const float* mean_rstd, const float* style,
                                const float* dstyle, const double* sums, const float* gscale, int mode, const float* res,
                                float rscale, int res_pool, const float* noise, float slope, float* out_f32b,
                                void* out_act, float* sums2, int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(g && x && mean_rstd && sums, "in_bwd_apply: null pointer");
  DGE_REQUIRE(mode == 0 || mode == 1, "in_bwd_apply: mode=%d", mode);
  DGE_REQUIRE(n > 0 && c >= 8 && c % 8 == 0 && h > 0 && w > 0, "in_bwd_apply: bad dims n=%d c=%d h=%d w=%d", n, c, h, w);
  DGE_REQUIRE(!dstyle || style, "in_bwd_apply: dstyle needs style (mean || std)");
  if (mode == 0) {
    DGE_REQUIRE(out_f32b, "in_bwd_apply: mode 0 writes out_f32b");
    DGE_REQUIRE(!res || !res_pool || (h % 2 == 0 && w % 2 == 0), "in_bwd_apply: pooled residual needs even h, w");
  } else {
    DGE_REQUIRE((out_act || out_f32b) && sums2 && (!out_act || ((planes == 1 || planes == 2) && c % 16 == 0)),
                "in_bwd_apply: mode 1 writes out_act (planes 1|2, c %% 16 == 0) and / or out_f32b, and sums2");
    TB_ZERO(sums2, (size_t)2 * c * sizeof(float));
  }
  dim3 grid(tb_splits((long long)h * w, (long long)n * (c / 8)), n * (c / 8));
  k_in_bwd_apply<<<grid, TB_THREADS, 0, TB_STREAM>>>(g, x, mean_rstd, style, dstyle, sums, gscale, mode, res, rscale,
                                                      res_pool, noise, slope, out_f32b, out_act, sums2, c, h, w, planes);
  count_launch();
  return check_launch("k_in_bwd_apply");
}

extern "C" int dge_affine_relu_bwd(const float* g, const float* x, const float* a, const float* b, float slope, int up,
                                   const float* skip, int skip_c, int skip_up, float* out_f32b, void* out_act,
                                   float* sums, int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(g && x && a && b && sums && (out_f32b || out_act), "affine_relu_bwd: null pointer");
  DGE_REQUIRE(n > 0 && c >= 8 && c % 8 == 0 && h > 0 && w > 0, "affine_relu_bwd: bad dims n=%d c=%d h=%d w=%d", n, c, h, w);
  DGE_REQUIRE(up == 1 || up == 2, "affine_relu_bwd: up=%d", up);
  DGE_REQUIRE(!skip || (skip_c >= 8 && skip_c % 8 == 0 && skip_c <= c && (skip_up == 1 || skip_up == 2)),
              "affine_relu_bwd: skip needs 8 <= skip_c <= c, a multiple of 8, and skip_up 1|2 (skip_c=%d skip_up=%d)",
              skip_c, skip_up);
  DGE_REQUIRE(!out_act || ((planes == 1 || planes == 2) && c % 16 == 0), "affine_relu_bwd: out_act needs planes 1|2, c %% 16 == 0");
  TB_ZERO(sums, (size_t)2 * n * c * sizeof(float));
  dim3 grid(tb_splits((long long)h * w, (long long)n * (c / 8)), n * (c / 8));
  k_affine_relu_bwd<<<grid, TB_THREADS, 0, TB_STREAM>>>(g, x, a, b, slope, up, skip, skip_c, skip ? skip_up : 1, out_f32b,
                                                        out_act, sums, c, h, w, planes);
  count_launch();
  return check_launch("k_affine_relu_bwd");
}

extern "C" int dge_from_rgb_bwd(const float* d_f, const float* f, const float* img, const float* wgt, float slope,
                                float* sums, float* d_img, int n, int cimg, int c, int h, int w, void* stream) {
  DGE_REQUIRE(d_f && f && img && sums, "from_rgb_bwd: null pointer");
  DGE_REQUIRE(!d_img || wgt, "from_rgb_bwd: d_img needs the weight");
  if (d_img) TB_ZERO(d_img, (size_t)n * cimg * h * w * sizeof(float));
  DGE_REQUIRE(n > 0 && cimg >= 1 && cimg <= 3 && c >= 8 && c % 8 == 0 && h > 0 && w > 0,
              "from_rgb_bwd: bad dims n=%d cimg=%d c=%d h=%d w=%d", n, cimg, c, h, w);
  TB_ZERO(sums, (size_t)4 * c * sizeof(float));
  dim3 grid(tb_splits((long long)h * w, (long long)n * (c / 8)), n * (c / 8));
  k_from_rgb_bwd<<<grid, TB_THREADS, 0, TB_STREAM>>>(d_f, f, img, wgt, slope, sums, d_img, cimg, c, h * w);
  count_launch();
  return check_launch("k_from_rgb_bwd");
}

extern "C" int dge_sg2_layer_bwd(const void* ya_act, const float* ya_scale, const float* dxs, const float* dimg,
                                 const float* rgbw, const float* noise, int64_t noise_bstride, float noise_scalar,
                                 const float* bias, const float* demod, float gain, float slope, void* out_act,
                                 int out_planes, float* out_f32b, float* sums, int n, int c, int h, int w, int planes,
                                 void* stream) {
  DGE_REQUIRE(ya_act && sums && (out_act || out_f32b), "sg2_layer_bwd: null pointer");
  DGE_REQUIRE(n > 0 && c >= 16 && c % 16 == 0 && h > 0 && w > 0, "sg2_layer_bwd: bad dims n=%d c=%d h=%d w=%d", n, c, h, w);
  DGE_REQUIRE((planes == 1 || planes == 2) && (!out_act || out_planes == 1 || out_planes == 2),
              "sg2_layer_bwd: planes=%d out_planes=%d", planes, out_planes);
  DGE_REQUIRE(!dimg == !rgbw, "sg2_layer_bwd: dimg and rgbw go together");
  DGE_REQUIRE(gain != 0.f, "sg2_layer_bwd: gain must be non-zero");
  TB_ZERO(sums, (size_t)5 * n * c * sizeof(float));
  Sg2BwdParams p;
  p.ya = ya_act; p.ya_scale = ya_scale; p.dxs = dxs; p.dimg = dimg; p.rgbw = rgbw; p.noise = noise;
  p.noise_bstride = noise_bstride; p.noise_scalar = noise_scalar; p.bias = bias; p.demod = demod; p.gain = gain;
  p.slope = slope; p.out_act = out_act; p.out_f32b = out_f32b; p.sums = sums; p.c = c; p.hw = h * w; p.planes = planes;
  p.out_planes = out_planes;
  dim3 grid(tb_splits(2ll * h * w, (long long)n * (c / 8)), n * (c / 8));   // two threads (channel halves) per pixel
  if (p.dimg && p.dxs) k_sg2_layer_bwd<true, true><<<grid, TB_THREADS, 0, TB_STREAM>>>(p);
  else if (p.dimg) k_sg2_layer_bwd<true, false><<<grid, TB_THREADS, 0, TB_STREAM>>>(p);
  else if (p.dxs) k_sg2_layer_bwd<false, true><<<grid, TB_THREADS, 0, TB_STREAM>>>(p);
  else k_sg2_layer_bwd<false, false><<<grid, TB_THREADS, 0, TB_STREAM>>>(p);
  count_launch();
  return check_launch("k_sg2_layer_bwd");
}

extern "C" int dge_up_fir_bwd_s2d(const float* dconv, void* out_act, int n, int c, int h, int w, int planes, int box,
                                  void* stream) {
  DGE_REQUIRE(dconv && out_act, "up_fir_bwd_s2d: null pointer");
  DGE_REQUIRE(n > 0 && c >= 16 && c % 16 == 0 && h > 0 && w > 0, "up_fir_bwd_s2d: bad dims n=%d c=%d h=%d w=%d", n, c, h, w);
  DGE_REQUIRE(planes == 1 || planes == 2, "up_fir_bwd_s2d: planes=%d", planes);
  const size_t work = (size_t)n * (c / 8) * (h + 1) * (w + 1);
  DGE_REQUIRE(work < (1ull << 31) && (size_t)4 * h * w < (1ull << 31),
              "up_fir_bwd_s2d: map too large for 32-bit indexing (n=%d c=%d h=%d w=%d)", n, c, h, w);
  size_t g = (work + TB_THREADS - 1) / TB_THREADS;
  if (g > (size_t)tb_sms() * 32) g = (size_t)tb_sms() * 32;
  if (box) k_up_fir_bwd_s2d<true><<<(int)g, TB_THREADS, 0, TB_STREAM>>>(dconv, out_act, n, c, h, w, planes);
  else k_up_fir_bwd_s2d<false><<<(int)g, TB_THREADS, 0, TB_STREAM>>>(dconv, out_act, n, c, h, w, planes);
  count_launch();
  return check_launch("k_up_fir_bwd_s2d");
}

extern "C" int dge_rgb_up_bwd(const float* d_out, float* d_in, int64_t planes, int h_in, int w_in, void* stream) {
  DGE_REQUIRE(d_out && d_in && planes > 0 && h_in > 0 && w_in > 0, "rgb_up_bwd: bad arguments");
  const size_t work = (size_t)planes * h_in * w_in;
  size_t g = (work + 255) / 256;
  if (g > (size_t)tb_sms() * 32) g = (size_t)tb_sms() * 32;
  k_rgb_up_bwd<<<(int)g, 256, 0, TB_STREAM>>>(d_out, d_in, (size_t)planes, h_in, w_in);
  count_launch();
  return check_launch("k_rgb_up_bwd");
}

static int tb_grid1d(size_t work, int block) {
  size_t g = (work + block - 1) / block;
  if (g > (size_t)tb_sms() * 32) g = (size_t)tb_sms() * 32;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int dge_lpips_input(const float* x, void* out_act, float shift0, float shift1, float shift2, float scale0,
                               float scale1, float scale2, int n, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(x && out_act && n > 0 && h > 0 && w > 0 && (planes == 1 || planes == 2) && scale0 != 0.f && scale1 != 0.f &&
                  scale2 != 0.f, "lpips_input: bad arguments");
  k_lpips_input<<<tb_grid1d((size_t)n * h * w, 256), 256, 0, TB_STREAM>>>(x, out_act, n, h, w, planes, shift0, shift1,
                                                                          shift2, 1.f / scale0, 1.f / scale1, 1.f / scale2);
  count_launch();
  return check_launch("k_lpips_input");
}

extern "C" int dge_maxpool_to_act(const float* x, void* out_act, int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(x && out_act && n > 0 && c >= 8 && c % 8 == 0 && h >= 2 && w >= 2 && (planes == 1 || planes == 2),
              "maxpool_to_act: bad arguments");
  k_maxpool_to_act<<<tb_grid1d((size_t)n * (c / 8) * (h / 2) * (w / 2), 256), 256, 0, TB_STREAM>>>(x, out_act, n, c, h, w,
                                                                                                  planes);
  count_launch();
  return check_launch("k_maxpool_to_act");
}

extern "C" int dge_relu_pool_bwd(const float* y_f32b, const void* y_act, int y_planes, const float* g_same,
                                 const float* g_pool, void* out_act, int n, int c, int h, int w, int planes, int clamp_pos,
                                 void* stream) {
  DGE_REQUIRE((y_f32b || y_act) && !(y_f32b && y_act) && out_act && (g_same || g_pool), "relu_pool_bwd: bad pointers");
  DGE_REQUIRE(n > 0 && c >= 8 && c % 8 == 0 && h > 0 && w > 0 && (planes == 1 || planes == 2) &&
                  (!y_act || y_planes == 1 || y_planes == 2), "relu_pool_bwd: bad dims");
  DGE_REQUIRE(!g_pool || (h >= 2 && w >= 2), "relu_pool_bwd: pooled gradient needs h, w >= 2");
  k_relu_pool_bwd<<<tb_grid1d((size_t)n * (c / 8) * ((h + 1) / 2) * ((w + 1) / 2), 256), 256, 0, TB_STREAM>>>(
      y_f32b, y_act, y_planes, g_same, g_pool, out_act, n, c, h, w, planes, clamp_pos);
  count_launch();
  return check_launch("k_relu_pool_bwd");
}

extern "C" int dge_lpips_dist(const float* f, const float* lin_w, float* out, const float* go, float* ga, float* gb,
                              int nb, int c, int h, int w, float eps, void* stream) {
  DGE_REQUIRE(f && lin_w && nb > 0 && c >= 8 && c % 8 == 0 && h > 0 && w > 0, "lpips_dist: bad arguments");
  const int mode = (gb ? 1 : 0) | (ga ? 2 : 0);
  DGE_REQUIRE(mode ? go != nullptr : out != nullptr, "lpips_dist: forward needs out, backward needs go");
  long long blocks = ((long long)h * w * LP_CS + 255) / 256, want = ((long long)tb_sms() * 8 + nb - 1) / nb;
  if (blocks > want) blocks = want;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)nb);
  k_lpips_dist<<<grid, 256, 0, TB_STREAM>>>(f, lin_w, out, go, ga, gb, mode, nb, c, h * w, eps);
  count_launch();
  return check_launch("k_lpips_dist");
}
