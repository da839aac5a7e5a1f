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
// blocks per (sample, channel group): enough to fill the chip ~8 blocks per SM, never more than the pixels allow
static int tb_splits(long long pixels, long long groups) {
  long long want = ((long long)tb_sms() * 8 + groups - 1) / groups;
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
__global__ void __launch_bounds__(TB_THREADS)
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
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hwp; i += (size_t)gridDim.x * blockDim.x) {
    const int py = (int)(i / wo), px = (int)(i - (size_t)py * wo);
    float d[8], r[8];
    load8_f32b(d_out, (size_t)ng * hwp + i, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      r[k] = gb * d[k];
      acc[16 + k] += r[k];
      d[k] *= ga4;
    }
    if (dres) store8_act_at(dres, (size_t)ng * planes * hwp + i, hwp, planes, r);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const size_t pix = (size_t)(2 * py + dy) * W + (2 * px + dx);
        float y[8], v[8];
        load8_f32b(y2, (size_t)ng * hw + pix, y);
        const float nz = noise ? __ldg(noise + (size_t)nidx * hw + pix) : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = y[k] > 0.f ? d[k] : d[k] * slope;
          acc[k] += v[k];
          acc[8 + k] = fmaf(v[k], nz, acc[8 + k]);
        }
        store8_act_at(dy2, (size_t)ng * planes * hw + pix, hw, planes, v);
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
  const size_t base = (size_t)ng * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)hw; i += (size_t)gridDim.x * blockDim.x) {
    float gv[8], xv[8];
    load8_f32b(g, base + i, gv);
    load8_f32b(x, base + i, xv);
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
__global__ void __launch_bounds__(TB_THREADS)
k_in_bwd_apply(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mr,
               const float* __restrict__ style, const float* __restrict__ dstyle, const double* __restrict__ sums,
               int mode, const float* __restrict__ res, float rscale, int res_pool, const float* __restrict__ noise,
               float slope, float* __restrict__ out_f32b, void* __restrict__ out_act, float* __restrict__ sums2, int c,
               int h, int w, int planes) {
  __shared__ float red[TB_THREADS / 32][16];
  const int C8 = c >> 3, ng = blockIdx.y, nidx = ng / C8, grp = ng - nidx * C8;
  const size_t hw = (size_t)h * w;
  const float inv_hw = 1.f / (float)hw;
  float m[8], r[8], a[8], b[8], cm[8], cs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = grp * 8 + k;
    const size_t o = ((size_t)nidx * c + ch) * 2;
    m[k] = __ldg(mr + o);
    r[k] = __ldg(mr + o + 1);
    a[k] = (float)(sums[o] / (double)hw);
    b[k] = (float)(sums[o + 1] / (double)hw);
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
  const int wr = res_pool ? (w >> 1) : w;
  const size_t rbase = (size_t)ng * (res_pool ? (hw >> 2) : hw);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    float gv[8], xv[8], v[8];
    load8_f32b(g, base + i, gv);
    load8_f32b(x, base + i, xv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xc = xv[k] - m[k];
      v[k] = r[k] * (gv[k] - a[k] - xc * r[k] * b[k]) + cm[k] + cs[k] * xc;
    }
    if (mode == 0) {
      if (res) {
        size_t ri = i;
        if (res_pool) {
          const int y = (int)(i / w), xx = (int)(i - (size_t)y * w);
          ri = (size_t)(y >> 1) * wr + (xx >> 1);
        }
        float rv[8];
        load8_f32b(res, rbase + ri, rv);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaf(rscale, rv[k], v[k]);
      }
      store8_f32b(out_f32b, base + i, v);
    } else {
      const float nz = noise ? __ldg(noise + (size_t)nidx * hw + i) : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = xv[k] > 0.f ? v[k] : v[k] * slope;
        acc[k] += v[k];
        acc[8 + k] = fmaf(v[k], nz, acc[8 + k]);
      }
      store8_act_at(out_act, (size_t)ng * planes * hw + i, hw, planes, v);
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
k_from_rgb_bwd(const float* __restrict__ d_f, const float* __restrict__ f, const float* __restrict__ img, float slope,
               float* __restrict__ sums, int cimg, int c, int hw) {
  __shared__ float red[TB_THREADS / 32][32];
  const int C8 = c >> 3, ng = blockIdx.y, nidx = ng / C8, grp = ng - nidx * C8;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const size_t base = (size_t)ng * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)hw; i += (size_t)gridDim.x * blockDim.x) {
    float dv[8], fv[8], px[3] = {0.f, 0.f, 0.f};
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
    }
  }
  const float t = block_sums<32>(acc, red);
  if (threadIdx.x < 32) atomicAdd(sums + (size_t)(grp * 8 + (threadIdx.x >> 2)) * 4 + (threadIdx.x & 3), t);
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
                                const float* dstyle, const double* sums, int mode, const float* res, float rscale,
                                int res_pool, const float* noise, float slope, float* out_f32b, void* out_act,
                                float* sums2, int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(g && x && mean_rstd && sums, "in_bwd_apply: null pointer");
  DGE_REQUIRE(mode == 0 || mode == 1, "in_bwd_apply: mode=%d", mode);
  DGE_REQUIRE(n > 0 && c >= 8 && c % 8 == 0 && h > 0 && w > 0, "in_bwd_apply: bad dims n=%d c=%d h=%d w=%d", n, c, h, w);
  DGE_REQUIRE(!dstyle || style, "in_bwd_apply: dstyle needs style (mean || std)");
  if (mode == 0) {
    DGE_REQUIRE(out_f32b, "in_bwd_apply: mode 0 writes out_f32b");
    DGE_REQUIRE(!res || !res_pool || (h % 2 == 0 && w % 2 == 0), "in_bwd_apply: pooled residual needs even h, w");
  } else {
    DGE_REQUIRE(out_act && sums2 && (planes == 1 || planes == 2) && c % 16 == 0,
                "in_bwd_apply: mode 1 writes out_act (planes 1|2, c %% 16 == 0) and sums2");
    TB_ZERO(sums2, (size_t)2 * c * sizeof(float));
  }
  dim3 grid(tb_splits((long long)h * w, (long long)n * (c / 8)), n * (c / 8));
  k_in_bwd_apply<<<grid, TB_THREADS, 0, TB_STREAM>>>(g, x, mean_rstd, style, dstyle, sums, mode, res, rscale, res_pool,
                                                      noise, slope, out_f32b, out_act, sums2, c, h, w, planes);
  count_launch();
  return check_launch("k_in_bwd_apply");
}

extern "C" int dge_from_rgb_bwd(const float* d_f, const float* f, const float* img, float slope, float* sums, int n,
                                int cimg, int c, int h, int w, void* stream) {
  DGE_REQUIRE(d_f && f && img && sums, "from_rgb_bwd: null pointer");
  DGE_REQUIRE(n > 0 && cimg >= 1 && cimg <= 3 && c >= 8 && c % 8 == 0 && h > 0 && w > 0,
              "from_rgb_bwd: bad dims n=%d cimg=%d c=%d h=%d w=%d", n, cimg, c, h, w);
  TB_ZERO(sums, (size_t)4 * c * sizeof(float));
  dim3 grid(tb_splits((long long)h * w, (long long)n * (c / 8)), n * (c / 8));
  k_from_rgb_bwd<<<grid, TB_THREADS, 0, TB_STREAM>>>(d_f, f, img, slope, sums, cimg, c, h * w);
  count_launch();
  return check_launch("k_from_rgb_bwd");
}
