// elementwise.cu -- the HBM-bound kernels around the conv: layout changes, weight packing, style/demod GEMVs,
// FIR up-sampling epilogue, ToRGB skip initialisation, encoder statistics / instance-norm / pooling / blend.
// Every kernel is one pass over its data with 16/32-byte vector accesses; one thread handles one
// (pixel, 8-channel group) unless noted.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "dge_common.cuh"

namespace dge {

// ---------------------------------------------------------------------------------------------
// error / bookkeeping
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};   // process-wide: autograd runs backward nodes on its own worker thread

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DGE_ERR_CUDA;
  }
  return DGE_OK;
}

static inline int grid_for(size_t work, int block) {
  size_t g = (work + block - 1) / block;
  if (g > (size_t)148 * 64) g = (size_t)148 * 64;  // grid-stride beyond this
  if (g < 1) g = 1;
  return (int)g;
}

#define LAUNCH_1D(kernel, work, stream, ...)                                     \
  do {                                                                           \
    kernel<<<grid_for((work), 256), 256, 0, (cudaStream_t)(stream)>>>(__VA_ARGS__); \
    count_launch();                                                              \
    return check_launch(#kernel);                                                \
  } while (0)

// ---------------------------------------------------------------------------------------------
// weight preparation
// ---------------------------------------------------------------------------------------------
// WPK [taps][Cin/8][planes][Cout][8]; one thread per (tap, cin-group, cout)
// transpose = 1 packs the data-gradient weight of the same layer: W'[i][o][ky][kx] = W[o][i][k-1-ky][k-1-kx], i.e. the
// returned operand has `cin` output columns and contracts over `cout` (cout/cin below are those of the PACKED operand;
// src_cin is the input-channel count of the stored OIHW tensor)
__global__ void k_pack_conv_weight(const float* __restrict__ w, uint4* __restrict__ out, int cout, int cin, int ks,
                                   int flip, float scale, int planes, int transpose, int src_cin) {
  const int taps = ks * ks, C8 = cin >> 3;
  const size_t total = (size_t)taps * C8 * cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    const int g = (int)((i / cout) % C8);
    const int tap = (int)(i / ((size_t)cout * C8));
    int ky = tap / ks, kx = tap % ks;
    if (flip) { ky = ks - 1 - ky; kx = ks - 1 - kx; }
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const size_t so = transpose ? (size_t)(g * 8 + k) * src_cin + co : (size_t)co * src_cin + g * 8 + k;
      v[k] = w[(so * ks + ky) * ks + kx] * scale;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    const size_t o = (((size_t)tap * C8 + g) * planes) * cout + co;
    out[o] = hi;
    if (planes == 2) out[o + cout] = lo;
  }
}

__global__ void k_weight_sqsum(const float* __restrict__ w, float* __restrict__ w2, int cout, int cin, int kk,
                               float scale) {
  const size_t total = (size_t)cout * cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < kk; ++k) {
      const float t = w[i * kk + k] * scale;
      s = fmaf(t, t, s);
    }
    w2[i] = s;
  }
}

// one warp per (n, o)
__global__ void k_demod(const float* __restrict__ w2, const float* __restrict__ style, float* __restrict__ d, int n,
                        int cout, int cin, float eps) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int idx = gw; idx < n * cout; idx += nw) {
    const int b = idx / cout, o = idx % cout;
    float s = 0.f;
    for (int i = lane; i < cin; i += 32) {
      const float st = style[(size_t)b * cin + i];
      s = fmaf(w2[(size_t)o * cin + i], st * st, s);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) d[idx] = rsqrtf(s + eps);
  }
}

__global__ void k_rgb_weights(const float* __restrict__ w, const float* __restrict__ style, float* __restrict__ out,
                              int n, int nch, int cin, float scale) {
  const size_t total = (size_t)n * nch * cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin);
    const int ch = (int)((i / cin) % nch);
    const int b = (int)(i / ((size_t)cin * nch));
    out[i] = w[(size_t)ch * cin + c] * scale * style[(size_t)b * cin + c];
  }
}

// one warp per (n, m)
__global__ void k_dense(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                        float* __restrict__ y, int n, int k, int m, float wscale, float bscale, float add_bias,
                        float slope, float gain) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int idx = gw; idx < n * m; idx += nw) {
    const int r = idx / m, o = idx % m;
    float s = 0.f;
    for (int i = lane; i < k; i += 32) s = fmaf(x[(size_t)r * k + i], w[(size_t)o * k + i], s);
#pragma unroll
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
      float v = s * wscale + (b ? b[o] * bscale : 0.f) + add_bias;
      v = (v < 0.f ? v * slope : v) * gain;
      y[idx] = v;
    }
  }
}

// one warp per row
// one block per (layer item, sample): style affine -> shared memory -> demod coefficients / ToRGB weights.
// Both contractions are latency-bound GEMVs: a warp works on 4 output rows at once with 16-byte loads (16 independent
// loads in flight per lane) against a vector held in shared memory.
__device__ __forceinline__ void gemv4_rows(const float* __restrict__ w, int ld, int len, const float* sv, int lane,
                                           float* acc) {
  acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
  for (int k = lane * 4; k < len; k += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(sv + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)j * ld + k));
      acc[j] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[j]))));
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int off = 16; off; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
}

__global__ void __launch_bounds__(256)
k_sg2_prep(const dge_sg2_prep_item* __restrict__ items, const float* __restrict__ wp, float* __restrict__ arena,
           int num_layers, int wdim) {
  __shared__ __align__(16) float s_x[2048];       // the w vector, then style^2 for the demod contraction
  __shared__ __align__(16) float s_style[2048];
  const dge_sg2_prep_item it = items[blockIdx.x];
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* x = wp + ((size_t)n * num_layers + it.wp_index) * wdim;
  for (int k = threadIdx.x; k < wdim; k += blockDim.x) s_x[k] = __ldg(x + k);
  __syncthreads();
  for (int c = warp * 4; c < it.cin; c += nwarps * 4) {   // cin, wdim are multiples of 4 (host-checked)
    float acc[4];
    gemv4_rows(it.st_w + (size_t)c * wdim, wdim, wdim, s_x, lane, acc);
    if (lane < 4) {
      const int cc = c + lane;
      const float a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
      const float v = a * it.st_wscale + (it.st_b ? __ldg(it.st_b + cc) * it.st_bscale : 0.f) + it.st_add_bias;
      s_style[cc] = v;
      if (it.style_off >= 0) arena[it.style_off + (size_t)n * it.cin + cc] = v;
    }
  }
  __syncthreads();
  if (it.w2 && it.demod_off >= 0) {
    for (int k = threadIdx.x; k < it.cin; k += blockDim.x) s_x[k] = s_style[k] * s_style[k];
    __syncthreads();
    for (int o = warp * 4; o < it.cout; o += nwarps * 4) {
      float acc[4];
      gemv4_rows(it.w2 + (size_t)o * it.cin, it.cin, it.cin, s_x, lane, acc);
      if (lane < 4) {
        const float a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
        arena[it.demod_off + (size_t)n * it.cout + o + lane] = rsqrtf(a + it.eps);
      }
    }
  }
  if (it.rgb_w && it.rgbw_off >= 0) {
    const int total = it.nch * it.cin;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int c = i % it.cin;
      arena[it.rgbw_off + (size_t)n * total + i] = __ldg(it.rgb_w + i) * it.rgb_scale * s_style[c];
    }
  }
}

// Transpose of k_sg2_prep (training: the frozen generator's gradient w.r.t. wp).  Block = (item, sample).
//   layer item : ds[c] = S[c] - style[c] * sum_o (D[o] * demod[o]^2) * w2[o][c]      (S = d style, D = demod * d demod)
//   ToRGB item : ds[c] = sum_j T[c][j] * rgb_w[j][c] * rgb_scale                       (T = d of the modulated ToRGB weights)
//   d_wp[n][wp_index][k] += st_wscale * sum_c ds[c] * st_w[c][k]
// gsrc[item] = (s_off, s_stride, d_off, t_off): float offsets into `sums` (-1 = absent)
__global__ void __launch_bounds__(256)
k_sg2_prep_bwd(const dge_sg2_prep_item* __restrict__ items, const float* __restrict__ arena,
               const long long* __restrict__ gsrc, const float* __restrict__ sums, float* __restrict__ d_wp,
               int num_layers, int wdim) {
  __shared__ float s_q[2048];
  __shared__ float s_ds[2048];
  const dge_sg2_prep_item it = items[blockIdx.x];
  const int n = blockIdx.y;
  const long long s_off = gsrc[4 * blockIdx.x], s_stride = gsrc[4 * blockIdx.x + 1], d_off = gsrc[4 * blockIdx.x + 2],
                  t_off = gsrc[4 * blockIdx.x + 3];
  if (t_off >= 0) {
    for (int c = threadIdx.x; c < it.cin; c += blockDim.x) {
      float ds = 0.f;
      for (int j = 0; j < it.nch; ++j)
        ds = fmaf(__ldg(sums + t_off + ((size_t)n * it.cin + c) * 5 + j), __ldg(it.rgb_w + (size_t)j * it.cin + c), ds);
      s_ds[c] = ds * it.rgb_scale;
    }
  } else {
    const bool dem = it.w2 && d_off >= 0 && it.demod_off >= 0;
    if (dem) {
      for (int o = threadIdx.x; o < it.cout; o += blockDim.x) {
        const float dm = arena[it.demod_off + (size_t)n * it.cout + o];
        s_q[o] = __ldg(sums + d_off + ((size_t)n * it.cout + o) * 5) * dm * dm;
      }
      __syncthreads();
    }
    for (int c = threadIdx.x; c < it.cin; c += blockDim.x) {
      float ds = s_off >= 0 ? __ldg(sums + s_off + ((size_t)n * it.cin + c) * s_stride) : 0.f;
      if (dem) {
        float r = 0.f;
        for (int o = 0; o < it.cout; ++o) r = fmaf(s_q[o], __ldg(it.w2 + (size_t)o * it.cin + c), r);
        ds -= arena[it.style_off + (size_t)n * it.cin + c] * r;
      }
      s_ds[c] = ds;
    }
  }
  __syncthreads();
  float* dst = d_wp + ((size_t)n * num_layers + it.wp_index) * wdim;
  for (int k = threadIdx.x; k < wdim; k += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < it.cin; ++c) acc = fmaf(s_ds[c], __ldg(it.st_w + (size_t)c * wdim + k), acc);
    atomicAdd(dst + k, acc * it.st_wscale);      // a layer and the ToRGB that shares its wp row both land here
  }
}

__global__ void k_pixel_norm(const float* __restrict__ x, float* __restrict__ y, int n, int k, float eps) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int r = gw; r < n; r += nw) {
    float s = 0.f;
    for (int i = lane; i < k; i += 32) {
      const float t = x[(size_t)r * k + i];
      s = fmaf(t, t, s);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float inv = 1.f / sqrtf(s / (float)k + eps);
    for (int i = lane; i < k; i += 32) y[(size_t)r * k + i] = x[(size_t)r * k + i] * inv;
  }
}

// ---------------------------------------------------------------------------------------------
// layout conversions.  index space: (n, c8, y, x) with x fastest
// ---------------------------------------------------------------------------------------------
struct Idx4 {
  int n, g, y, x;
};
__device__ __forceinline__ Idx4 decode4(size_t i, int C8, int H, int W) {
  Idx4 r;
  r.x = (int)(i % W);
  size_t t = i / W;
  r.y = (int)(t % H);
  t /= H;
  r.g = (int)(t % C8);
  r.n = (int)(t / C8);
  return r;
}

__global__ void k_nchw_to_act(const float* __restrict__ x, long long bstride, const float* __restrict__ scale,
                              void* __restrict__ act, int n, int c, int h, int w, int planes) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, h, w);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = q.g * 8 + k;
      float t = x[(size_t)q.n * bstride + ((size_t)ch * h + q.y) * w + q.x];
      if (scale) t *= scale[(size_t)q.n * c + ch];
      v[k] = t;
    }
    store8_act(act, q.n, q.g, q.y, q.x, C8, planes, h, w, v);
  }
}

__global__ void k_nchw_to_f32b(const float* __restrict__ x, float* __restrict__ out, int n, int c, int h, int w) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, h, w);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = x[(((size_t)q.n * c + q.g * 8 + k) * h + q.y) * w + q.x];
    store8_f32b(out, i, v);
  }
}

__global__ void k_f32b_to_nchw(const float* __restrict__ x, float* __restrict__ out, int n, int c, int h, int w) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, h, w);
    float v[8];
    load8_f32b(x, i, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[(((size_t)q.n * c + q.g * 8 + k) * h + q.y) * w + q.x] = v[k];
  }
}

__global__ void k_act_to_nchw(const void* __restrict__ act, float* __restrict__ out, int n, int c, int h, int w,
                              int planes) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, h, w);
    float v[8];
    load8_act(act, q.n, q.g, q.y, q.x, C8, planes, h, w, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[(((size_t)q.n * c + q.g * 8 + k) * h + q.y) * w + q.x] = v[k];
  }
}

__global__ void k_f32b_to_act(const float* __restrict__ x, void* __restrict__ act, int n, int c, int h, int w,
                              int planes) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, h, w);
    float v[8];
    load8_f32b(x, i, v);
    store8_act(act, q.n, q.g, q.y, q.x, C8, planes, h, w, v);
  }
}

// standalone ToRGB (1x1 modulated conv, no demod): NCHW in, NCHW out; one thread per (n, y, x)
__global__ void k_to_rgb_nchw(const float* __restrict__ x, const float* __restrict__ rgb_w,
                              const float* __restrict__ bias, float* __restrict__ out, int n, int c, int nch, int hw) {
  const size_t total = (size_t)n * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const size_t pix = i % hw;
    for (int ch = 0; ch < nch; ++ch) {
      float s = bias ? bias[ch] : 0.f;
      const float* wv = rgb_w + ((size_t)b * nch + ch) * c;
      for (int k = 0; k < c; ++k) s = fmaf(x[((size_t)b * c + k) * hw + pix], wv[k], s);
      out[((size_t)b * nch + ch) * hw + pix] = s;
    }
  }
}

// skip-branch RGB upsample: zero-insert x2, pad (2,1), 4x4 FIR (f = [1,3,3,1]/4 per axis after the x4 gain):
//   out[2m]   = (x[m-1] + 3 x[m]) / 4 ,  out[2m+1] = (3 x[m] + x[m+1]) / 4   per axis
__global__ void k_rgb_init(const float* __restrict__ in, const float* __restrict__ bias, float* __restrict__ out,
                           int n, int nch, int ho, int wo) {
  const size_t total = (size_t)n * nch * ho * wo;
  const int hin = ho >> 1, win = wo >> 1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo);
    size_t tt = i / wo;
    const int y = (int)(tt % ho);
    tt /= ho;
    const int ch = (int)(tt % nch);
    float v = bias ? bias[ch] : 0.f;
    if (in) {
      const float* src = in + tt * (size_t)hin * win;
      const int my = y >> 1, mx = x >> 1;
      int y0, y1, x0, x1;
      float wy0, wy1, wx0, wx1;
      if (y & 1) { y0 = my; y1 = my + 1; wy0 = 0.75f; wy1 = 0.25f; } else { y0 = my - 1; y1 = my; wy0 = 0.25f; wy1 = 0.75f; }
      if (x & 1) { x0 = mx; x1 = mx + 1; wx0 = 0.75f; wx1 = 0.25f; } else { x0 = mx - 1; x1 = mx; wx0 = 0.25f; wx1 = 0.75f; }
      float s = 0.f;
      const bool vy0 = y0 >= 0 && y0 < hin, vy1 = y1 >= 0 && y1 < hin;
      const bool vx0 = x0 >= 0 && x0 < win, vx1 = x1 >= 0 && x1 < win;
      if (vy0 && vx0) s = fmaf(wy0 * wx0, src[(size_t)y0 * win + x0], s);
      if (vy0 && vx1) s = fmaf(wy0 * wx1, src[(size_t)y0 * win + x1], s);
      if (vy1 && vx0) s = fmaf(wy1 * wx0, src[(size_t)y1 * win + x0], s);
      if (vy1 && vx1) s = fmaf(wy1 * wx1, src[(size_t)y1 * win + x1], s);
      v += s;
    }
    out[i] = v;
  }
}

// vectorised form: one thread = 2 output rows x 4 output columns (12 input reads, two 16-byte stores); wo % 4 == 0
__global__ void k_rgb_init_v4(const float* __restrict__ in, const float* __restrict__ bias, float* __restrict__ out,
                              int n, int nch, int ho, int wo) {
  const int hin = ho >> 1, win = wo >> 1, wq = win >> 1;
  const size_t total = (size_t)n * nch * hin * wq;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % wq);
    size_t tt = i / wq;
    const int my = (int)(tt % hin);
    tt /= hin;
    const int ch = (int)(tt % nch);
    const float b = bias ? bias[ch] : 0.f;
    const float* src = in + tt * (size_t)hin * win;
    const int c0 = 2 * q;
    float r[3][4];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = my - 1 + dy;
      const bool vy = yy >= 0 && yy < hin;
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const int xx = c0 - 1 + dx;
        r[dy][dx] = (vy && xx >= 0 && xx < win) ? __ldg(src + (size_t)yy * win + xx) : 0.f;
      }
    }
    float hz[3][4];   // horizontal pass: 4 output columns per input row
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      hz[dy][0] = 0.25f * r[dy][0] + 0.75f * r[dy][1];
      hz[dy][1] = 0.75f * r[dy][1] + 0.25f * r[dy][2];
      hz[dy][2] = 0.25f * r[dy][1] + 0.75f * r[dy][2];
      hz[dy][3] = 0.75f * r[dy][2] + 0.25f * r[dy][3];
    }
    float4 o0, o1;
    o0.x = b + 0.25f * hz[0][0] + 0.75f * hz[1][0]; o1.x = b + 0.75f * hz[1][0] + 0.25f * hz[2][0];
    o0.y = b + 0.25f * hz[0][1] + 0.75f * hz[1][1]; o1.y = b + 0.75f * hz[1][1] + 0.25f * hz[2][1];
    o0.z = b + 0.25f * hz[0][2] + 0.75f * hz[1][2]; o1.z = b + 0.75f * hz[1][2] + 0.25f * hz[2][2];
    o0.w = b + 0.25f * hz[0][3] + 0.75f * hz[1][3]; o1.w = b + 0.75f * hz[1][3] + 0.25f * hz[2][3];
    float* dst = out + (tt * (size_t)ho + 2 * my) * wo + 4 * q;
    *reinterpret_cast<float4*>(dst) = o0;
    *reinterpret_cast<float4*>(dst + wo) = o1;
  }
}

// ---------------------------------------------------------------------------------------------
// encoder pieces
// ---------------------------------------------------------------------------------------------
// one thread per pixel: the (<= 4) image channels are read once, every output channel group is produced from weights
// staged in shared memory (broadcast reads), stores are 32-byte vectors contiguous across the warp
__global__ void __launch_bounds__(256)
k_from_rgb(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ b,
           float* __restrict__ out, int n, int cimg, int c, int h, int wd, float slope) {
  extern __shared__ __align__(16) float sw[];            // [c][4] weights (zero-padded beyond cimg) then [c] bias
  float* sb = sw + 4 * c;
  for (int i = threadIdx.x; i < c * 4; i += blockDim.x) sw[i] = ((i & 3) < cimg) ? w[(i >> 2) * cimg + (i & 3)] : 0.f;
  for (int i = threadIdx.x; i < c; i += blockDim.x) sb[i] = b ? b[i] : 0.f;
  __syncthreads();
  const int C8 = c >> 3;
  const size_t hw = (size_t)h * wd, total = (size_t)n * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bn = i / hw, pix = i - bn * hw;
    float px[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ci = 0; ci < 4; ++ci)
      if (ci < cimg) px[ci] = __ldg(img + (bn * cimg + ci) * hw + pix);
    for (int g = 0; g < C8; ++g) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ch = g * 8 + k;
        const float4 wv = *reinterpret_cast<const float4*>(sw + 4 * ch);   // one 16-byte broadcast read per channel
        const float a = fmaf(px[0], wv.x, fmaf(px[1], wv.y, fmaf(px[2], wv.z, fmaf(px[3], wv.w, sb[ch]))));
        v[k] = a < 0.f ? a * slope : a;
      }
      store8_f32b(out, (bn * C8 + g) * hw + pix, v);
    }
  }
}

// grid (splits, n*C8); fp64 accumulation (naive fp32 E[x^2]-E[x]^2 loses the variance when |mean| >> std)
__global__ void k_instance_stats_partial(const float* __restrict__ x, double* __restrict__ scratch, int c, int hw) {
  const int C8 = c >> 3;
  const int ng = blockIdx.y;  // n*C8 + g
  const int nidx = ng / C8, g = ng % C8;
  double s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.0;
  // fp32 partial sums over groups of 4 pixels (exact enough: 4 terms), fp64 across groups: a quarter of the fp64 work,
  // and the 4 loads of a group are independent (the kernel ran at 73 % of the copy bandwidth with one load in flight)
  const float* xb = x + (size_t)ng * hw * 8;
  const unsigned stride = gridDim.x * blockDim.x, uhw = (unsigned)hw;
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < uhw; i0 += 4 * stride) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned i = i0 + u * stride;
      if (i < uhw) load8_f32b(xb, i, v[u]);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[u][k] = 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float a = (v[0][k] + v[1][k]) + (v[2][k] + v[3][k]);
      const float q = fmaf(v[0][k], v[0][k], v[1][k] * v[1][k]) + fmaf(v[2][k], v[2][k], v[3][k] * v[3][k]);
      s1[k] += (double)a;
      s2[k] += (double)q;
    }
  }
  __shared__ double red[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], off);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], off);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[warp][2 * k] = s1[k];
      red[warp][2 * k + 1] = s2[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) t += red[wv][threadIdx.x];
    const int k = threadIdx.x >> 1, which = threadIdx.x & 1;
    atomicAdd(&scratch[((size_t)nidx * c + g * 8 + k) * 2 + which], t);
  }
}

// small maps (h*w <= 4096): one block per (sample, channel group) does the whole reduction and writes the results --
// one launch instead of memset + partial + final (the three launches cost ~16 us where the data is a few KB)
__global__ void __launch_bounds__(256)
k_instance_stats_small(const float* __restrict__ x, float* __restrict__ style, float* __restrict__ mean_rstd, int c,
                       int hw, float eps) {
  const int C8 = c >> 3;
  const int ng = blockIdx.x;
  const int nidx = ng / C8, g = ng % C8;
  double s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.0;
  const size_t base = (size_t)ng * hw;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    float v[8];
    load8_f32b(x, base + i, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s1[k] += (double)v[k];
      s2[k] += (double)v[k] * (double)v[k];
    }
  }
  __shared__ double red[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], off);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], off);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[warp][2 * k] = s1[k];
      red[warp][2 * k + 1] = s2[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int k = threadIdx.x;
    double a = 0.0, q = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) {
      a += red[wv][2 * k];
      q += red[wv][2 * k + 1];
    }
    const double m = a / hw;
    double var = q / hw - m * m;
    if (var < 0.0) var = 0.0;
    const int ch = g * 8 + k;
    if (style) {
      style[(size_t)nidx * 2 * c + ch] = (float)m;
      style[(size_t)nidx * 2 * c + c + ch] = (float)sqrt(var);
    }
    if (mean_rstd) {
      mean_rstd[2 * ((size_t)nidx * c + ch)] = (float)m;
      mean_rstd[2 * ((size_t)nidx * c + ch) + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
}

__global__ void k_instance_stats_final(const double* __restrict__ scratch, float* __restrict__ style,
                                       float* __restrict__ mean_rstd, int n, int c, int hw, float eps) {
  const int total = n * c;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / c, ch = i % c;
    const double m = scratch[2 * (size_t)i] / hw;
    double var = scratch[2 * (size_t)i + 1] / hw - m * m;
    if (var < 0.0) var = 0.0;
    if (style) {
      style[(size_t)b * 2 * c + ch] = (float)m;
      style[(size_t)b * 2 * c + c + ch] = (float)sqrt(var);
    }
    if (mean_rstd) {
      mean_rstd[2 * (size_t)i] = (float)m;
      mean_rstd[2 * (size_t)i + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
}

__global__ void k_instance_norm(const float* __restrict__ x, const float* __restrict__ mr,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                void* __restrict__ out_act, float* __restrict__ out_f32b, int n, int c, int h, int w,
                                int planes) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, h, w);
    float v[8];
    load8_f32b(x, i, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const size_t s = ((size_t)q.n * c + q.g * 8 + k) * 2;
      v[k] = (v[k] - __ldg(mr + s)) * __ldg(mr + s + 1);
      if (gamma) v[k] = v[k] * __ldg(gamma + q.g * 8 + k) + __ldg(beta + q.g * 8 + k);
    }
    if (out_f32b) store8_f32b(out_f32b, i, v);
    if (out_act) store8_act(out_act, q.n, q.g, q.y, q.x, C8, planes, h, w, v);
  }
}

__device__ __forceinline__ void pooled8(const float* src, int n, int g, int y, int x, int C8, int ho, int wo,
                                        float* s) {
  float t[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      load8_f32b(src, f32b_idx32(n, g, 2 * y + dy, 2 * x + dx, C8, 2 * ho, 2 * wo), t);
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] += t[k];
    }
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] *= 0.25f;
}

// instance norm + zero-padded 3x3 blur; optional space-to-depth output (E_Blur.py:69-72)
__global__ void k_instance_norm_blur(const float* __restrict__ x, const float* __restrict__ mr, void* __restrict__ out,
                                     int s2d, int n, int c, int h, int w, int planes) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * h * w;
  const float bl[3] = {0.25f, 0.5f, 0.25f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, h, w);
    float m[8], r[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const size_t s = ((size_t)q.n * c + q.g * 8 + k) * 2;
      m[k] = __ldg(mr + s);
      r[k] = __ldg(mr + s + 1);
      acc[k] = 0.f;
    }
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = q.y + dy;
      if (y < 0 || y >= h) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = q.x + dx;
        if (xx < 0 || xx >= w) continue;
        float v[8];
        load8_f32b(x, f32b_idx32(q.n, q.g, y, xx, C8, h, w), v);
        const float wgt = bl[dy + 1] * bl[dx + 1];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, (v[k] - m[k]) * r[k], acc[k]);
      }
    }
    if (s2d) {
      const int ph = 2 * (q.y & 1) + (q.x & 1);
      store8_act(out, q.n, ph * C8 + q.g, q.y >> 1, q.x >> 1, 4 * C8, planes, h >> 1, w >> 1, acc);
    } else {
      store8_act(out, q.n, q.g, q.y, q.x, C8, planes, h, w, acc);
    }
  }
}

// instance norm of a 2x2 block of pixels + the 2x2 mean of the RAW input: one pass over x feeds both the conv_1
// operand (E.py:58) and the residual branch's avg_pool2d (E.py:78).  Thread = one pooled pixel of one channel group.
__global__ void k_instance_norm_pool(const float* __restrict__ x, const float* __restrict__ mr,
                                     void* __restrict__ out_act, void* __restrict__ out_pool, int n, int c, int ho,
                                     int wo, int planes) {
  const int C8 = c >> 3, h = 2 * ho, w = 2 * wo;
  const size_t total = (size_t)n * C8 * ho * wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, ho, wo);
    float m[8], r[8], s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const size_t o = ((size_t)q.n * c + q.g * 8 + k) * 2;
      m[k] = __ldg(mr + o);
      r[k] = __ldg(mr + o + 1);
      s[k] = 0.f;
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float v[8];
        load8_f32b(x, f32b_idx32(q.n, q.g, 2 * q.y + dy, 2 * q.x + dx, C8, h, w), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          s[k] += v[k];
          v[k] = (v[k] - m[k]) * r[k];
        }
        store8_act(out_act, q.n, q.g, 2 * q.y + dy, 2 * q.x + dx, C8, planes, h, w, v);
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] *= 0.25f;
    store8_act(out_pool, q.n, q.g, q.y, q.x, C8, planes, ho, wo, s);
  }
}

// FromRGB (net.py:231-240) that also accumulates the per-(sample, channel) sum and sum of squares of its OUTPUT
// (the first block's instance statistics, E.py:51-53,58) -- saves one full read of the 1024^2 feature map.
// grid (blocks, n); C8 channel groups (c <= 32).  fp32 partials over a thread's few pixels, fp64 across threads.
template <int C8T>
__global__ void __launch_bounds__(256)
k_from_rgb_stats(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ b,
                 float* __restrict__ out, double* __restrict__ scratch, int cimg, int h, int wd, float slope) {
  constexpr int C = C8T * 8;
  __shared__ __align__(16) float sw4[C * 4];
  __shared__ float sb[C];
  __shared__ double red[8][2 * C];
  for (int i = threadIdx.x; i < C * 4; i += blockDim.x) sw4[i] = ((i & 3) < cimg) ? w[(i >> 2) * cimg + (i & 3)] : 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = b ? b[i] : 0.f;
  __syncthreads();
  const int bn = blockIdx.y;
  const size_t hw = (size_t)h * wd;
  float s1[C], s2[C];
#pragma unroll
  for (int k = 0; k < C; ++k) s1[k] = s2[k] = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    float px[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ci = 0; ci < 4; ++ci)
      if (ci < cimg) px[ci] = __ldg(img + ((size_t)bn * cimg + ci) * hw + i);
#pragma unroll
    for (int g = 0; g < C8T; ++g) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ch = g * 8 + k;
        // weights as 16-byte broadcast reads of a [ch][4] table (zero-padded beyond cimg): 2 LDS per channel, not 4
        const float4 wv = *reinterpret_cast<const float4*>(sw4 + 4 * ch);
        const float a = fmaf(px[0], wv.x, fmaf(px[1], wv.y, fmaf(px[2], wv.z, fmaf(px[3], wv.w, sb[ch]))));
        v[k] = fmaxf(a, a * slope);
        s1[ch] += v[k];
        s2[ch] = fmaf(v[k], v[k], s2[ch]);
      }
      store8_f32b(out, ((size_t)bn * C8T + g) * hw + i, v);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < C; ++k) {
    // fp32 tree over the warp (512 pixels in all), fp64 from there on
    float a = s1[k], q = s2[k];
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    if (lane == 0) {
      red[warp][2 * k] = (double)a;
      red[warp][2 * k + 1] = (double)q;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    double t = 0.0;
    for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
    atomicAdd(&scratch[(size_t)bn * 2 * C + threadIdx.x], t);   // [n][c][2] like k_instance_stats_partial
  }
}

// x: F32B at (2*ho, 2*wo) -> ACT at (ho, wo)
__global__ void k_avgpool_to_act(const float* __restrict__ x, void* __restrict__ out, int n, int c, int ho, int wo,
                                 int planes) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * ho * wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, ho, wo);
    float s[8];
    pooled8(x, q.n, q.g, q.y, q.x, C8, ho, wo, s);
    store8_act(out, q.n, q.g, q.y, q.x, C8, planes, ho, wo, s);
  }
}

__global__ void k_blend(const float* __restrict__ a_src, const float* __restrict__ b_src, float* __restrict__ out,
                        float a, float b, int pool, int n, int c, int ho, int wo) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * ho * wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, ho, wo);
    float va[8], vb[8];
    if (pool & 1) pooled8(a_src, q.n, q.g, q.y, q.x, C8, ho, wo, va); else load8_f32b(a_src, i, va);
    if (pool & 2) pooled8(b_src, q.n, q.g, q.y, q.x, C8, ho, wo, vb); else load8_f32b(b_src, i, vb);
#pragma unroll
    for (int k = 0; k < 8; ++k) va[k] = a * va[k] + b * vb[k];
    store8_f32b(out, i, va);
  }
}

// ---------------------------------------------------------------------------------------------
// StyleGAN1 (model/stylegan1/net.py:141-169): what sits between a conv and the next instance norm.
//   mode 0: src = conv_1 output (F32B, same size)            -> Blur 3x3 [1,2,1]^2/16, zero padding (net.py:48-58)
//   mode 1: src = raw stride-2 transposed-conv map (2H+1)^2  -> 2x2 box sum (== the 4-shift `transform_kernel`
//           of lreq.py:127-131 with stride 2 / padding 1), then the same zero-padded Blur
//   mode 2: src = feature map, no filtering (first block: x = const)
// then  + noise_w[c]*noise[n,y,x] + bias[c] -> leaky_relu(slope)   (net.py:148-152)
// ---------------------------------------------------------------------------------------------
__global__ void k_sg1_post(const float* __restrict__ src, int mode, const float* __restrict__ noise,
                           const float* __restrict__ noise_w, const float* __restrict__ bias, float slope,
                           float* __restrict__ out, int n, int c, int ho, int wo) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * ho * wo;
  const int hs = mode == 1 ? ho + 1 : ho, ws = mode == 1 ? wo + 1 : wo;
  const float bl[3] = {0.25f, 0.5f, 0.25f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, ho, wo);
    float acc[8];
    if (mode == 2) {
      load8_f32b(src, i, acc);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = q.y + dy;
        if (y < 0 || y >= ho) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = q.x + dx;
          if (x < 0 || x >= wo) continue;
          const float wgt = bl[dy + 1] * bl[dx + 1];
          float v[8];
          if (mode == 0) {
            load8_f32b(src, f32b_idx32(q.n, q.g, y, x, C8, hs, ws), v);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, v[k], acc[k]);
          } else {
            float b4[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) b4[k] = 0.f;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                load8_f32b(src, f32b_idx32(q.n, q.g, y + a, x + b, C8, hs, ws), v);
#pragma unroll
                for (int k = 0; k < 8; ++k) b4[k] += v[k];
              }
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, b4[k], acc[k]);
          }
        }
      }
    }
    const float nz = noise ? noise[((size_t)q.n * ho + q.y) * wo + q.x] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = q.g * 8 + k;
      float v = acc[k];
      if (noise) v = fmaf(nz, noise_w[ch], v);
      if (bias) v += bias[ch];
      acc[k] = v < 0.f ? v * slope : v;
    }
    store8_f32b(out, i, acc);
  }
}

// instance norm + style_mod (net.py:32-34, 154-156): y = (x-mean)*rstd*(style[n][0][c]+1) + style[n][1][c];
// the input may hold a single sample that is broadcast over the batch (first block: x = const);
// optional nearest x2 upsample of the result (upscale2d of the NEXT block, net.py:37-43, 143)
__global__ void k_instance_norm_style(const float* __restrict__ x, int in_n, const float* __restrict__ mr,
                                      const float* __restrict__ style, int up, void* __restrict__ out_act,
                                      float* __restrict__ out_f32b, int n, int c, int h, int w, int planes) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, h, w);
    const int sn = in_n == 1 ? 0 : q.n;
    float v[8];
    load8_f32b(x, f32b_idx32(sn, q.g, q.y, q.x, C8, h, w), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = q.g * 8 + k;
      const size_t s = ((size_t)sn * c + ch) * 2;
      float t = (v[k] - __ldg(mr + s)) * __ldg(mr + s + 1);
      if (style) t = fmaf(t, __ldg(style + (size_t)q.n * 2 * c + ch) + 1.f, __ldg(style + (size_t)q.n * 2 * c + c + ch));
      v[k] = t;
    }
    for (int dy = 0; dy < up; ++dy)
      for (int dx = 0; dx < up; ++dx) {
        if (out_f32b) store8_f32b(out_f32b, f32b_idx32(q.n, q.g, q.y * up + dy, q.x * up + dx, C8, h * up, w * up), v);
        if (out_act) store8_act(out_act, q.n, q.g, q.y * up + dy, q.x * up + dx, C8, planes, h * up, w * up, v);
      }
  }
}

// 1x1 conv F32B -> NCHW (ToRGB, net.py:244-253); one thread per pixel
__global__ void k_to_rgb_f32b(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                              float* __restrict__ out, int n, int c, int nch, int h, int w) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int b = (int)(i / ((size_t)w * h));
    for (int ch = 0; ch < nch; ++ch) {
      float s = bias ? bias[ch] : 0.f;
      for (int g = 0; g < C8; ++g) {
        float v[8];
        load8_f32b(x, f32b_idx32(b, g, y, xx, C8, h, w), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) s = fmaf(v[k], __ldg(wt + (size_t)ch * c + g * 8 + k), s);
      }
      out[(((size_t)b * nch + ch) * h + y) * w + xx] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// BigGAN pieces
// ---------------------------------------------------------------------------------------------
__global__ void k_cbn_coeffs(const float* __restrict__ scale, const float* __restrict__ offset,
                             const float* __restrict__ weight, const float* __restrict__ bias,
                             const float* __restrict__ mean, const float* __restrict__ var, float eps,
                             float* __restrict__ a_out, float* __restrict__ b_out, int n, int c) {
  const int total = n * c;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ch = i % c;
    const float inv = 1.f / sqrtf(var[ch] + eps);
    const float w = scale ? 1.f + scale[i] : weight[ch];
    const float o = offset ? offset[i] : bias[ch];
    const float a = w * inv;
    a_out[i] = a;
    b_out[i] = o - mean[ch] * a;
  }
}

// grid = (splits, n * C/8): a block owns one (sample, 8-channel group), keeps its 16 coefficients in registers and walks its
// pixels with 32-bit indices (the flat form spent most of its instructions on three 64-bit divisions, 16 coefficient loads
// and the output index products per element and stayed at 74 % of the copy bandwidth).
__global__ void __launch_bounds__(256)
k_affine_act(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b, int relu, int up,
             void* __restrict__ out_act, float* __restrict__ out_f32b, int groups, int c, int h, int w, int planes) {
  const int C8 = c >> 3;
  for (int ng = blockIdx.y; ng < groups; ng += gridDim.y) {      // (grid.y is capped at 65535)
  const int nidx = ng / C8, grp = ng - nidx * C8;
  float av[8], bv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    av[k] = __ldg(a + (size_t)nidx * c + grp * 8 + k);
    bv[k] = __ldg(b + (size_t)nidx * c + grp * 8 + k);
  }
  const unsigned hw = (unsigned)(h * w), uw = (unsigned)w, ow = (unsigned)(w * up), ohw = hw * (unsigned)(up * up);
  const float* xb = x + (size_t)ng * hw * 8;
  float* fb = out_f32b ? out_f32b + (size_t)ng * ohw * 8 : nullptr;
  uint4* ab = out_act ? reinterpret_cast<uint4*>(out_act) + (size_t)ng * planes * ohw : nullptr;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += stride) {
    float v[8];
    load8_f32b(xb, i, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t = fmaf(v[k], av[k], bv[k]);
      v[k] = (relu && t < 0.f) ? 0.f : t;
    }
    if (up == 1) {
      if (fb) store8_f32b(fb, i, v);
      if (ab) {
        uint4 hi, lo;
        split8(v, hi, lo);
        ab[i] = hi;
        if (planes == 2) ab[i + ohw] = lo;
      }
    } else {
      const unsigned y = i / uw, xx = i - y * uw;
      uint4 hi, lo;
      if (ab) split8(v, hi, lo);
      for (int dy = 0; dy < up; ++dy)
        for (int dx = 0; dx < up; ++dx) {
          const unsigned o = (y * up + dy) * ow + xx * up + dx;
          if (fb) store8_f32b(fb, o, v);
          if (ab) {
            ab[o] = hi;
            if (planes == 2) ab[o + ohw] = lo;
          }
        }
    }
  }
  }
}

__global__ void k_maxpool2_f32b(const float* __restrict__ x, float* __restrict__ out, int n, int c, int ho, int wo) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * C8 * ho * wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const Idx4 q = decode4(i, C8, ho, wo);
    float m[8], t[8];
    load8_f32b(x, f32b_idx32(q.n, q.g, 2 * q.y, 2 * q.x, C8, 2 * ho, 2 * wo), m);
#pragma unroll
    for (int d = 1; d < 4; ++d) {
      load8_f32b(x, f32b_idx32(q.n, q.g, 2 * q.y + (d >> 1), 2 * q.x + (d & 1), C8, 2 * ho, 2 * wo), t);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], t[k]);
    }
    store8_f32b(out, i, m);
  }
}

// one thread per pixel; three passes over the channel groups (max, sum, write) -- the re-reads hit L1/L2
// one WARP per pixel: lane l holds channel groups l, l+32, ... (<= 8 of them, c <= 2048) in registers, so the tensor is
// read once and the max / sum are warp shuffles.  (The one-thread-per-pixel form below has only h*w threads per
// sample -- 4096 for BigGAN's 64x64 attention map -- and reads the tensor three times.)
__global__ void __launch_bounds__(256)
k_channel_softmax_warp(const float* __restrict__ x, void* __restrict__ out, int n, int c, int h, int w, int planes) {
  const int C8 = c >> 3;
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)n * h * w;
  const size_t warp0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t i = warp0; i < total; i += nwarps) {
    const int xx = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int b = (int)(i / ((size_t)w * h));
    float v[8][8];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = lane + 32 * j;
      if (g < C8) {
        load8_f32b(x, f32b_idx32(b, g, y, xx, C8, h, w), v[j]);
#pragma unroll
        for (int k = 0; k < 8; ++k) mx = fmaxf(mx, v[j][k]);
      }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (lane + 32 * j < C8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[j][k] = expf(v[j][k] - mx);
          sum += v[j][k];
        }
      }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = lane + 32 * j;
      if (g < C8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[j][k] *= inv;
        store8_act(out, b, g, y, xx, C8, planes, h, w, v[j]);
      }
    }
  }
}

__global__ void k_channel_softmax_to_act(const float* __restrict__ x, void* __restrict__ out, int n, int c, int h, int w,
                                         int planes) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int b = (int)(i / ((size_t)w * h));
    float mx = -INFINITY;
    for (int g = 0; g < C8; ++g) {
      float v[8];
      load8_f32b(x, f32b_idx32(b, g, y, xx, C8, h, w), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) mx = fmaxf(mx, v[k]);
    }
    float sum = 0.f;
    for (int g = 0; g < C8; ++g) {
      float v[8];
      load8_f32b(x, f32b_idx32(b, g, y, xx, C8, h, w), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += expf(v[k] - mx);
    }
    const float inv = 1.f / sum;
    for (int g = 0; g < C8; ++g) {
      float v[8];
      load8_f32b(x, f32b_idx32(b, g, y, xx, C8, h, w), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = expf(v[k] - mx) * inv;
      store8_act(out, b, g, y, xx, C8, planes, h, w, v);
    }
  }
}

__global__ void k_tanh_slice_nchw(const float* __restrict__ x, float* __restrict__ out, int n, int c, int nch, int hw) {
  const size_t total = (size_t)n * nch * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i % hw;
    const int ch = (int)((i / hw) % nch);
    const size_t b = i / ((size_t)hw * nch);
    out[i] = tanhf(x[(b * c + ch) * hw + pix]);
  }
}

// ---------------------------------------------------------------------------------------------
// PGGAN: pixel-norm over channels (one thread per pixel; the second pass over the channel groups hits L1/L2)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float pixel_rnorm(const float* x, int n, int y, int xx, int C8, int h, int w, float eps) {
  float ss = 0.f;
  for (int g = 0; g < C8; ++g) {
    float v[8];
    load8_f32b(x, f32b_idx32(n, g, y, xx, C8, h, w), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) ss = fmaf(v[k], v[k], ss);
  }
  return 1.f / sqrtf(ss / (float)(C8 * 8) + eps);
}

__global__ void k_pixelnorm_to_act(const float* __restrict__ x, void* __restrict__ out, int n, int c, int h, int w,
                                   int up, float eps, int planes) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int b = (int)(i / ((size_t)w * h));
    const float rn = pixel_rnorm(x, b, y, xx, C8, h, w, eps);
    for (int g = 0; g < C8; ++g) {
      float v[8];
      load8_f32b(x, f32b_idx32(b, g, y, xx, C8, h, w), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= rn;
      for (int dy = 0; dy < up; ++dy)
        for (int dx = 0; dx < up; ++dx)
          store8_act(out, b, g, y * up + dy, xx * up + dx, C8, planes, h * up, w * up, v);
    }
  }
}

__global__ void k_pixelnorm_to_rgb(const float* __restrict__ x, const float* __restrict__ wt,
                                   const float* __restrict__ bias, float* __restrict__ out, int n, int c, int nch, int h,
                                   int w, float eps) {
  const int C8 = c >> 3;
  const size_t total = (size_t)n * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int b = (int)(i / ((size_t)w * h));
    const float rn = pixel_rnorm(x, b, y, xx, C8, h, w, eps);
    for (int ch = 0; ch < nch; ++ch) {
      float s = 0.f;
      for (int g = 0; g < C8; ++g) {
        float v[8];
        load8_f32b(x, f32b_idx32(b, g, y, xx, C8, h, w), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) s = fmaf(v[k] * rn, __ldg(wt + (size_t)ch * c + g * 8 + k), s);
      }
      out[(((size_t)b * nch + ch) * h + y) * w + xx] = s + (bias ? bias[ch] : 0.f);
    }
  }
}

__global__ void k_upsample_nearest_nchw(const float* __restrict__ x, float* __restrict__ out, size_t planes, int h,
                                        int w) {
  const size_t total = planes * 4 * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % (2 * w));
    const int yo = (int)((i / (2 * w)) % (2 * h));
    const size_t pl = i / ((size_t)4 * h * w);
    out[i] = x[(pl * h + (yo >> 1)) * w + (xo >> 1)];
  }
}

__global__ void k_axpby(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, float a,
                        float b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = a * x[i] + b * y[i];
}

// ---------------------------------------------------------------------------------------------
// image / latent losses (training_utils.py:54-99, metric/pytorch_ssim.py:18-38): one-pass reductions.
// Every kernel block-reduces in fp64 and does one atomicAdd(double) per block and output.
// ---------------------------------------------------------------------------------------------
template <int NOUT>
__device__ __forceinline__ void block_reduce_add(double* vals, double* out) {
  __shared__ double red[NOUT][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NOUT; ++k) {
#pragma unroll
    for (int off = 16; off; off >>= 1) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], off);
    if (lane == 0) red[k][warp] = vals[k];
  }
  __syncthreads();
  if (threadIdx.x < NOUT) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
    atomicAdd(&out[threadIdx.x], t);
  }
}

// out[0..5] += sum a, sum b, sum a^2, sum b^2, sum a*b, sum (a-b)^2
__global__ void __launch_bounds__(256) k_pair_moments(const float* __restrict__ a, const float* __restrict__ b,
                                                      size_t n, double* __restrict__ out) {
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double x = a[i], y = b[i], d = x - y;
    v[0] += x; v[1] += y; v[2] += x * x; v[3] += y * y; v[4] += x * y; v[5] += d * d;
  }
  block_reduce_add<6>(v, out);
}

// KLDivLoss(log softmax(b), softmax(a)) summed: softmax over a dimension of size D and stride `inner`
// (torch's implicit-dim rule picks dim 1 for 4-D and dim 0 for 3-D inputs, training_utils.py:68-69)
__global__ void __launch_bounds__(256) k_softmax_kl(const float* __restrict__ a, const float* __restrict__ b,
                                                    size_t outer, int D, size_t inner, double* __restrict__ out) {
  double v[1] = {0};
  const size_t total = outer * inner;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = i / inner, in = i % inner;
    const float* pa = a + o * D * inner + in;
    const float* pb = b + o * D * inner + in;
    float ma = -INFINITY, mb = -INFINITY;
    for (int d = 0; d < D; ++d) {
      ma = fmaxf(ma, pa[d * inner]);
      mb = fmaxf(mb, pb[d * inner]);
    }
    float sa = 0.f, sb = 0.f;
    for (int d = 0; d < D; ++d) {
      sa += expf(pa[d * inner] - ma);
      sb += expf(pb[d * inner] - mb);
    }
    const float lsa = logf(sa), lsb = logf(sb);
    for (int d = 0; d < D; ++d) {
      const float la = pa[d * inner] - ma - lsa;   // log softmax(a)
      const float t = expf(pa[d * inner] - ma) / sa;  // softmax(a), as the reference computes it
      const float lq = logf(expf(pb[d * inner] - mb) / sb);  // log(softmax(b)) -- reference takes log of the softmax
      (void)lsb;
      // xlogy convention of kl_div: 0 where target == 0
      if (t > 0.f) v[0] += (double)(t * (logf(t) - lq));
      (void)la;
    }
  }
  block_reduce_add<1>(v, out);
}

// f x f average pooling of an NCHW tensor (repeated F.avg_pool2d(.,2,2), training_utils.py:81-84)
__global__ void k_avgpool_nchw(const float* __restrict__ x, float* __restrict__ out, size_t planes, int ho, int wo,
                               int f) {
  const size_t total = planes * ho * wo;
  const int wi = wo * f, hi = ho * f;
  const float inv = 1.f / (float)(f * f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % wo);
    const int yo = (int)((i / wo) % ho);
    const size_t pl = i / ((size_t)wo * ho);
    const float* src = x + pl * hi * wi + (size_t)yo * f * wi + (size_t)xo * f;
    float s = 0.f;
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) s += src[(size_t)dy * wi + dx];
    out[i] = s * inv;
  }
}

// SSIM map mean: 11x11 Gaussian (sigma 1.5) depthwise, zero padding 5, C1 = 1e-4, C2 = 9e-4; out[0] += sum of map
__constant__ float c_gauss11[11];
__global__ void __launch_bounds__(256) k_ssim_sum(const float* __restrict__ a, const float* __restrict__ b,
                                                  size_t planes, int h, int w, double* __restrict__ out) {
  double v[1] = {0};
  const size_t total = planes * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const size_t pl = i / ((size_t)w * h);
    const float* pa = a + pl * h * w;
    const float* pb = b + pl * h * w;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int dy = -5; dy <= 5; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      const float gy = c_gauss11[dy + 5];
      for (int dx = -5; dx <= 5; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= w) continue;
        const float g = gy * c_gauss11[dx + 5];
        const float va = pa[(size_t)yy * w + xx], vb = pb[(size_t)yy * w + xx];
        m1 = fmaf(g, va, m1);
        m2 = fmaf(g, vb, m2);
        s11 = fmaf(g, va * va, s11);
        s22 = fmaf(g, vb * vb, s22);
        s12 = fmaf(g, va * vb, s12);
      }
    }
    const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu12 = m1 * m2;
    const float sg1 = s11 - mu1_sq, sg2 = s22 - mu2_sq, sg12 = s12 - mu12;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    v[0] += (double)(((2.f * mu12 + C1) * (2.f * sg12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sg1 + sg2 + C2)));
  }
  block_reduce_add<1>(v, out);
}

// SSIM backward w.r.t. the second image (the training scripts back-propagate 1 - ssim(imgs1, imgs2) into imgs2,
// training_utils.py:87-88, E_align_s2.py:184-205).  With m = blur(b), Ebb = blur(b^2), Eab = blur(a b) (gaussian, zero pad):
//   pass 1: per pixel the partials of the SSIM map S w.r.t. those three blurred quantities -> G[0..2]
//   pass 2: d b[q] = go/numel * ( blur(G0)[q] + 2 b[q] blur(G1)[q] + a[q] blur(G2)[q] )      (the blur is self-adjoint)
__global__ void __launch_bounds__(256) k_ssim_grad_maps(const float* __restrict__ a, const float* __restrict__ b,
                                                        size_t planes, int h, int w, float* __restrict__ G) {
  const size_t total = planes * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const size_t pl = i / ((size_t)w * h);
    const float* pa = a + pl * h * w;
    const float* pb = b + pl * h * w;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int dy = -5; dy <= 5; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      const float gy = c_gauss11[dy + 5];
      for (int dx = -5; dx <= 5; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= w) continue;
        const float g = gy * c_gauss11[dx + 5];
        const float va = pa[(size_t)yy * w + xx], vb = pb[(size_t)yy * w + xx];
        m1 = fmaf(g, va, m1);
        m2 = fmaf(g, vb, m2);
        s11 = fmaf(g, va * va, s11);
        s22 = fmaf(g, vb * vb, s22);
        s12 = fmaf(g, va * vb, s12);
      }
    }
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float n1 = 2.f * m1 * m2 + C1, n2 = 2.f * (s12 - m1 * m2) + C2;
    const float d1 = m1 * m1 + m2 * m2 + C1, d2 = (s11 - m1 * m1) + (s22 - m2 * m2) + C2;
    const float inv = 1.f / (d1 * d2), S = n1 * n2 * inv;
    // dS/dm2: through n1 (2 m1), n2 (-2 m1), d1 (2 m2), d2 (-2 m2)
    G[i] = 2.f * m1 * (n2 - n1) * inv - 2.f * m2 * S / d1 + 2.f * m2 * S / d2;
    G[total + i] = -S / d2;                 // dS/dEbb
    G[2 * total + i] = 2.f * n1 * inv;      // dS/dEab
  }
}

__global__ void __launch_bounds__(256) k_ssim_grad_apply(const float* __restrict__ a, const float* __restrict__ b,
                                                         const float* __restrict__ G, const float* __restrict__ go,
                                                         size_t planes, int h, int w, float* __restrict__ db) {
  const size_t total = planes * h * w;
  const float scale = __ldg(go) / (float)total;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const size_t pl = i / ((size_t)w * h);
    const float* g0 = G + pl * h * w;
    const float* g1 = g0 + total;
    const float* g2 = g1 + total;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    for (int dy = -5; dy <= 5; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      const float gy = c_gauss11[dy + 5];
      for (int dx = -5; dx <= 5; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= w) continue;
        const float g = gy * c_gauss11[dx + 5];
        const size_t o = (size_t)yy * w + xx;
        b0 = fmaf(g, g0[o], b0);
        b1 = fmaf(g, g1[o], b1);
        b2 = fmaf(g, g2[o], b2);
      }
    }
    db[i] = scale * (b0 + 2.f * b[i] * b1 + a[i] * b2);
  }
}

// ---------------------------------------------------------------------------------------------
// Grad-CAM / Grad-CAM++ maps (metric/grad_cam.py:101-194) from the hooked feature / gradient tensors (NCHW fp32).
// The reference does this per image on the host in NumPy (float64 for the ++ variant) + cv2.resize.
// ---------------------------------------------------------------------------------------------
// first-max argmax per row (np.argmax) -> idx[n]; then mode with smallest-index tie break (np.argmax(np.bincount))
__global__ void k_argmax_rows(const float* __restrict__ logits, int n, int k, long long* __restrict__ idx) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float* p = logits + (size_t)r * k;
  float best = p[0];
  int bi = 0;
  for (int i = 1; i < k; ++i)
    if (p[i] > best) { best = p[i]; bi = i; }
  idx[r] = bi;
}
__global__ void k_mode(const long long* __restrict__ idx, int n, long long* __restrict__ mode) {
  if (blockIdx.x || threadIdx.x) return;
  long long best = -1;
  int bc = 0;
  for (int i = 0; i < n; ++i) {
    int c = 0;
    for (int j = 0; j < n; ++j) c += idx[j] == idx[i];
    if (c > bc || (c == bc && idx[i] < best)) { bc = c; best = idx[i]; }
  }
  *mode = best;
}

// one block per (n, c): plus == 1: w = sum(relu(g) * (1/sum relu(g)))  (0 if the sum is 0)   [Grad-CAM++ as coded, :170-178]
//                       plus == 0: w = mean(g)                                               [Grad-CAM, :117]
__global__ void __launch_bounds__(128) k_gradcam_weights(const float* __restrict__ grad, int hw, int plus,
                                                         double* __restrict__ wout) {
  const float* g = grad + (size_t)blockIdx.x * hw;
  __shared__ float sred[4];
  __shared__ double dred[4];
  float s = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) s += plus ? fmaxf(g[i], 0.f) : g[i];
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  const float tot = sred[0] + sred[1] + sred[2] + sred[3];
  if (!plus) {
    if (threadIdx.x == 0) wout[blockIdx.x] = (double)(tot / (float)hw);     // np.mean of float32 stays float32
    return;
  }
  const float inv = tot > 0.f ? 1.f / tot : 0.f;                            // float32 array element (:173-174)
  double d = 0.0;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) d += (double)fmaxf(g[i], 0.f) * (double)inv;
  for (int off = 16; off; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
  if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) wout[blockIdx.x] = dred[0] + dred[1] + dred[2] + dred[3];
}

// one block per image: cam = sum_c feat_c * w_c  [relu if !plus]; cam -= min; cam /= max   (:179-186 / :118-125)
__global__ void __launch_bounds__(256) k_gradcam_map(const float* __restrict__ feat, const double* __restrict__ w,
                                                     int c, int hw, int plus, double* __restrict__ cam) {
  const int n = blockIdx.x;
  __shared__ double smin[8], smax[8];
  double lmin = INFINITY, lmax = -INFINITY;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    double v;
    if (plus) {
      double a = 0.0;
      for (int ch = 0; ch < c; ++ch) a += (double)feat[((size_t)n * c + ch) * hw + i] * w[(size_t)n * c + ch];
      v = a;
    } else {
      float a = 0.f;                                                       // float32 path of the base class
      for (int ch = 0; ch < c; ++ch) a += feat[((size_t)n * c + ch) * hw + i] * (float)w[(size_t)n * c + ch];
      v = (double)fmaxf(a, 0.f);
    }
    cam[(size_t)n * hw + i] = v;
    lmin = fmin(lmin, v);
    lmax = fmax(lmax, v);
  }
  for (int off = 16; off; off >>= 1) {
    lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, off));
    lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lmin; smax[threadIdx.x >> 5] = lmax; }
  __syncthreads();
  double mn = smin[0], mx = smax[0];
  for (int i = 1; i < 8; ++i) { mn = fmin(mn, smin[i]); mx = fmax(mx, smax[i]); }
  const double denom = plus ? (mx - mn) : (double)((float)mx - (float)mn);
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const size_t o = (size_t)n * hw + i;
    cam[o] = plus ? (cam[o] - mn) / denom : (double)(((float)cam[o] - (float)mn) / (float)denom);
  }
}

// cv2.resize(src, (wo, ho)) INTER_LINEAR for a single-channel float image: half-pixel centres, float coefficients,
// edge clamping (OpenCV resizeGeneric / HResizeLinear / VResizeLinear)
__global__ void k_resize_bilinear(const double* __restrict__ src, int n, int hi, int wi, int ho, int wo, int f32path,
                                  double* __restrict__ dst) {
  const size_t total = (size_t)n * ho * wo;
  const double sx_scale = (double)wi / wo, sy_scale = (double)hi / ho;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int dx = (int)(i % wo), dy = (int)((i / wo) % ho);
    const size_t b = i / ((size_t)wo * ho);
    float fx = (float)((dx + 0.5) * sx_scale - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= wi - 1) { fx = 0.f; sx = wi - 1; }
    float fy = (float)((dy + 0.5) * sy_scale - 0.5);
    int sy = (int)floorf(fy);
    fy -= sy;
    if (sy < 0) { fy = 0.f; sy = 0; }
    if (sy >= hi - 1) { fy = 0.f; sy = hi - 1; }
    const int sx1 = sx + 1 < wi ? sx + 1 : sx, sy1 = sy + 1 < hi ? sy + 1 : sy;
    const double* p = src + b * (size_t)hi * wi;
    if (f32path) {
      const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
      const float r0 = (float)p[(size_t)sy * wi + sx] * a0 + (float)p[(size_t)sy * wi + sx1] * a1;
      const float r1 = (float)p[(size_t)sy1 * wi + sx] * a0 + (float)p[(size_t)sy1 * wi + sx1] * a1;
      dst[i] = (double)(r0 * b0 + r1 * b1);
    } else {
      const double a0 = (double)(1.f - fx), a1 = (double)fx, b0 = (double)(1.f - fy), b1 = (double)fy;
      const double r0 = p[(size_t)sy * wi + sx] * a0 + p[(size_t)sy * wi + sx1] * a1;
      const double r1 = p[(size_t)sy1 * wi + sx] * a0 + p[(size_t)sy1 * wi + sx1] * a1;
      dst[i] = r0 * b0 + r1 * b1;
    }
  }
}

// mask2cam (metric/grad_cam.py:234-251): JET colour map of the mask + overlay on the image
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
// heat[n][c][p] = lut_rgb[(uint8)(255*mask[n][p])][c] / 255 ; cam = heat + img
__global__ void k_jet_overlay(const double* __restrict__ mask, const float* __restrict__ img,
                              const float* __restrict__ lut_rgb, float* __restrict__ heat, float* __restrict__ cam,
                              int n, int hw) {
  const size_t total = (size_t)n * 3 * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i % hw;
    const int c = (int)((i / hw) % 3);
    const size_t b = i / ((size_t)3 * hw);
    const unsigned char q = (unsigned char)(long long)(255.0 * mask[b * hw + p]);   // np.uint8(255 * j): truncation
    const float hv = lut_rgb[q * 3 + c] / 255.f;
    heat[i] = hv;
    cam[i] = img[i];          // cam starts as a copy of the images (:239); image i gets its overlay inside the loop
  }
}
// out[0] = min(x[0:n]), out[1] = max(x[0:n]); out must be pre-set to (+inf, -inf)
__global__ void k_minmax(const float* __restrict__ x, size_t n, float* __restrict__ out) {
  float mn = INFINITY, mx = -INFINITY;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    mn = fminf(mn, x[i]);
    mx = fmaxf(mx, x[i]);
  }
  for (int off = 16; off; off >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_f(out, mn);
    atomic_max_f(out + 1, mx);
  }
}
// x = (x - *sub) ; or x = x / *div   (device scalars)
__global__ void k_sub_or_div(float* __restrict__ x, size_t n, const float* __restrict__ sub,
                             const float* __restrict__ div) {
  const float s = sub ? *sub : 0.f, d = div ? *div : 1.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = (x[i] - s) / d;
}

// ---------------------------------------------------------------------------------------------
// LREQAdam (model/utils/custom_adam.py:24-76): beta1 == 0 Adam, per-tensor step size
//   v = beta2*v + (1-beta2)*g*g ;  p -= step[t] * g / (sqrt(v) + eps)
// multi-tensor: block b works on tensor blk_tensor[b], elements [blk_off[b], blk_off[b] + chunk)
// ---------------------------------------------------------------------------------------------
__global__ void k_lreq_adam(float* const* __restrict__ params, const float* const* __restrict__ grads,
                            float* const* __restrict__ vs, const long long* __restrict__ numel,
                            const float* __restrict__ step, const int* __restrict__ blk_tensor,
                            const long long* __restrict__ blk_off, int chunk, float beta2, float eps) {
  const int t = blk_tensor[blockIdx.x];
  const long long off = blk_off[blockIdx.x];
  long long end = off + chunk;
  if (end > numel[t]) end = numel[t];
  float* p = params[t];
  const float* g = grads[t];
  float* v = vs[t];
  const float st = step[t], omb = 1.f - beta2;
  // 128-bit path when the three streams of this block are 16-byte aligned (gradients may be views into the flat
  // all-reduce bucket at any 4-byte offset): 5 streams of 4 B per element, HBM-bound
  if ((((uintptr_t)(p + off) | (uintptr_t)(g + off) | (uintptr_t)(v + off)) & 15) == 0) {
    const long long n4 = (end - off) >> 2;
    float4* p4 = reinterpret_cast<float4*>(p + off);
    const float4* g4 = reinterpret_cast<const float4*>(g + off);
    float4* v4 = reinterpret_cast<float4*>(v + off);
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 gq = g4[i];
      float4 vq = v4[i], pq = p4[i];
      vq.x = vq.x * beta2 + omb * gq.x * gq.x;
      vq.y = vq.y * beta2 + omb * gq.y * gq.y;
      vq.z = vq.z * beta2 + omb * gq.z * gq.z;
      vq.w = vq.w * beta2 + omb * gq.w * gq.w;
      pq.x = pq.x - st * (gq.x / (sqrtf(vq.x) + eps));
      pq.y = pq.y - st * (gq.y / (sqrtf(vq.y) + eps));
      pq.z = pq.z - st * (gq.z / (sqrtf(vq.z) + eps));
      pq.w = pq.w - st * (gq.w / (sqrtf(vq.w) + eps));
      v4[i] = vq;
      p4[i] = pq;
    }
    for (long long i = off + (n4 << 2) + threadIdx.x; i < end; i += blockDim.x) {
      const float gi = g[i];
      const float vi = v[i] * beta2 + omb * gi * gi;
      v[i] = vi;
      p[i] = p[i] - st * (gi / (sqrtf(vi) + eps));
    }
    return;
  }
  for (long long i = off + threadIdx.x; i < end; i += blockDim.x) {
    const float gi = g[i];
    const float vi = v[i] * beta2 + omb * gi * gi;
    v[i] = vi;
    p[i] = p[i] - st * (gi / (sqrtf(vi) + eps));
  }
}

}  // namespace dge

// =================================================================================================
// C ABI
// =================================================================================================
using namespace dge;

extern "C" {

const char* dge_last_error(void) { return g_err; }
int dge_version(void) { return 100; }
int64_t dge_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void dge_launch_count_reset(void) { g_launches.store(0, std::memory_order_relaxed); }

int dge_device_ok(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return DGE_ERR_CUDA;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return DGE_ERR_CUDA;
  }
  if (prop.major != 10) {
    set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    return DGE_ERR_UNSUPPORTED;
  }
  return DGE_OK;
}

int dge_pack_conv_weight(const float* w, void* wpk, int cout, int cin, int ksize, int flip, float scale, int planes,
                         void* stream) {
  DGE_REQUIRE(w && wpk, "pack_conv_weight: null pointer");
  DGE_REQUIRE(cin % 16 == 0 && cout % 16 == 0 && cin > 0 && cout > 0, "pack_conv_weight: cin=%d cout=%d must be multiples of 16", cin, cout);
  DGE_REQUIRE(ksize == 1 || ksize == 3 || ksize == 4, "pack_conv_weight: ksize=%d", ksize);
  DGE_REQUIRE(planes == 1 || planes == 2, "pack_conv_weight: planes=%d", planes);
  LAUNCH_1D(k_pack_conv_weight, (size_t)ksize * ksize * (cin / 8) * cout, stream, w, (uint4*)wpk, cout, cin, ksize,
            flip, scale, planes, 0, cin);
}

int dge_pack_conv_weight_dgrad(const float* w, void* wpk, int cout, int cin, int ksize, float scale, int planes,
                               void* stream) {
  DGE_REQUIRE(w && wpk, "pack_conv_weight_dgrad: null pointer");
  DGE_REQUIRE(cin % 16 == 0 && cout % 16 == 0 && cin > 0 && cout > 0,
              "pack_conv_weight_dgrad: cin=%d cout=%d must be multiples of 16", cin, cout);
  DGE_REQUIRE(ksize == 1 || ksize == 3, "pack_conv_weight_dgrad: ksize=%d", ksize);
  DGE_REQUIRE(planes == 1 || planes == 2, "pack_conv_weight_dgrad: planes=%d", planes);
  // packed operand: cin output columns, contraction over cout, spatially flipped
  LAUNCH_1D(k_pack_conv_weight, (size_t)ksize * ksize * (cout / 8) * cin, stream, w, (uint4*)wpk, cin, cout, ksize, 1,
            scale, planes, 1, cin);
}

int dge_weight_sqsum(const float* w, float* w2, int cout, int cin, int ksize, float scale, void* stream) {
  DGE_REQUIRE(w && w2 && cout > 0 && cin > 0 && ksize > 0, "weight_sqsum: bad args");
  LAUNCH_1D(k_weight_sqsum, (size_t)cout * cin, stream, w, w2, cout, cin, ksize * ksize, scale);
}

int dge_demod(const float* w2, const float* style, float* d, int n, int cout, int cin, float eps, void* stream) {
  DGE_REQUIRE(w2 && style && d && n > 0 && cout > 0 && cin > 0, "demod: bad args");
  LAUNCH_1D(k_demod, (size_t)n * cout * 32, stream, w2, style, d, n, cout, cin, eps);
}

int dge_rgb_weights(const float* w, const float* style, float* rgb_w, int n, int nch, int cin, float scale,
                    void* stream) {
  DGE_REQUIRE(w && style && rgb_w && n > 0 && nch > 0 && cin > 0, "rgb_weights: bad args");
  LAUNCH_1D(k_rgb_weights, (size_t)n * nch * cin, stream, w, style, rgb_w, n, nch, cin, scale);
}

int dge_dense(const float* x, const float* w, const float* b, float* y, int n, int k, int m, float wscale,
              float bscale, float add_bias, float slope, float gain, void* stream) {
  DGE_REQUIRE(x && w && y && n > 0 && k > 0 && m > 0, "dense: bad args");
  LAUNCH_1D(k_dense, (size_t)n * m * 32, stream, x, w, b, y, n, k, m, wscale, bscale, add_bias, slope, gain);
}

int dge_sg2_prep(const dge_sg2_prep_item* items, int n_items, const float* wp, float* arena, int n, int num_layers,
                 int wdim, void* stream) {
  DGE_REQUIRE(items && wp && arena && n_items > 0 && n > 0 && num_layers > 0 && wdim > 0 && wdim % 4 == 0 && wdim <= 2048,
              "sg2_prep: bad args (w dimension must be a multiple of 4, <= 2048; channel counts multiples of 4, <= 2048)");
  dim3 grid(n_items, n);
  k_sg2_prep<<<grid, 256, 0, (cudaStream_t)stream>>>(items, wp, arena, num_layers, wdim);
  count_launch();
  return check_launch("k_sg2_prep");
}

int dge_sg2_prep_bwd(const dge_sg2_prep_item* items, int n_items, const float* arena, const int64_t* gsrc,
                     const float* sums, float* d_wp, int n, int num_layers, int wdim, void* stream) {
  DGE_REQUIRE(items && arena && gsrc && sums && d_wp && n_items > 0 && n > 0 && num_layers > 0 && wdim > 0,
              "sg2_prep_bwd: bad args");
  cudaError_t e = cudaMemsetAsync(d_wp, 0, (size_t)n * num_layers * wdim * sizeof(float), (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("sg2_prep_bwd: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    return DGE_ERR_CUDA;
  }
  dim3 grid(n_items, n);
  k_sg2_prep_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(items, arena, (const long long*)gsrc, sums, d_wp, num_layers, wdim);
  count_launch();
  return check_launch("k_sg2_prep_bwd");
}

int dge_pixel_norm(const float* x, float* y, int n, int k, float eps, void* stream) {
  DGE_REQUIRE(x && y && n > 0 && k > 0, "pixel_norm: bad args");
  LAUNCH_1D(k_pixel_norm, (size_t)n * 32, stream, x, y, n, k, eps);
}

#define REQ_NCHW(name)                                                                                      \
  DGE_REQUIRE(n > 0 && c > 0 && c % 8 == 0 && h > 0 && w > 0, name ": bad dims n=%d c=%d h=%d w=%d (c %% 8)", n, c, h, w)

int dge_nchw_to_act(const float* x, int64_t x_bstride, const float* scale, void* act, int n, int c, int h, int w,
                    int planes, void* stream) {
  DGE_REQUIRE(x && act, "nchw_to_act: null pointer");
  REQ_NCHW("nchw_to_act");
  DGE_REQUIRE(planes == 1 || planes == 2, "nchw_to_act: planes=%d", planes);
  LAUNCH_1D(k_nchw_to_act, (size_t)n * (c / 8) * h * w, stream, x, (long long)x_bstride, scale, act, n, c, h, w, planes);
}
int dge_nchw_to_f32b(const float* x, float* out, int n, int c, int h, int w, void* stream) {
  DGE_REQUIRE(x && out, "nchw_to_f32b: null pointer");
  REQ_NCHW("nchw_to_f32b");
  LAUNCH_1D(k_nchw_to_f32b, (size_t)n * (c / 8) * h * w, stream, x, out, n, c, h, w);
}
int dge_f32b_to_nchw(const float* x, float* out, int n, int c, int h, int w, void* stream) {
  DGE_REQUIRE(x && out, "f32b_to_nchw: null pointer");
  REQ_NCHW("f32b_to_nchw");
  LAUNCH_1D(k_f32b_to_nchw, (size_t)n * (c / 8) * h * w, stream, x, out, n, c, h, w);
}
int dge_act_to_nchw(const void* act, float* out, int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(act && out, "act_to_nchw: null pointer");
  REQ_NCHW("act_to_nchw");
  DGE_REQUIRE(planes == 1 || planes == 2, "act_to_nchw: planes=%d", planes);
  LAUNCH_1D(k_act_to_nchw, (size_t)n * (c / 8) * h * w, stream, act, out, n, c, h, w, planes);
}

int dge_f32b_to_act(const float* x, void* act, int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(x && act, "f32b_to_act: null pointer");
  REQ_NCHW("f32b_to_act");
  DGE_REQUIRE(planes == 1 || planes == 2, "f32b_to_act: planes=%d", planes);
  LAUNCH_1D(k_f32b_to_act, (size_t)n * (c / 8) * h * w, stream, x, act, n, c, h, w, planes);
}
int dge_to_rgb_nchw(const float* x, const float* rgb_w, const float* bias, float* out, int n, int c, int nch, int h,
                    int w, void* stream) {
  DGE_REQUIRE(x && rgb_w && out && n > 0 && c > 0 && nch > 0 && h > 0 && w > 0, "to_rgb_nchw: bad args");
  LAUNCH_1D(k_to_rgb_nchw, (size_t)n * h * w, stream, x, rgb_w, bias, out, n, c, nch, h * w);
}

int dge_rgb_init(const float* img_in, const float* bias, float* img_out, int n, int nch, int h_out, int w_out,
                 void* stream) {
  DGE_REQUIRE(img_out && n > 0 && nch > 0 && h_out > 0 && w_out > 0, "rgb_init: bad args");
  DGE_REQUIRE(!img_in || (h_out % 2 == 0 && w_out % 2 == 0), "rgb_init: odd output size with an input image");
  if (img_in && w_out % 4 == 0 && (reinterpret_cast<uintptr_t>(img_out) & 15) == 0) {
    LAUNCH_1D(k_rgb_init_v4, (size_t)n * nch * (h_out / 2) * (w_out / 4), stream, img_in, bias, img_out, n, nch, h_out,
              w_out);
  }
  LAUNCH_1D(k_rgb_init, (size_t)n * nch * h_out * w_out, stream, img_in, bias, img_out, n, nch, h_out, w_out);
}

int dge_from_rgb(const float* img, const float* w, const float* b, float* out, int n, int cimg, int c, int h, int wd,
                 float slope, void* stream) {
  DGE_REQUIRE(img && w && out, "from_rgb: null pointer");
  DGE_REQUIRE(n > 0 && cimg > 0 && cimg <= 4 && c > 0 && c % 8 == 0 && c <= 2048 && h > 0 && wd > 0, "from_rgb: bad dims");
  k_from_rgb<<<grid_for((size_t)n * h * wd, 256), 256, (size_t)(5 * c) * sizeof(float), (cudaStream_t)stream>>>(
      img, w, b, out, n, cimg, c, h, wd, slope);
  count_launch();
  return check_launch("k_from_rgb");
}

int dge_instance_stats(const float* x, double* scratch, float* style, float* mean_rstd, int n, int c, int h, int w,
                       float eps, void* stream) {
  DGE_REQUIRE(x && scratch && (style || mean_rstd), "instance_stats: null pointer");
  REQ_NCHW("instance_stats");
  cudaStream_t st = (cudaStream_t)stream;
  const int hw = h * w;
  if (hw <= 4096) {
    k_instance_stats_small<<<n * (c / 8), 256, 0, st>>>(x, style, mean_rstd, c, hw, eps);
    count_launch();
    return check_launch("k_instance_stats_small");
  }
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * (size_t)n * c, st);
  if (e != cudaSuccess) {
    set_error("instance_stats memset: %s", cudaGetErrorString(e));
    return DGE_ERR_CUDA;
  }
  int splits = (hw + 256 * 16 - 1) / (256 * 16);
  if (splits < 1) splits = 1;
  if (splits > 256) splits = 256;
  dim3 grid(splits, n * (c / 8));
  k_instance_stats_partial<<<grid, 256, 0, st>>>(x, scratch, c, hw);
  count_launch();
  int r = check_launch("k_instance_stats_partial");
  if (r) return r;
  k_instance_stats_final<<<grid_for((size_t)n * c, 256), 256, 0, st>>>(scratch, style, mean_rstd, n, c, hw, eps);
  count_launch();
  return check_launch("k_instance_stats_final");
}

int dge_instance_norm(const float* x, const float* mean_rstd, void* out_act, float* out_f32b, int n, int c, int h,
                      int w, int planes, void* stream) {
  DGE_REQUIRE(x && mean_rstd && (out_act || out_f32b), "instance_norm: null pointer");
  REQ_NCHW("instance_norm");
  DGE_REQUIRE(!out_act || planes == 1 || planes == 2, "instance_norm: planes=%d", planes);
  LAUNCH_1D(k_instance_norm, (size_t)n * (c / 8) * h * w, stream, x, mean_rstd, (const float*)nullptr,
            (const float*)nullptr, out_act, out_f32b, n, c, h, w, planes);
}

int dge_instance_norm_affine(const float* x, const float* mean_rstd, const float* gamma, const float* beta,
                             void* out_act, float* out_f32b, int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(x && mean_rstd && (out_act || out_f32b), "instance_norm_affine: null pointer");
  DGE_REQUIRE(!gamma == !beta, "instance_norm_affine: gamma and beta must be given together");
  REQ_NCHW("instance_norm_affine");
  DGE_REQUIRE(!out_act || planes == 1 || planes == 2, "instance_norm_affine: planes=%d", planes);
  LAUNCH_1D(k_instance_norm, (size_t)n * (c / 8) * h * w, stream, x, mean_rstd, gamma, beta, out_act, out_f32b, n, c, h,
            w, planes);
}

int dge_instance_norm_blur(const float* x, const float* mean_rstd, void* out_act, int s2d, int n, int c, int h, int w,
                           int planes, void* stream) {
  DGE_REQUIRE(x && mean_rstd && out_act, "instance_norm_blur: null pointer");
  REQ_NCHW("instance_norm_blur");
  DGE_REQUIRE(planes == 1 || planes == 2, "instance_norm_blur: planes=%d", planes);
  DGE_REQUIRE(!s2d || (h % 2 == 0 && w % 2 == 0), "instance_norm_blur: space-to-depth needs even h, w");
  LAUNCH_1D(k_instance_norm_blur, (size_t)n * (c / 8) * h * w, stream, x, mean_rstd, out_act, s2d, n, c, h, w, planes);
}

int dge_instance_norm_pool(const float* x, const float* mean_rstd, void* out_act, void* out_pool_act, int n, int c, int h,
                           int w, int planes, void* stream) {
  DGE_REQUIRE(x && mean_rstd && out_act && out_pool_act, "instance_norm_pool: null pointer");
  REQ_NCHW("instance_norm_pool");
  DGE_REQUIRE(h % 2 == 0 && w % 2 == 0, "instance_norm_pool: odd input size %dx%d", h, w);
  DGE_REQUIRE(planes == 1 || planes == 2, "instance_norm_pool: planes=%d", planes);
  LAUNCH_1D(k_instance_norm_pool, (size_t)n * (c / 8) * (h / 2) * (w / 2), stream, x, mean_rstd, out_act, out_pool_act, n,
            c, h / 2, w / 2, planes);
}

int dge_from_rgb_stats(const float* img, const float* w, const float* b, float* out, double* scratch, float* style,
                       float* mean_rstd, int n, int cimg, int c, int h, int wd, float slope, float eps, void* stream) {
  DGE_REQUIRE(img && w && out && scratch && (style || mean_rstd), "from_rgb_stats: null pointer");
  DGE_REQUIRE(n > 0 && cimg > 0 && cimg <= 4 && (c == 16 || c == 32) && h > 0 && wd > 0 && slope >= 0.f && slope <= 1.f,
              "from_rgb_stats: bad dims (c must be 16 or 32; use dge_from_rgb + dge_instance_stats otherwise)");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * (size_t)n * c, st);
  if (e != cudaSuccess) {
    set_error("from_rgb_stats memset: %s", cudaGetErrorString(e));
    return DGE_ERR_CUDA;
  }
  const size_t hw = (size_t)h * wd;
  int bx = (int)((hw + 256 * 16 - 1) / (256 * 16));   // ~16 pixels per thread: the block reduction amortises
  if (bx < 1) bx = 1;
  if (bx > 2048) bx = 2048;
  dim3 grid(bx, n);
  if (c == 16)
    k_from_rgb_stats<2><<<grid, 256, 0, st>>>(img, w, b, out, scratch, cimg, h, wd, slope);
  else
    k_from_rgb_stats<4><<<grid, 256, 0, st>>>(img, w, b, out, scratch, cimg, h, wd, slope);
  count_launch();
  int r = check_launch("k_from_rgb_stats");
  if (r) return r;
  k_instance_stats_final<<<grid_for((size_t)n * c, 256), 256, 0, st>>>(scratch, style, mean_rstd, n, c, (int)hw, eps);
  count_launch();
  return check_launch("k_instance_stats_final");
}

int dge_avgpool_to_act(const float* x, void* out_act, int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(x && out_act, "avgpool_to_act: null pointer");
  REQ_NCHW("avgpool_to_act");
  DGE_REQUIRE(h % 2 == 0 && w % 2 == 0, "avgpool_to_act: odd input size %dx%d", h, w);
  DGE_REQUIRE(planes == 1 || planes == 2, "avgpool_to_act: planes=%d", planes);
  LAUNCH_1D(k_avgpool_to_act, (size_t)n * (c / 8) * (h / 2) * (w / 2), stream, x, out_act, n, c, h / 2, w / 2, planes);
}

int dge_blend(const float* a_src, const float* b_src, float* out, float a, float b, int pool, int n, int c, int h_out,
              int w_out, void* stream) {
  DGE_REQUIRE(a_src && b_src && out, "blend: null pointer");
  DGE_REQUIRE(n > 0 && c > 0 && c % 8 == 0 && h_out > 0 && w_out > 0, "blend: bad dims");
  LAUNCH_1D(k_blend, (size_t)n * (c / 8) * h_out * w_out, stream, a_src, b_src, out, a, b, pool, n, c, h_out, w_out);
}

int dge_cbn_coeffs(const float* scale, const float* offset, const float* weight, const float* bias, const float* mean,
                   const float* var, float eps, float* a_out, float* b_out, int n, int c, void* stream) {
  DGE_REQUIRE(mean && var && a_out && b_out && n > 0 && c > 0, "cbn_coeffs: bad args");
  DGE_REQUIRE((scale && offset) || (weight && bias), "cbn_coeffs: need (scale, offset) or (weight, bias)");
  LAUNCH_1D(k_cbn_coeffs, (size_t)n * c, stream, scale, offset, weight, bias, mean, var, eps, a_out, b_out, n, c);
}
int dge_affine_act(const float* x, const float* a, const float* b, int relu, int up, void* out_act, float* out_f32b,
                   int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(x && a && b && (out_act || out_f32b), "affine_act: null pointer");
  REQ_NCHW("affine_act");
  DGE_REQUIRE((up == 1 || up == 2) && (!out_act || planes == 1 || planes == 2), "affine_act: up=%d planes=%d", up, planes);
  DGE_REQUIRE((long long)h * w * up * up < (1ll << 31), "affine_act: map too large for 32-bit indexing (h=%d w=%d)", h, w);
  {
    const long long groups = (long long)n * (c / 8), pixels = (long long)h * w;
    long long splits = (148ll * 12 + groups - 1) / groups, cap = (pixels + 255) / 256;
    if (splits > cap) splits = cap;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    DGE_REQUIRE(groups < (1ll << 31), "affine_act: too many (sample, channel group) pairs (%lld)", groups);
    dim3 grid((unsigned)splits, (unsigned)(groups < 65535 ? groups : 65535));
    k_affine_act<<<grid, 256, 0, (cudaStream_t)stream>>>(x, a, b, relu, up, out_act, out_f32b, (int)groups, c, h, w, planes);
    count_launch();
    return check_launch("k_affine_act");
  }
}
int dge_maxpool2_f32b(const float* x, float* out, int n, int c, int h_out, int w_out, void* stream) {
  DGE_REQUIRE(x && out && n > 0 && c > 0 && c % 8 == 0 && h_out > 0 && w_out > 0, "maxpool2_f32b: bad args");
  LAUNCH_1D(k_maxpool2_f32b, (size_t)n * (c / 8) * h_out * w_out, stream, x, out, n, c, h_out, w_out);
}
int dge_channel_softmax_to_act(const float* x, void* out_act, int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(x && out_act, "channel_softmax_to_act: null pointer");
  REQ_NCHW("channel_softmax_to_act");
  DGE_REQUIRE(planes == 1 || planes == 2, "channel_softmax_to_act: planes=%d", planes);
  if (c <= 2048) LAUNCH_1D(k_channel_softmax_warp, (size_t)n * h * w * 32, stream, x, out_act, n, c, h, w, planes);
  LAUNCH_1D(k_channel_softmax_to_act, (size_t)n * h * w, stream, x, out_act, n, c, h, w, planes);
}
int dge_tanh_slice_nchw(const float* x, float* out, int n, int c, int nch, int hw, void* stream) {
  DGE_REQUIRE(x && out && n > 0 && c >= nch && nch > 0 && hw > 0, "tanh_slice_nchw: bad args");
  LAUNCH_1D(k_tanh_slice_nchw, (size_t)n * nch * hw, stream, x, out, n, c, nch, hw);
}

int dge_sg1_post(const float* src, int mode, const float* noise, const float* noise_w, const float* bias, float slope,
                 float* out_f32b, int n, int c, int h_out, int w_out, void* stream) {
  DGE_REQUIRE(src && out_f32b && mode >= 0 && mode <= 2, "sg1_post: bad args");
  DGE_REQUIRE(n > 0 && c > 0 && c % 8 == 0 && h_out > 0 && w_out > 0, "sg1_post: bad dims");
  DGE_REQUIRE(!noise == !noise_w, "sg1_post: noise and noise_w must be given together");
  LAUNCH_1D(k_sg1_post, (size_t)n * (c / 8) * h_out * w_out, stream, src, mode, noise, noise_w, bias, slope, out_f32b, n,
            c, h_out, w_out);
}
int dge_instance_norm_style(const float* x, int in_n, const float* mean_rstd, const float* style, int up, void* out_act,
                            float* out_f32b, int n, int c, int h, int w, int planes, void* stream) {
  DGE_REQUIRE(x && mean_rstd && (out_act || out_f32b), "instance_norm_style: null pointer");
  REQ_NCHW("instance_norm_style");
  DGE_REQUIRE((in_n == n || in_n == 1) && (up == 1 || up == 2), "instance_norm_style: in_n=%d up=%d", in_n, up);
  DGE_REQUIRE(!out_act || planes == 1 || planes == 2, "instance_norm_style: planes=%d", planes);
  LAUNCH_1D(k_instance_norm_style, (size_t)n * (c / 8) * h * w, stream, x, in_n, mean_rstd, style, up, out_act, out_f32b,
            n, c, h, w, planes);
}
int dge_to_rgb_f32b(const float* x, const float* w, const float* bias, float* out, int n, int c, int nch, int h, int wd,
                    void* stream) {
  DGE_REQUIRE(x && w && out && n > 0 && c > 0 && c % 8 == 0 && nch > 0 && h > 0 && wd > 0, "to_rgb_f32b: bad args");
  LAUNCH_1D(k_to_rgb_f32b, (size_t)n * h * wd, stream, x, w, bias, out, n, c, nch, h, wd);
}

int dge_pixelnorm_to_act(const float* x, void* out_act, int n, int c, int h, int w, int up, float eps, int planes,
                         void* stream) {
  DGE_REQUIRE(x && out_act, "pixelnorm_to_act: null pointer");
  REQ_NCHW("pixelnorm_to_act");
  DGE_REQUIRE((up == 1 || up == 2) && (planes == 1 || planes == 2), "pixelnorm_to_act: up=%d planes=%d", up, planes);
  LAUNCH_1D(k_pixelnorm_to_act, (size_t)n * h * w, stream, x, out_act, n, c, h, w, up, eps, planes);
}
int dge_pixelnorm_to_rgb(const float* x, const float* w, const float* bias, float* out, int n, int c, int nch, int h,
                         int wd, float eps, void* stream) {
  DGE_REQUIRE(x && w && out && n > 0 && c > 0 && c % 8 == 0 && nch > 0 && h > 0 && wd > 0, "pixelnorm_to_rgb: bad args");
  LAUNCH_1D(k_pixelnorm_to_rgb, (size_t)n * h * wd, stream, x, w, bias, out, n, c, nch, h, wd, eps);
}
int dge_upsample_nearest_nchw(const float* x, float* out, int64_t planes, int h, int w, void* stream) {
  DGE_REQUIRE(x && out && planes > 0 && h > 0 && w > 0, "upsample_nearest_nchw: bad args");
  LAUNCH_1D(k_upsample_nearest_nchw, (size_t)planes * 4 * h * w, stream, x, out, (size_t)planes, h, w);
}
int dge_axpby(const float* x, const float* y, float* out, float a, float b, int64_t n, void* stream) {
  DGE_REQUIRE(x && y && out && n > 0, "axpby: bad args");
  LAUNCH_1D(k_axpby, (size_t)n, stream, x, y, out, a, b, (size_t)n);
}

int dge_pair_moments(const float* a, const float* b, int64_t n, double* out6, void* stream) {
  DGE_REQUIRE(a && b && out6 && n > 0, "pair_moments: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(out6, 0, 6 * sizeof(double), st) != cudaSuccess) { set_error("pair_moments: memset failed"); return DGE_ERR_CUDA; }
  LAUNCH_1D(k_pair_moments, (size_t)n / 4 + 1, stream, a, b, (size_t)n, out6);
}
int dge_softmax_kl_sum(const float* a, const float* b, int64_t outer, int d, int64_t inner, double* out1, void* stream) {
  DGE_REQUIRE(a && b && out1 && outer > 0 && d > 0 && inner > 0, "softmax_kl_sum: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(out1, 0, sizeof(double), st) != cudaSuccess) { set_error("softmax_kl_sum: memset failed"); return DGE_ERR_CUDA; }
  LAUNCH_1D(k_softmax_kl, (size_t)outer * inner, stream, a, b, (size_t)outer, d, (size_t)inner, out1);
}
int dge_avgpool_nchw(const float* x, float* out, int64_t planes, int h_out, int w_out, int factor, void* stream) {
  DGE_REQUIRE(x && out && planes > 0 && h_out > 0 && w_out > 0 && factor >= 1, "avgpool_nchw: bad args");
  LAUNCH_1D(k_avgpool_nchw, (size_t)planes * h_out * w_out, stream, x, out, (size_t)planes, h_out, w_out, factor);
}
// gaussian(11, 1.5) normalised, as metric/pytorch_ssim.py:8-10 (fp32 torch.Tensor arithmetic)
static int ssim_window_init() {
  static bool init = false;
  if (!init) {
    float g[11], sum = 0.f;
    for (int i = 0; i < 11; ++i) { g[i] = (float)exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); sum += g[i]; }
    for (int i = 0; i < 11; ++i) g[i] /= sum;
    if (cudaMemcpyToSymbol(c_gauss11, g, sizeof(g)) != cudaSuccess) { set_error("ssim: constant upload failed"); return DGE_ERR_CUDA; }
    init = true;
  }
  return DGE_OK;
}

int dge_ssim_grad(const float* a, const float* b, const float* go, float* scratch3, float* db, int64_t planes, int h, int w,
                  void* stream) {
  DGE_REQUIRE(a && b && go && scratch3 && db && planes > 0 && h > 0 && w > 0, "ssim_grad: bad args");
  if (int r = ssim_window_init()) return r;
  const size_t work = (size_t)planes * h * w;
  k_ssim_grad_maps<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(a, b, (size_t)planes, h, w, scratch3);
  count_launch();
  if (int r = check_launch("k_ssim_grad_maps")) return r;
  LAUNCH_1D(k_ssim_grad_apply, work, stream, a, b, scratch3, go, (size_t)planes, h, w, db);
}

int dge_ssim_sum(const float* a, const float* b, int64_t planes, int h, int w, double* out1, void* stream) {
  DGE_REQUIRE(a && b && out1 && planes > 0 && h > 0 && w > 0, "ssim_sum: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (int r = ssim_window_init()) return r;
  if (cudaMemsetAsync(out1, 0, sizeof(double), st) != cudaSuccess) { set_error("ssim_sum: memset failed"); return DGE_ERR_CUDA; }
  LAUNCH_1D(k_ssim_sum, (size_t)planes * h * w, stream, a, b, (size_t)planes, h, w, out1);
}

int dge_argmax_mode(const float* logits, int n, int k, int64_t* idx_out, int64_t* mode_out, void* stream) {
  DGE_REQUIRE(logits && idx_out && mode_out && n > 0 && k > 0, "argmax_mode: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  k_argmax_rows<<<(n + 127) / 128, 128, 0, st>>>(logits, n, k, (long long*)idx_out);
  count_launch();
  k_mode<<<1, 32, 0, st>>>((const long long*)idx_out, n, (long long*)mode_out);
  count_launch();
  return check_launch("k_argmax_rows/k_mode");
}
int dge_gradcam(const float* feature, const float* gradient, int plus, double* w_scratch, double* cam_scratch,
                double* out, int n, int c, int h, int w, int h_out, int w_out, void* stream) {
  DGE_REQUIRE(feature && gradient && w_scratch && cam_scratch && out, "gradcam: null pointer");
  DGE_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && h_out > 0 && w_out > 0, "gradcam: bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  k_gradcam_weights<<<n * c, 128, 0, st>>>(gradient, h * w, plus, w_scratch);
  count_launch();
  k_gradcam_map<<<n, 256, 0, st>>>(feature, w_scratch, c, h * w, plus, cam_scratch);
  count_launch();
  k_resize_bilinear<<<grid_for((size_t)n * h_out * w_out, 256), 256, 0, st>>>(cam_scratch, n, h, w, h_out, w_out,
                                                                            plus ? 0 : 1, out);
  count_launch();
  return check_launch("gradcam");
}

int dge_mask2cam(const double* mask, const float* img, const float* lut_rgb, float* heat, float* cam, float* scratch4,
                 int n, int h, int w, void* stream) {
  DGE_REQUIRE(mask && img && lut_rgb && heat && cam && scratch4 && n > 0 && h > 0 && w > 0, "mask2cam: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t per = (size_t)3 * h * w;
  k_jet_overlay<<<grid_for((size_t)n * per, 256), 256, 0, st>>>(mask, img, lut_rgb, heat, cam, n, h * w);
  count_launch();
  const float init[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
  for (int i = 0; i < n; ++i) {
    // cam[i] -= np.max(np.min(cam), 0)  -- the min over the WHOLE array as it is at this point (:248); then /= max(cam[i])
    if (cudaMemcpyAsync(scratch4, init, sizeof(init), cudaMemcpyHostToDevice, st) != cudaSuccess) {
      set_error("mask2cam: scratch init failed");
      return DGE_ERR_CUDA;
    }
    k_axpby<<<grid_for(per, 256), 256, 0, st>>>(heat + i * per, img + i * per, cam + i * per, 1.f, 1.f, per);   // :246
    count_launch();
    k_minmax<<<grid_for((size_t)n * per / 4 + 1, 256), 256, 0, st>>>(cam, (size_t)n * per, scratch4);
    k_sub_or_div<<<grid_for(per, 256), 256, 0, st>>>(cam + i * per, per, scratch4, nullptr);
    k_minmax<<<grid_for(per / 4 + 1, 256), 256, 0, st>>>(cam + i * per, per, scratch4 + 2);
    k_sub_or_div<<<grid_for(per, 256), 256, 0, st>>>(cam + i * per, per, nullptr, scratch4 + 3);
    for (int k = 0; k < 4; ++k) count_launch();
  }
  return check_launch("mask2cam");
}

int dge_lreq_adam_step(void* const* params, const void* const* grads, void* const* vs, const int64_t* numel,
                       const float* step, const int32_t* blk_tensor, const int64_t* blk_off, int n_blocks, int chunk,
                       float beta2, float eps, void* stream) {
  DGE_REQUIRE(params && grads && vs && numel && step && blk_tensor && blk_off, "lreq_adam_step: null pointer");
  DGE_REQUIRE(n_blocks > 0 && chunk > 0, "lreq_adam_step: bad launch shape");
  k_lreq_adam<<<n_blocks, 256, 0, (cudaStream_t)stream>>>((float* const*)params, (const float* const*)grads,
                                                          (float* const*)vs, (const long long*)numel, step, blk_tensor,
                                                          (const long long*)blk_off, chunk, beta2, eps);
  count_launch();
  return check_launch("k_lreq_adam");
}

}  // extern "C"
