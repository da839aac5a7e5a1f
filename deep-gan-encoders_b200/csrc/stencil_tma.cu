// stencil_tma.cu -- HBM-bound stencil kernels staged through shared memory by TMA (sm_100a).
//
// dge_up_fir_epilogue: the second half of StyleGAN2's x2 up-sampling layer
//   model/stylegan2_generator.py:603-615 (UpsamplingLayer: pad (1,1,1,1) + 4x4 FIR [1,3,3,1]^2/16 * 4) and
//   :907-921 (demodulate, noise, bias, leaky-ReLU * sqrt2) over the raw (2H+1)x(2W+1) transposed-conv map that
//   dge_conv_forward(DGE_CONV_UP3X3) wrote:
//     out[y][x] = sum_{a,b<4} f[a] f[b] t[y+a-1][x+b-1],   f = [1,3,3,1]/4,  t = 0 outside the map.
//
// One persistent CTA per SM slot walks (sample, channel group, 16-row strip, 60-pixel span) tiles.  A tile's
// (16+3) x (60+3) halo window of 32-byte channel groups is ONE TMA box (the F32B tensor viewed as uint64 elements so
// the 63-pixel row is a single contiguous inner run); out-of-map rows/columns are zero-filled by the TMA unit, which
// is exactly the reference's zero padding.  Two stages: the window of tile i+1 lands while tile i is filtered, so
// the kernel is paced by HBM, not by load latency (the previous register-ring version reached 2.6 TB/s).
#include "dge_common.cuh"
#include "tma_ptx.cuh"

namespace dge {

#ifndef DGE_FIR_TH
#define DGE_FIR_TH 16
#endif
#ifndef DGE_FIR_OCC
#define DGE_FIR_OCC 2
#endif
constexpr int FIR_TH = DGE_FIR_TH, FIR_TW = 60;         // output tile
constexpr int FIR_RPG = FIR_TH / 2;                     // output rows per thread (two row groups per CTA)
constexpr int FIR_WH = FIR_TH + 3, FIR_WW = FIR_TW + 3; // halo window (rows y0-1 .. y0+17, columns x0-1 .. x0+61)
constexpr int FIR_STAGE_BYTES = FIR_WH * FIR_WW * 32;   // 38304 (bytes one TMA box delivers)
constexpr int FIR_STAGE_STRIDE = (FIR_STAGE_BYTES + 127) / 128 * 128;   // stage buffers stay 128-byte aligned
constexpr int FIR_THREADS = 256;                        // 2 row groups x 64 column lanes (60 active) x 2 channel halves

struct FirParams {
  const float* demod;
  const float* noise;
  long long noise_bstride;
  float noise_scalar;
  const float* bias;
  float slope, gain;
  const float* out_scale;
  void* out_act;
  float* out_nchw;
  int n, c, ho, wo, planes;
  int tiles_x, tiles_y, total_tiles;
};

__device__ __forceinline__ void lds8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__global__ void __launch_bounds__(FIR_THREADS, DGE_FIR_OCC)
k_up_fir_tma(const __grid_constant__ CUtensorMap tm, const __grid_constant__ FirParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* win[2] = {reinterpret_cast<float*>(smem), reinterpret_cast<float*>(smem + FIR_STAGE_STRIDE)};
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * FIR_STAGE_STRIDE);
  const int C8 = p.c >> 3;
  const int per_ng = p.tiles_x * p.tiles_y;

  auto issue = [&](int tile, int stage) {
    const int ng = tile / per_ng, r = tile - ng * per_ng;
    const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
    mbar_arrive_expect_tx(&full[stage], FIR_STAGE_BYTES);
    // coordinates in uint64 elements: 4 per 32-byte channel group
    tma_load_3d(stage ? win[1] : win[0], &tm, &full[stage], 4 * (tx * FIR_TW - 1), ty * FIR_TH - 1, ng);
  };

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int t0 = blockIdx.x, stride = gridDim.x;
  if (threadIdx.x == 0) {
    if (t0 < p.total_tiles) issue(t0, 0);
    if (t0 + stride < p.total_tiles) issue(t0 + stride, 1);
  }
  // thread = (row group of 8 output rows, pixel column, channel half): consecutive lanes read consecutive 16-byte
  // half-groups of the window (conflict-free LDS.128) and write the two halves of one 16-byte ACT chunk
  const int l = threadIdx.x & 127, rg = threadIdx.x >> 7;
  const int xi = l >> 1, half = l & 1;
  const float f[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  uint32_t ph0 = 0, ph1 = 0;
  int stage = 0;
  int cur_ng = -1;
  float dm[4] = {1.f, 1.f, 1.f, 1.f}, bs[4] = {0.f, 0.f, 0.f, 0.f}, gs[4] = {1.f, 1.f, 1.f, 1.f};   // gs = gain * next style
  const bool lrelu_max = p.slope >= 0.f && p.slope <= 1.f;
  for (int tile = t0; tile < p.total_tiles; tile += stride) {
    const int ng = tile / per_ng, r = tile - ng * per_ng;
    const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
    const int nidx = ng / C8, g = ng - nidx * C8;
    const int y0 = ty * FIR_TH + FIR_RPG * rg, x = tx * FIR_TW + xi;
    const bool active = xi < FIR_TW && x < p.wo && y0 < p.ho;
    if (ng != cur_ng) {   // per-(sample, channel group) parameters: reloaded once per ~1000 tiles
      cur_ng = ng;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ch = g * 8 + 4 * half + k;
        dm[k] = p.demod ? __ldg(p.demod + (size_t)nidx * p.c + ch) : 1.f;
        bs[k] = p.bias ? __ldg(p.bias + ch) : 0.f;
        gs[k] = (p.out_scale ? __ldg(p.out_scale + (size_t)nidx * p.c + ch) : 1.f) * p.gain;
      }
    }
    // the tile's noise values are requested before waiting for the window (their latency overlaps the TMA wait).
    // Row addresses are formed ONCE per tile and stepped by the row pitch: the per-row 64-bit index products (ACT index,
    // noise index) were 28 % of the instructions of this kernel, which issues 2 of every 3 cycles (ncu source page)
    float nz[FIR_RPG];
    const float* nzp = p.noise ? p.noise + (size_t)nidx * p.noise_bstride + (size_t)y0 * p.wo + x : nullptr;
#pragma unroll
    for (int oy = 0; oy < FIR_RPG; ++oy) {
      // (raw value: multiplying by the strength here would stall on the load before the TMA wait)
      nz[oy] = (nzp && active && y0 + oy < p.ho) ? __ldg(nzp + oy * p.wo) : 0.f;
    }
    uint2* orow = p.out_act ? reinterpret_cast<uint2*>(reinterpret_cast<uint4*>(p.out_act) +
                                                       act_idx16(nidx, g, 0, y0, x, C8, p.planes, p.ho, p.wo)) + half
                            : nullptr;
    const size_t lo_off = (size_t)2 * p.ho * p.wo;                 // hi -> lo plane, in uint2 units
    float* nrow = p.out_nchw ? p.out_nchw + (((size_t)nidx * p.c + g * 8 + 4 * half) * p.ho + y0) * p.wo + x : nullptr;
    const size_t nch_off = (size_t)p.ho * p.wo;
    mbar_wait(&full[stage], stage ? ph1 : ph0);
    if (stage) ph1 ^= 1; else ph0 ^= 1;
    if (active) {
      // window rows 8*rg .. 8*rg+10 feed output rows y0 .. y0+7: horizontal 4-tap pass per row into a 4-row ring,
      // vertical pass from the ring
      const float* w0 = (stage ? win[1] : win[0]) + ((size_t)(FIR_RPG * rg) * FIR_WW + xi) * 8 + 4 * half;
      float ring[4][4];
#pragma unroll
      for (int rr = 0; rr < FIR_RPG + 3; ++rr) {
        float h[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float4 v = *reinterpret_cast<const float4*>(w0 + ((size_t)rr * FIR_WW + b) * 8);
          h[0] = fmaf(f[b], v.x, h[0]);
          h[1] = fmaf(f[b], v.y, h[1]);
          h[2] = fmaf(f[b], v.z, h[2]);
          h[3] = fmaf(f[b], v.w, h[3]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) ring[rr & 3][k] = h[k];
        if (rr < 3) continue;
        const int oy = rr - 3, y = y0 + oy;
        if (y >= p.ho) continue;
        float acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // window rows oy .. oy+3 live in ring slots (rr-3 .. rr) & 3 with taps f[0..3]
          const float v = f[0] * ring[(rr - 3) & 3][k] + f[1] * ring[(rr - 2) & 3][k] + f[2] * ring[(rr - 1) & 3][k] +
                          f[3] * ring[rr & 3][k];
          const float z = fmaf(v, dm[k], fmaf(nz[oy], p.noise_scalar, bs[k]));
          acc[k] = lrelu_max ? fmaxf(z, z * p.slope) : (z < 0.f ? z * p.slope : z);   // (gain applied with the scale)
        }
        if (nrow) {
#pragma unroll
          for (int k = 0; k < 4; ++k) nrow[k * nch_off + (size_t)oy * p.wo] = acc[k] * p.gain;
        }
        if (orow) {
          uint32_t hw[2], lw[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float a = acc[2 * i] * gs[2 * i], b2 = acc[2 * i + 1] * gs[2 * i + 1];
            const __nv_bfloat162 hb = __floats2bfloat162_rn(a, b2);
            hw[i] = *reinterpret_cast<const uint32_t*>(&hb);
            const __nv_bfloat162 lb = __floats2bfloat162_rn(a - __uint_as_float(hw[i] << 16),
                                                            b2 - __uint_as_float(hw[i] & 0xffff0000u));
            lw[i] = *reinterpret_cast<const uint32_t*>(&lb);
          }
          uint2* o = orow + (size_t)oy * (2 * p.wo);
          *o = make_uint2(hw[0], hw[1]);
          if (p.planes == 2) o[lo_off] = make_uint2(lw[0], lw[1]);
        }
      }
    }
    __syncthreads();   // every thread is done reading this stage
    if (threadIdx.x == 0 && tile + 2 * stride < p.total_tiles) issue(tile + 2 * stride, stage);
    stage ^= 1;
  }
}

}  // namespace dge

using namespace dge;

extern "C" int dge_up_fir_epilogue(const float* raw_up, const float* demod, const float* noise, int64_t noise_bstride,
                                   float noise_scalar, const float* bias, float slope, float gain,
                                   const float* out_scale, void* out_act, float* out_nchw, int n, int c, int h_out,
                                   int w_out, int planes, void* stream) {
  DGE_REQUIRE(raw_up && (out_act || out_nchw), "up_fir_epilogue: null pointer");
  DGE_REQUIRE(n > 0 && c > 0 && c % 8 == 0 && h_out > 0 && w_out > 0 && h_out % 2 == 0 && w_out % 2 == 0,
              "up_fir_epilogue: bad dims n=%d c=%d h=%d w=%d", n, c, h_out, w_out);
  DGE_REQUIRE(!out_act || planes == 1 || planes == 2, "up_fir_epilogue: planes=%d", planes);
  DGE_REQUIRE((reinterpret_cast<uintptr_t>(raw_up) & 15) == 0, "up_fir_epilogue: raw_up must be 16-byte aligned");
  const int hi = h_out + 1, wi = w_out + 1, C8 = c / 8;
  CUtensorMap tm;
  {
    // F32B [n*C8][hi][wi][8 floats] viewed as uint64 [n*C8][hi][4*wi]
    const uint64_t dims[3] = {(uint64_t)4 * wi, (uint64_t)hi, (uint64_t)n * C8};
    const uint64_t strides[2] = {(uint64_t)wi * 32, (uint64_t)hi * wi * 32};
    const uint32_t box[3] = {4 * FIR_WW, FIR_WH, 1};
    const int r = make_tmap(&tm, raw_up, 3, dims, strides, box);
    if (r) return r;
  }
  FirParams p;
  p.demod = demod; p.noise = noise; p.noise_bstride = noise_bstride; p.noise_scalar = noise_scalar;
  p.bias = bias; p.slope = slope; p.gain = gain; p.out_scale = out_scale;
  p.out_act = out_act; p.out_nchw = out_nchw;
  p.n = n; p.c = c; p.ho = h_out; p.wo = w_out; p.planes = planes;
  p.tiles_x = (w_out + FIR_TW - 1) / FIR_TW;
  p.tiles_y = (h_out + FIR_TH - 1) / FIR_TH;
  const long long total = (long long)n * C8 * p.tiles_x * p.tiles_y;
  DGE_REQUIRE(total < (1ll << 31), "up_fir_epilogue: too many tiles");
  p.total_tiles = (int)total;
  const size_t smem = 2 * FIR_STAGE_STRIDE + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_up_fir_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(k_up_fir_tma) failed: %s", cudaGetErrorString(e));
      return DGE_ERR_CUDA;
    }
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = DGE_FIR_OCC * sms;
  if (grid > p.total_tiles) grid = p.total_tiles;
  k_up_fir_tma<<<grid, FIR_THREADS, smem, (cudaStream_t)stream>>>(tm, p);
  count_launch();
  return check_launch("k_up_fir_tma");
}
