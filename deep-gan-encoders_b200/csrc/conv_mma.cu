// conv_mma.cu -- im2col-free implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Replaces the reference's ATen chains around F.conv2d / F.conv_transpose2d:
//   ModulateConvBlock.forward   model/stylegan2_generator.py:855-922 (plain :897-904, up :879-895)
//   BEBlock.forward convs       model/E/E.py:59-62, 72-75, 81-84 ; ln.Conv2d.forward model/utils/lreq.py:126-156
//
// Formulation: y = d[n,o] * conv(x * s[n,i], W) -- the style scale is folded into the
// PRODUCER of x (its epilogue multiplies by the next layer's style), so one shared weight tensor serves the
// whole batch and the conv is a dense GEMM:  D[pixel, cout] = sum_{tap, cin} A[pixel+tap, cin] * B[tap][cout, cin].
//
// Data path per CTA (persistent, one output tile = 16x8 pixels x <=256 output columns):
//   TMA  : one 18x10-pixel halo patch per K-chunk of channels (ACT layout, 16-byte channel groups -> the
//          smem image IS the canonical no-swizzle K-major UMMA layout, so all 9 taps are just different
//          descriptor start addresses into the same patch -- no im2col, no 9x re-read);
//          one weight slab per (tap, K-chunk).
//   MMA  : tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM), M=128, N<=128 per instruction.
//          planes==2 => split precision: x = xh+xl, w = wh+wl, acc += xh*wh + xh*wl + xl*wh  (bf16x3,
//          ~2^-16 relative operand error -> meets the 1e-3 parity bar that plain bf16 misses).
//   TMEM : two accumulator stages so the epilogue of tile i overlaps the mainloop of tile i+1.
//   EPI  : 4 warps, one TMEM lane (= pixel) per thread: demod, noise, bias, lrelu, residual blend, fused
//          ToRGB partial sums, next-layer style scale, bf16 hi/lo split, 16/32-byte vector stores.
#include <stdarg.h>
#include <stdlib.h>

#include "dge_common.cuh"
#include "tma_ptx.cuh"

namespace dge {

// One MMA covers an M block of 16x8 output pixels (BW=8 -> one UMMA core-matrix row group per block row).  A CTA tile is
// msub = 1, 2 or 4 such blocks (16x8, 16x16, 32x16 pixels) that share one halo patch (tile_h+2) x (tile_w+2): the
// per-tile fixed costs (barrier round trips, tile decode, parameter checks) amortise over up to 512 pixels, the
// epilogue overlaps the TMEM load of one block with the arithmetic of the previous one, and the halo overhead drops
// from 1.41x to 1.20x.
constexpr int BH = 16, BW = 8;
constexpr int MAX_B_SLOTS = 8;
constexpr int NUM_EPI_WARPS = 8;          // two warps per TMEM lane quarter; they split a tile's (M block, column group) units
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = 64 + NUM_EPI_THREADS;   // warp0: TMA, warp1: MMA + TMEM alloc, warps 2..9: epilogue; 1 CTA / SM

struct TapEntry {
  int16_t a_off;  // patch pixel offset dy*pw+dx of this tap
  int16_t w_tap;  // tap index into WPK
  int16_t phase;  // output phase (0 for stride-1 convs, 2*py+px for the transposed conv)
  int16_t in_phase;  // -1: applies to every K chunk; else only to chunks of this input phase (space-to-depth conv)
};

// division by a runtime constant: q = (umulhi(n, mul) + n) >> shr   (host computes mul/shr, n < 2^31)
struct FastDiv {
  uint32_t mul, shr;
};
__device__ __forceinline__ uint32_t fast_div(uint32_t n, FastDiv d) {
  return (uint32_t)(((uint64_t)__umulhi(n, d.mul) + n) >> d.shr);
}
static FastDiv make_fast_div(uint32_t d) {
  FastDiv f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.mul = (uint32_t)((((1ull << l) - d) << 32) / d + 1);
  f.shr = l;
  return f;
}

constexpr int EPI_TAB_COLS = 256;                      // widest CTA column block
constexpr int EPI_TAB_BYTES = 7 * EPI_TAB_COLS * 4;    // A, B, NW, S, RGB[3] rows of the per-(sample, N tile) table

struct IssueEnt {
  int16_t a_off;   // patch pixel offset of the tap (16-byte units into the A slab)
  int16_t dcol;    // accumulator column block of the tap's output phase
  int16_t first;   // 1: first tap of its phase in issue order (clears the accumulator on the first K chunk)
  int16_t w_tap;
};

struct ConvKParams {
  int N, H, W;             // input dims
  int dom_h, dom_w;        // tile domain (H,W) or (H+1,W+1) for the transposed conv
  int tiles_x, tiles_y;
  int Cin, Cout, P;
  int cin_w;               // input channels per weight tap (== Cin, or Cin/4 for the space-to-depth stride-2 conv)
  int kc, nchunks;         // channels per A chunk, number of chunks
  int ntile, cw, nsub;     // CTA columns, columns per phase, columns per MMA
  int n_ntiles, np;        // N tiles, phases per CTA tile
  int planes;
  int total_tiles;
  int sched_tiles;         // scheduling units: tiles, or tile PAIRS (two M tiles, same N tile) in pair mode
  int mtiles;              // M tiles (sample x tile rows x tile columns)
  int pair;                // 1: CTA pair (cta_group::2) launch
  int ksplit;              // split-K factor: unit = (tile or pair, K part); partial sums are ADDED atomically to `ws`
  int sched_units;         // sched_tiles * ksplit
  FastDiv div_ksplit;
  float* ws;               // split-K scratch, F32B [n][cout/8][h][w][8] (zeroed by the host wrapper)
  int nsub_local;          // weight rows this CTA holds in smem per slab (= nsub, or nsub/2 in pair mode)
  int ntaps;
  TapEntry taps[16];
  int a_slot_bytes, b_slot_bytes, b_sub_bytes, b_slots;
  int a_slots, b_region_bytes, resident, acc_stages, b_rb;
  int tmem_cols;
  int msub;                // M blocks per tile
  int tile_h, tile_w;      // tile extent in pixels
  int pw, ph, patch_bytes; // halo patch extent and bytes of one (channel-group, plane) slab
  int sb_off[4];           // patch pixel offset of M block sb
  int sb_y[4], sb_x[4];    // pixel offset of M block sb inside the tile
  int sb_cols;             // TMEM columns of one M block = np * acc_cols
  int np_shift;            // log2(np)
  IssueEnt ilist[16];      // issue order: [input phase (space-to-depth conv only)][tpc]
  int n_cph;               // input phases in ilist (1, or 4 for the space-to-depth conv)
  int tpc;                 // taps issued per K chunk (ntaps, or 4 for the space-to-depth conv)
  int stack;               // 1: hi|lo weight planes stacked on N (A_hi x [B_hi|B_lo] + A_lo x B_hi: 2 MMAs instead of 3)
  int acc_cols;            // TMEM columns of one phase block = cw * (stack ? 2 : 1)
  int tm_stride;           // TMEM columns of one accumulator stage = msub * np * acc_cols
  FastDiv div_ntiles, div_per_img, div_tiles_x;
  int act_mode;            // 0: identity (slope == 1), 1: max(v, v*slope) (0 <= slope <= 1), 2: select form
  // epilogue
  const float* demod;
  const float* noise;
  long long noise_bstride;
  const float* noise_w;
  float noise_scalar;
  const float* bias;
  float slope, gain;
  const float* preact_add;
  int preact_c, preact_up;
  const float* blend_src;
  int blend_pool;
  float blend_a, blend_b;
  void* out_act;
  int out_planes;
  const float* out_scale;
  float* out_f32b;
  float* out_nchw;
  const float* rgb_w;
  float* rgb_out;
  float* out_raw_up;
  int out_pool;
  // checker kernel only
  const void* x;
  const void* wpk;
};

// role timing (experiments only: `make exp` builds tools/_exp/libdge_exp.so with -DDGE_ROLE_TIMING)
#ifdef DGE_ROLE_TIMING
__device__ unsigned long long g_role_cycles[16];
#define ROLE_T0() long long _rt0 = clock64()
#define ROLE_ACC(var) do { long long _t = clock64(); (var) += _t - _rt0; _rt0 = _t; } while (0)
#define ROLE_FLUSH(slot, var) do { if (lane == 0) atomicAdd(&g_role_cycles[slot], (unsigned long long)(var)); } while (0)
#else
#define ROLE_T0() do {} while (0)
#define ROLE_ACC(var) do {} while (0)
#define ROLE_FLUSH(slot, var) do {} while (0)
#endif

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// PAIR = the CTA pair form (cta_group::2): one MMA spans both SMs of a 2-CTA cluster (M = 256: 128 output pixels per
// CTA), each CTA supplies its own A patch and HALF of the weight rows, so the weight operand is fetched from shared
// memory once per pair instead of once per SM.  The commit is multicast to the same barrier offset in both CTAs.
template <bool PAIR>
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  if (PAIR)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  if (PAIR)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on a barrier of the pair's leader CTA (`bar_cluster` = mapa address)
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no-swizzle ("interleaved") shared-memory matrix descriptor:
//   element (row r, 16-byte k-group j) lives at start + (r/8)*SBO + (r%8)*16 + j*LBO.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc_bf16(int n, int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// epilogues (shared by the tcgen05 kernel and the checker kernel)
// ---------------------------------------------------------------------------------------------
struct PixelCtx {
  int n, y, x;
  bool valid;
  float nz;
};

// value part of the pointwise epilogue: demod, noise, bias, pre-activation residual, activation, gain
__device__ __forceinline__ void epi_value16(const ConvKParams& p, const PixelCtx& px, int c0, float* v) {
  const int H = p.H, W = p.W;
  if (p.demod) {
    const float4* d = reinterpret_cast<const float4*>(p.demod + (size_t)px.n * p.Cout + c0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 t = __ldg(d + q);
      v[4 * q] *= t.x; v[4 * q + 1] *= t.y; v[4 * q + 2] *= t.z; v[4 * q + 3] *= t.w;
    }
  }
  if (p.noise) {
    if (p.noise_w) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(px.nz, __ldg(p.noise_w + c0 + j), v[j]);
    } else {
      const float t = px.nz * p.noise_scalar;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += t;
    }
  }
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += __ldg(p.bias + c0 + j);
  }
  if (p.preact_add && px.valid) {
    float r[8];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      // residual may have more channels than the output (BigGAN channel drop) and half its resolution (nearest x2)
      load8_f32b(p.preact_add, f32b_idx32(px.n, (c0 >> 3) + g, px.y / p.preact_up, px.x / p.preact_up, p.preact_c >> 3,
                                          H / p.preact_up, W / p.preact_up), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * g + j] += r[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = (v[j] < 0.f ? v[j] * p.slope : v[j]) * p.gain;
}

// pointwise epilogue on 16 consecutive output channels c0..c0+15 of one pixel (generic form: the CUDA-core checker
// kernel and the split-K finish kernel; the tcgen05 kernel has its own table-driven version, epi_group16)
__device__ __forceinline__ void epi_pointwise16(const ConvKParams& p, const PixelCtx& px, int c0, float* v,
                                                float* rgb) {
  const int H = p.H, W = p.W, C8 = p.Cout >> 3;
  epi_value16(p, px, c0, v);
  if (!px.valid) return;
  if (p.blend_src) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float s[8];
      const int c8 = (c0 >> 3) + g;
      if (p.blend_pool) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            load8_f32b(p.blend_src, f32b_idx32(px.n, c8, 2 * px.y + dy, 2 * px.x + dx, C8, 2 * H, 2 * W), t);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += t[j];
          }
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] *= 0.25f;
      } else {
        load8_f32b(p.blend_src, f32b_idx32(px.n, c8, px.y, px.x, C8, H, W), s);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * g + j] = p.blend_a * s[j] + p.blend_b * v[8 * g + j];
    }
  }
  if (p.rgb_w) {
    const float* w = p.rgb_w + (size_t)px.n * 3 * p.Cout + c0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float acc = rgb[ch];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc = fmaf(v[j], __ldg(w + ch * p.Cout + j), acc);
      rgb[ch] = acc;
    }
  }
  if (p.out_f32b) {
    store8_f32b(p.out_f32b, f32b_idx32(px.n, c0 >> 3, px.y, px.x, C8, H, W), v);
    store8_f32b(p.out_f32b, f32b_idx32(px.n, (c0 >> 3) + 1, px.y, px.x, C8, H, W), v + 8);
  }
  if (p.out_nchw) {
#pragma unroll
    for (int j = 0; j < 16; ++j) p.out_nchw[(((size_t)px.n * p.Cout + c0 + j) * H + px.y) * W + px.x] = v[j];
  }
  if (p.out_act) {
    float t[16];
    if (p.out_scale) {
      const float* s = p.out_scale + (size_t)px.n * p.Cout + c0;
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = v[j] * __ldg(s + j);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = v[j];
    }
    store8_act(p.out_act, px.n, c0 >> 3, px.y, px.x, C8, p.out_planes, H, W, t);
    store8_act(p.out_act, px.n, (c0 >> 3) + 1, px.y, px.x, C8, p.out_planes, H, W, t + 8);
  }
}

// raw transposed-conv epilogue: pixel (Y,X) of the (H+1)x(W+1) domain, phase (py,px) -> t[2Y+py][2X+px]
template <bool SPLITK>
__device__ __forceinline__ void epi_rawup16(const ConvKParams& p, int n, int Y, int X, bool valid, int phase, int c0,
                                            const float* v) {
  const int py = phase >> 1, pxs = phase & 1;
  const int Ho = 2 * p.H + 1, Wo = 2 * p.W + 1;
  const int oy = 2 * Y + py, ox = 2 * X + pxs;
  if (!valid || oy >= Ho || ox >= Wo) return;
  const int C8 = p.Cout >> 3;
  if (SPLITK) {   // split-K: the raw map is linear in the accumulator -- partial sums are added in place
    float4* o0 = reinterpret_cast<float4*>(p.out_raw_up) + f32b_idx32(n, c0 >> 3, oy, ox, C8, Ho, Wo) * 2;
    float4* o1 = reinterpret_cast<float4*>(p.out_raw_up) + f32b_idx32(n, (c0 >> 3) + 1, oy, ox, C8, Ho, Wo) * 2;
    atomicAdd(o0, make_float4(v[0], v[1], v[2], v[3]));
    atomicAdd(o0 + 1, make_float4(v[4], v[5], v[6], v[7]));
    atomicAdd(o1, make_float4(v[8], v[9], v[10], v[11]));
    atomicAdd(o1 + 1, make_float4(v[12], v[13], v[14], v[15]));
    return;
  }
  store8_f32b(p.out_raw_up, f32b_idx32(n, c0 >> 3, oy, ox, C8, Ho, Wo), v);
  store8_f32b(p.out_raw_up, f32b_idx32(n, (c0 >> 3) + 1, oy, ox, C8, Ho, Wo), v + 8);
}

struct TileCoord {
  int n, y0, x0, nt, p0, co0;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int tile) {
  TileCoord t;
  const int mt = (int)fast_div((uint32_t)tile, p.div_ntiles);
  t.nt = tile - mt * p.n_ntiles;
  const int per_img = p.tiles_x * p.tiles_y;
  t.n = (int)fast_div((uint32_t)mt, p.div_per_img);
  const int r = mt - t.n * per_img;
  const int ty = (int)fast_div((uint32_t)r, p.div_tiles_x);
  t.y0 = ty * p.tile_h;
  t.x0 = (r - ty * p.tiles_x) * p.tile_w;
  // an N tile holds `np` phases x `cw` channels: channels [nt*cw, (nt+1)*cw) of every phase
  t.p0 = 0;
  t.co0 = t.nt * p.cw;
  return t;
}

// ---------------------------------------------------------------------------------------------
// lean epilogue of the tcgen05 kernel
//   Per-channel parameters come from a shared-memory table (rows A, B, NW, S, RGB0..2 of EPI_TAB_COLS floats) that the
//   128 epilogue threads refill only when the (sample, N tile) changes.  With g = gain (lrelu is positively
//   homogeneous, so the gain is folded into the affine part):
//     A[c] = demod[n][c]*g   B[c] = bias[c]*g   NW[c] = noise weight*g   S[c] = next-layer style   RGB = ToRGB weights
//     v = max(z, z*slope),  z = acc*A + B + nz*NW   (+ residual*g)
// ---------------------------------------------------------------------------------------------
struct EpiPix {
  int n, y, x;
  bool valid;
  float nz;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_THREADS) : "memory"); }

__device__ __forceinline__ void epi_fill_table(const ConvKParams& p, float* tab, int n, int co0, int et) {
  for (int c = et; c < p.cw; c += NUM_EPI_THREADS) {
    const int co = co0 + c;
    tab[c] = (p.demod ? __ldg(p.demod + (size_t)n * p.Cout + co) : 1.f) * p.gain;
    tab[EPI_TAB_COLS + c] = (p.bias ? __ldg(p.bias + co) : 0.f) * p.gain;
    tab[2 * EPI_TAB_COLS + c] = (p.noise ? (p.noise_w ? __ldg(p.noise_w + co) : p.noise_scalar) : 0.f) * p.gain;
    tab[3 * EPI_TAB_COLS + c] = p.out_scale ? __ldg(p.out_scale + (size_t)n * p.Cout + co) : 1.f;
    if (p.rgb_w) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        tab[(4 + ch) * EPI_TAB_COLS + c] = __ldg(p.rgb_w + ((size_t)n * 3 + ch) * p.Cout + co);
    }
  }
}

__device__ __forceinline__ void tm_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// the "+r" operands tie the loaded registers to the wait so no use can be scheduled ahead of it
__device__ __forceinline__ void tm_ld16_wait(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// 8 floats -> bf16 hi chunk and bf16 lo chunk (x ~= hi + lo): 2 packed converts, 2 unpacks and 2 subtracts per pair
__device__ __forceinline__ void split8_fast(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&hb);
    const float r0 = v[2 * i] - __uint_as_float(h[i] << 16);
    const float r1 = v[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
    const __nv_bfloat162 lb = __floats2bfloat162_rn(r0, r1);
    l[i] = *reinterpret_cast<const uint32_t*>(&lb);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// r += s (fp32 bit patterns): the two column halves of a stacked accumulator
__device__ __forceinline__ void epi_fold16(uint32_t* r, const uint32_t* s) {
#pragma unroll
  for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(s[j]));
}

// experiment switches (timing only, wrong results): -DDGE_X_NOSTORE / _NORGB / _NOMATH
template <int EPI, bool SPLITK>
__device__ __forceinline__ void epi_group16(const ConvKParams& p, const float* tab, const EpiPix& px, int co0, int p0,
                                            int q, int c, const uint32_t* r, float* rgb, const float* bl = nullptr) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  if (EPI == 1) {
    epi_rawup16<SPLITK>(p, px.n, px.y, px.x, px.valid, p0 + q, co0 + c, v);
    __syncwarp();
    return;
  }
  if (SPLITK) {
    // split-K: this CTA holds a partial sum over its K chunks -- add it to the scratch map; the nonlinear epilogue
    // runs in conv_splitk_finish_kernel once every part has landed
    if (px.valid) {
      const size_t HWs = (size_t)p.H * p.W;
      float4* o = reinterpret_cast<float4*>(p.ws) +
                  ((((size_t)px.n * (p.Cout >> 3) + ((co0 + c) >> 3)) * HWs + (size_t)px.y * p.W + px.x) * 2);
      atomicAdd(o, make_float4(v[0], v[1], v[2], v[3]));
      atomicAdd(o + 1, make_float4(v[4], v[5], v[6], v[7]));
      atomicAdd(o + 2 * HWs, make_float4(v[8], v[9], v[10], v[11]));
      atomicAdd(o + 2 * HWs + 1, make_float4(v[12], v[13], v[14], v[15]));
    }
    __syncwarp();
    return;
  }
  const float4* A4 = reinterpret_cast<const float4*>(tab + c);
  const float4* B4 = reinterpret_cast<const float4*>(tab + EPI_TAB_COLS + c);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 a = A4[k], b = B4[k];
    v[4 * k] = fmaf(v[4 * k], a.x, b.x);
    v[4 * k + 1] = fmaf(v[4 * k + 1], a.y, b.y);
    v[4 * k + 2] = fmaf(v[4 * k + 2], a.z, b.z);
    v[4 * k + 3] = fmaf(v[4 * k + 3], a.w, b.w);
  }
  if (p.noise) {
    const float4* N4 = reinterpret_cast<const float4*>(tab + 2 * EPI_TAB_COLS + c);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 w = N4[k];
      v[4 * k] = fmaf(px.nz, w.x, v[4 * k]);
      v[4 * k + 1] = fmaf(px.nz, w.y, v[4 * k + 1]);
      v[4 * k + 2] = fmaf(px.nz, w.z, v[4 * k + 2]);
      v[4 * k + 3] = fmaf(px.nz, w.w, v[4 * k + 3]);
    }
  }
  const int H = p.H, W = p.W, C8 = p.Cout >> 3, g0 = (co0 + c) >> 3;
  if (EPI != 3 && EPI != 4 && p.preact_add && px.valid) {
    float t[8];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      // residual may have more channels than the output (BigGAN channel drop) and half its resolution (nearest x2)
      load8_f32b(p.preact_add, f32b_idx32(px.n, g0 + g, px.y / p.preact_up, px.x / p.preact_up, p.preact_c >> 3,
                                          H / p.preact_up, W / p.preact_up), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * g + j] = fmaf(t[j], p.gain, v[8 * g + j]);
    }
  }
  if (p.act_mode == 1) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], v[j] * p.slope);
  } else if (p.act_mode == 2) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = v[j] < 0.f ? v[j] * p.slope : v[j];
  }
  if (EPI == 4 || (EPI != 3 && p.out_pool)) {
    // 2x2 mean across the lanes (x ^ 1 <-> lane ^ 1, y ^ 1 <-> lane ^ 8); the even/even lane stores the pooled pixel
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float sm = v[j] + __shfl_xor_sync(0xffffffffu, v[j], 1);
      sm += __shfl_xor_sync(0xffffffffu, sm, 8);
      v[j] = 0.25f * sm;
    }
    if (px.valid && !(lane & 9)) {
      const size_t HWp = (size_t)(H >> 1) * (W >> 1);
      const size_t o = ((size_t)px.n * C8 + g0) * HWp + (size_t)(px.y >> 1) * (W >> 1) + (px.x >> 1);
      store8_f32b(p.out_f32b, o, v);
      store8_f32b(p.out_f32b, o + HWp, v + 8);
    }
    __syncwarp();
    return;
  }
  if (px.valid) {
    if (EPI != 3 && EPI != 4 && p.blend_src) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float sv[8];
        if (p.blend_pool) {
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) sv[j] = 0.f;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              load8_f32b(p.blend_src, f32b_idx32(px.n, g0 + g, 2 * px.y + dy, 2 * px.x + dx, C8, 2 * H, 2 * W), t);
#pragma unroll
              for (int j = 0; j < 8; ++j) sv[j] += t[j];
            }
#pragma unroll
          for (int j = 0; j < 8; ++j) sv[j] *= 0.25f;
        } else if (bl) {
#pragma unroll
          for (int j = 0; j < 8; ++j) sv[j] = bl[8 * g + j];   // fetched one unit ahead (see the unit loop)
        } else {
          load8_f32b(p.blend_src, f32b_idx32(px.n, g0 + g, px.y, px.x, C8, H, W), sv);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * g + j] = p.blend_a * sv[j] + p.blend_b * v[8 * g + j];
      }
    }
#ifndef DGE_X_NORGB
    if (p.rgb_w) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float4* R4 = reinterpret_cast<const float4*>(tab + (4 + ch) * EPI_TAB_COLS + c);
        float acc = rgb[ch];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 w = R4[k];
          acc = fmaf(v[4 * k], w.x, acc);
          acc = fmaf(v[4 * k + 1], w.y, acc);
          acc = fmaf(v[4 * k + 2], w.z, acc);
          acc = fmaf(v[4 * k + 3], w.w, acc);
        }
        rgb[ch] = acc;
      }
    }
#endif
    const size_t HW = (size_t)H * W, pix = (size_t)px.y * W + px.x;
#ifdef DGE_X_NOSTORE
    float xsum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) xsum += v[j];
    if (xsum != -1.2345e30f) { __syncwarp(); return; }
#endif
    if (p.out_f32b) {
      const size_t o = ((size_t)px.n * C8 + g0) * HW + pix;
      store8_f32b(p.out_f32b, o, v);
      store8_f32b(p.out_f32b, o + HW, v + 8);
    }
    if (EPI != 3 && EPI != 4 && p.out_nchw) {
#pragma unroll
      for (int j = 0; j < 16; ++j) p.out_nchw[((size_t)px.n * p.Cout + co0 + c + j) * HW + pix] = v[j];
    }
    if (p.out_act) {
      const float4* S4 = reinterpret_cast<const float4*>(tab + 3 * EPI_TAB_COLS + c);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 w = S4[k];
        v[4 * k] *= w.x; v[4 * k + 1] *= w.y; v[4 * k + 2] *= w.z; v[4 * k + 3] *= w.w;
      }
      uint4* o = reinterpret_cast<uint4*>(p.out_act) + ((size_t)px.n * C8 + g0) * p.out_planes * HW + pix;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint4 hi, lo;
        split8_fast(v + 8 * g, hi, lo);
        o[(size_t)g * p.out_planes * HW] = hi;
        if (p.out_planes == 2) o[(size_t)g * p.out_planes * HW + HW] = lo;
      }
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// the tcgen05 kernel
// ---------------------------------------------------------------------------------------------
constexpr int MAX_A_SLOTS = 4;
constexpr int BAR_A_FULL = 0, BAR_A_EMPTY = MAX_A_SLOTS, BAR_TM_FULL = 2 * MAX_A_SLOTS, BAR_TM_EMPTY = BAR_TM_FULL + 2,
              BAR_B_FULL = BAR_TM_EMPTY + 2, BAR_B_EMPTY = BAR_B_FULL + MAX_B_SLOTS, BAR_COUNT = BAR_B_EMPTY + MAX_B_SLOTS;

// all MMAs of one (tap, K-chunk) into accumulator column block d.
//   MODE 0: plain bf16 (1 MMA per K-step)
//   MODE 1: split precision, 3 MMAs per K-step: hi*hi, hi*lo, lo*hi
//   MODE 2: split precision with the weight planes stacked on N: A_hi x [B_hi|B_lo] (N = 2*cw, the hi and lo rows
//           of one k-group are contiguous in smem) + A_lo x B_hi -- 2 MMAs per K-step.  For cw <= 32 an MMA is bound
//           by the 4 KB A-operand read from shared memory (32 clk) whatever N is, so this is 1.5x fewer tensor-pipe
//           cycles; the epilogue adds the two column halves.
//   KSTEPS > 0 unrolls the K loop (straight-line issue); 0 = runtime count.
struct IssueConsts {
  uint32_t idesc, idesc2, a_kstep, b_kstep, a_lo16, b_lo16;
  int ksteps;
};
template <int MODE, bool PAIR>
__device__ __forceinline__ void issue_kstep(uint32_t d, uint64_t da, uint64_t db, const IssueConsts& c,
                                            uint32_t accumulate) {
  if (MODE == 2) {
    tc_mma_bf16<PAIR>(d, da, db, c.idesc2, accumulate);
    tc_mma_bf16<PAIR>(d, da + c.a_lo16, db, c.idesc, 1u);
  } else {
    tc_mma_bf16<PAIR>(d, da, db, c.idesc, accumulate);
    if (MODE == 1) {
      tc_mma_bf16<PAIR>(d, da, db + c.b_lo16, c.idesc, 1u);
      tc_mma_bf16<PAIR>(d, da + c.a_lo16, db, c.idesc, 1u);
    }
  }
}
template <int MODE, int KSTEPS, bool PAIR>
__device__ __forceinline__ void issue_tap(uint32_t d, uint64_t da, uint64_t db, const IssueConsts& c,
                                          uint32_t accumulate) {
  if (KSTEPS > 0) {
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k)
      issue_kstep<MODE, PAIR>(d, da + (uint32_t)k * c.a_kstep, db + (uint32_t)k * c.b_kstep, c,
                              k == 0 ? accumulate : 1u);
  } else {
#pragma unroll 1
    for (int k = 0; k < c.ksteps; ++k) {
      issue_kstep<MODE, PAIR>(d, da, db, c, accumulate);
      accumulate = 1u;
      da += c.a_kstep;
      db += c.b_kstep;
    }
  }
}

// split-K: scheduling unit u = su * ksplit + ks works on K chunks [c0, c1) of tile unit su
template <bool SPLITK>
__device__ __forceinline__ int sched_split(const ConvKParams& p, int unit, int& c0, int& c1) {
  if (!SPLITK) {
    c0 = 0;
    c1 = p.nchunks;
    return unit;
  }
  const int su = (int)fast_div((uint32_t)unit, p.div_ksplit);
  const int ks = unit - su * p.ksplit;
  c0 = (ks * p.nchunks) / p.ksplit;
  c1 = ((ks + 1) * p.nchunks) / p.ksplit;
  return su;
}

// per-CTA barrier / shared-memory handles of the MMA role
struct MmaBars {
  uint64_t *a_full, *a_empty, *b_full, *b_empty, *tm_full, *tm_empty;
  uint32_t a_smem16, b_region16, a_slot16, b_slot16, tm_base;
};

// The MMA role's tile loop.  A single warp issues every MMA of the CTA, and a taken branch costs it ~25 cycles, so for
// the 16/32-channel layers (18..54 small MMAs per tile) the issue path -- not the tensor pipe -- is the critical path
// unless the per-chunk code is straight-line: TPC (taps per chunk) and KSTEPS > 0 unroll everything between two
// barrier waits; all operands derive from kernel parameters and uniform counters.  TPC == 0 / KSTEPS == 0 are the
// generic runtime-loop forms.
template <int MODE, bool RESIDENT, int TPC, int KSTEPS, bool PAIR, bool SPLITK>
__device__ __forceinline__ void mma_tiles(const ConvKParams& p, const MmaBars& m, const IssueConsts& ic,
                                          uint64_t a_desc0, uint64_t b_desc0) {
  uint32_t a_slot = 0, a_ph = 0, b_slot = 0, b_ph = 0, acc = 0, acc_ph = 0;
  [[maybe_unused]] long long rt_tm = 0, rt_a = 0, rt_issue = 0;
  [[maybe_unused]] const int lane = threadIdx.x & 31;
  const int tpc = TPC > 0 ? TPC : p.tpc;
  if (RESIDENT) {
    mbar_wait(&m.b_full[0], 0);
    tc_fence_after();
  }
  ROLE_T0();
  const int t_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  for (int unit = t_first; unit < p.sched_units; unit += t_step) {
    int ch0, ch1;
    sched_split<SPLITK>(p, unit, ch0, ch1);
    ROLE_ACC(rt_issue);
    mbar_wait(&m.tm_empty[acc], acc_ph ^ 1);
    ROLE_ACC(rt_tm);
    tc_fence_after();
    const uint32_t d_base = m.tm_base + acc * p.tm_stride;
    uint32_t b_res16 = m.b_region16 + (uint32_t)(ch0 * tpc) * m.b_slot16;   // resident: slabs in (chunk, tap) order
    for (int ch = ch0; ch < ch1; ++ch) {
      ROLE_ACC(rt_issue);
      mbar_wait(&m.a_full[a_slot], a_ph);
      ROLE_ACC(rt_a);
      tc_fence_after();
      const uint64_t a_desc = a_desc0 + (m.a_smem16 + a_slot * m.a_slot16);
      const int l0 = (p.n_cph > 1) ? ((ch * p.kc) / p.cin_w) * tpc : 0;
      const uint32_t first_chunk = ch == ch0 ? 1u : 0u;
      if (RESIDENT) {
        if (elect_one_sync()) {
#pragma unroll 1
          for (int sb = 0; sb < p.msub; ++sb) {
            const uint64_t a_sb = a_desc + (uint32_t)p.sb_off[sb];
            const uint32_t d_sb = d_base + (uint32_t)(sb * p.sb_cols);
            if (TPC > 0) {
#pragma unroll
              for (int j = 0; j < TPC; ++j) {
                const IssueEnt ent = p.ilist[l0 + j];
                issue_tap<MODE, KSTEPS, PAIR>(d_sb + (uint32_t)ent.dcol, a_sb + (uint32_t)ent.a_off,
                                        b_desc0 + (b_res16 + (uint32_t)j * m.b_slot16), ic,
                                        (first_chunk & (uint32_t)ent.first) ^ 1u);
              }
            } else {
#pragma unroll 1
              for (int j = 0; j < tpc; ++j) {
                const IssueEnt ent = p.ilist[l0 + j];
                issue_tap<MODE, KSTEPS, PAIR>(d_sb + (uint32_t)ent.dcol, a_sb + (uint32_t)ent.a_off,
                                        b_desc0 + (b_res16 + (uint32_t)j * m.b_slot16), ic,
                                        (first_chunk & (uint32_t)ent.first) ^ 1u);
              }
            }
          }
          tc_commit<PAIR>(&m.a_empty[a_slot]);
        }
        __syncwarp();
        b_res16 += (uint32_t)tpc * m.b_slot16;
      } else {
#pragma unroll 1
        for (int j = 0; j < tpc; ++j) {
          const IssueEnt ent = p.ilist[l0 + j];
          mbar_wait(&m.b_full[b_slot], b_ph);
          tc_fence_after();
          if (elect_one_sync()) {
#pragma unroll 1
            for (int sb = 0; sb < p.msub; ++sb)
              issue_tap<MODE, KSTEPS, PAIR>(d_base + (uint32_t)(sb * p.sb_cols + ent.dcol),
                                      a_desc + (uint32_t)(p.sb_off[sb] + ent.a_off),
                                      b_desc0 + (m.b_region16 + b_slot * m.b_slot16), ic,
                                      (first_chunk & (uint32_t)ent.first) ^ 1u);
            tc_commit<PAIR>(&m.b_empty[b_slot]);
          }
          __syncwarp();
          if (++b_slot == (uint32_t)p.b_slots) { b_slot = 0; b_ph ^= 1; }
        }
        if (elect_one_sync()) tc_commit<PAIR>(&m.a_empty[a_slot]);
        __syncwarp();
      }
      if (++a_slot == (uint32_t)p.a_slots) { a_slot = 0; a_ph ^= 1; }
    }
    if (elect_one_sync()) tc_commit<PAIR>(&m.tm_full[acc]);
    __syncwarp();
    if (++acc == (uint32_t)p.acc_stages) { acc = 0; acc_ph ^= 1; }
  }
  ROLE_ACC(rt_issue);
  ROLE_FLUSH(2, rt_tm);
  ROLE_FLUSH(3, rt_a);
  ROLE_FLUSH(4, rt_issue);
}

// pick the unrolled form for the shapes the hot layers use; everything else runs the generic loops
template <int MODE, bool RESIDENT, bool PAIR, bool SPLITK>
__device__ __forceinline__ void mma_dispatch(const ConvKParams& p, const MmaBars& m, const IssueConsts& ic,
                                             uint64_t a_desc0, uint64_t b_desc0) {
  const int key = p.tpc * 8 + ic.ksteps;
  if (RESIDENT) {
    switch (key) {
      case 9 * 8 + 1: mma_tiles<MODE, true, 9, 1, PAIR, SPLITK>(p, m, ic, a_desc0, b_desc0); break;
      case 9 * 8 + 2: mma_tiles<MODE, true, 9, 2, PAIR, SPLITK>(p, m, ic, a_desc0, b_desc0); break;
      case 9 * 8 + 4: mma_tiles<MODE, true, 9, 4, PAIR, SPLITK>(p, m, ic, a_desc0, b_desc0); break;
      case 1 * 8 + 1: mma_tiles<MODE, true, 1, 1, PAIR, SPLITK>(p, m, ic, a_desc0, b_desc0); break;
      case 1 * 8 + 2: mma_tiles<MODE, true, 1, 2, PAIR, SPLITK>(p, m, ic, a_desc0, b_desc0); break;
      case 1 * 8 + 4: mma_tiles<MODE, true, 1, 4, PAIR, SPLITK>(p, m, ic, a_desc0, b_desc0); break;
      default: mma_tiles<MODE, true, 0, 0, PAIR, SPLITK>(p, m, ic, a_desc0, b_desc0); break;
    }
  } else {
    switch (ic.ksteps) {
      case 2: mma_tiles<MODE, false, 0, 2, PAIR, SPLITK>(p, m, ic, a_desc0, b_desc0); break;
      case 4: mma_tiles<MODE, false, 0, 4, PAIR, SPLITK>(p, m, ic, a_desc0, b_desc0); break;
      default: mma_tiles<MODE, false, 0, 0, PAIR, SPLITK>(p, m, ic, a_desc0, b_desc0); break;
    }
  }
}

template <bool PAIR>
__device__ __forceinline__ uint32_t mapa_or_local(uint64_t* bar) {
  return PAIR ? mapa_shared(smem_u32(bar), 0) : smem_u32(bar);
}
template <bool PAIR>
__device__ __forceinline__ void arrive_expect_tx(uint64_t* bar, uint32_t leader_addr, uint32_t bytes) {
  if (PAIR) mbar_arrive_expect_tx_cluster(leader_addr, bytes); else mbar_arrive_expect_tx(bar, bytes);
}
template <bool PAIR>
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* tm, uint64_t* bar, uint32_t leader_addr, int c0,
                                         int c1, int c2, int c3) {
  if (PAIR) tma_load_4d_2sm(dst, tm, leader_addr, c0, c1, c2, c3); else tma_load_4d(dst, tm, bar, c0, c1, c2, c3);
}

// scheduling unit -> tile index of THIS CTA.  Pair mode: unit u covers M tiles 2*(u / n_ntiles) + {0, 1} of N tile
// u % n_ntiles; a pair whose second M tile does not exist re-runs the last tile with every store masked.
template <bool PAIR>
__device__ __forceinline__ int sched_tile(const ConvKParams& p, int unit, uint32_t rank, bool& live) {
  live = true;
  if (!PAIR) return unit;
  const int mtp = (int)fast_div((uint32_t)unit, p.div_ntiles);
  const int nt = unit - mtp * p.n_ntiles;
  int mt = 2 * mtp + (int)rank;
  if (mt >= p.mtiles) { mt = p.mtiles - 1; live = false; }
  return mt * p.n_ntiles + nt;
}

template <int EPI, bool PAIR, bool SPLITK>  // EPI: 0 = pointwise, 1 = raw up, 2 = pointwise with the residual blend
                                           // fetched ahead, 3 = pointwise without the rare options (residuals,
                                           // blend, pooled / NCHW outputs: the generator's and conv_1's epilogue);
                                           // SPLITK: partial sums over a K range, added atomically
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_mma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ ConvKParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_smem = smem;
  uint8_t* b_smem = a_smem + p.a_slots * p.a_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + p.b_region_bytes);
  uint64_t* a_full = bars + BAR_A_FULL;
  uint64_t* a_empty = bars + BAR_A_EMPTY;
  uint64_t* tm_full = bars + BAR_TM_FULL;
  uint64_t* tm_empty = bars + BAR_TM_EMPTY;
  uint64_t* b_full = bars + BAR_B_FULL;
  uint64_t* b_empty = bars + BAR_B_EMPTY;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  float* epi_tab = reinterpret_cast<float*>(bars + BAR_COUNT + 2);   // EPI_TAB_BYTES, 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // pair mode: the "full" barriers of the LEADER CTA (rank 0) collect the TMA bytes of both CTAs (one
  // arrive.expect_tx per CTA); the leader's tm_empty collects the epilogue warps of both; "empty"/tm_full barriers
  // are local and receive the leader's multicast commits.
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const uint32_t npeers = PAIR ? 2u : 1u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_A_SLOTS; ++i) {
      mbar_init(&a_full[i], npeers);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tm_full[i], 1);
      mbar_init(&tm_empty[i], npeers * NUM_EPI_WARPS);
    }
    for (int i = 0; i < MAX_B_SLOTS; ++i) {
      mbar_init(&b_full[i], npeers);
      mbar_init(&b_empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int t_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int kc8p = (p.kc >> 3) * p.planes;

  if (warp == 0) {
    // =================================== TMA producer ===================================
    // Whole warp runs the (warp-uniform) control flow; one elected lane issues.  Keeping the flow uniform lets
    // the compiler hold addresses/descriptors in uniform registers instead of waterfall loops.
    uint32_t a_slot = 0, a_ph = 0, b_slot = 0, b_ph = 0;
    if (p.resident) {
      // weights of every (chunk, tap) stay in smem for the whole kernel: one bulk load, one barrier
      if (elect_one_sync()) {
        const uint32_t bfull0 = mapa_or_local<PAIR>(&b_full[0]);
        arrive_expect_tx<PAIR>(&b_full[0], bfull0, (uint32_t)p.b_region_bytes);
        int slot = 0;
        for (int ch = 0; ch < p.nchunks; ++ch) {
          const int cph = (ch * p.kc) / p.cin_w, cb = ch * kc8p - cph * (p.cin_w >> 3) * p.planes;
          for (int e = 0; e < p.ntaps; ++e) {
            if (p.taps[e].in_phase >= 0 && p.taps[e].in_phase != cph) continue;
            tma_load<PAIR>(b_smem + slot * p.b_slot_bytes, &tmB, &b_full[0], bfull0, 0,
                           (int)(rank * p.nsub_local) / p.b_rb, cb, p.taps[e].w_tap);
            ++slot;
          }
        }
      }
      __syncwarp();
    }
    [[maybe_unused]] long long rt_wait = 0, rt_work = 0;
    ROLE_T0();
    for (int unit = t_first; unit < p.sched_units; unit += t_step) {
      bool live;
      int ch0, ch1;
      const int su = sched_split<SPLITK>(p, unit, ch0, ch1);
      const TileCoord t = decode_tile(p, sched_tile<PAIR>(p, su, rank, live));
      for (int ch = ch0; ch < ch1; ++ch) {
        ROLE_ACC(rt_work);
        mbar_wait(&a_empty[a_slot], a_ph ^ 1);
        ROLE_ACC(rt_wait);
        if (elect_one_sync()) {
          const uint32_t afull = mapa_or_local<PAIR>(&a_full[a_slot]);
          arrive_expect_tx<PAIR>(&a_full[a_slot], afull, (uint32_t)p.a_slot_bytes);
          tma_load<PAIR>(a_smem + a_slot * p.a_slot_bytes, &tmA, &a_full[a_slot], afull, 2 * (t.x0 - 1), t.y0 - 1,
                         ch * kc8p, t.n);
        }
        __syncwarp();
        if (++a_slot == (uint32_t)p.a_slots) { a_slot = 0; a_ph ^= 1; }
        if (p.resident) continue;
        const int cph = (ch * p.kc) / p.cin_w, cb = ch * kc8p - cph * (p.cin_w >> 3) * p.planes;
        for (int e = 0; e < p.ntaps; ++e) {
          const int q = p.taps[e].phase - t.p0;
          if (q < 0 || q >= p.np) continue;
          if (p.taps[e].in_phase >= 0 && p.taps[e].in_phase != cph) continue;
          mbar_wait(&b_empty[b_slot], b_ph ^ 1);
          if (elect_one_sync()) {
            const uint32_t bfull = mapa_or_local<PAIR>(&b_full[b_slot]);
            arrive_expect_tx<PAIR>(&b_full[b_slot], bfull, (uint32_t)p.b_slot_bytes);
            tma_load<PAIR>(b_smem + b_slot * p.b_slot_bytes, &tmB, &b_full[b_slot], bfull, 0,
                           (t.co0 + (int)(rank * p.nsub_local)) / p.b_rb, cb, p.taps[e].w_tap);
          }
          __syncwarp();
          if (++b_slot == (uint32_t)p.b_slots) { b_slot = 0; b_ph ^= 1; }
        }
      }
    }
    ROLE_ACC(rt_work);
    ROLE_FLUSH(0, rt_wait);
    ROLE_FLUSH(1, rt_work);
  } else if (warp == 1) {
    // =================================== MMA issuer ======================================
    // Warp-uniform control flow with elected issue; every MMA operand derives from provably uniform sources (kernel
    // parameters in the constant bank, uniform counters, a shuffled TMEM base) -- see mma_tiles.
    const uint32_t a_lbo = p.planes * p.patch_bytes, a_sbo = p.pw * 16;
    const uint32_t nb = p.nsub_local * 16;  // bytes of one (k-group, plane) slab of the B block held by this CTA
    const uint32_t b_lbo = p.planes * nb, b_sbo = 128;
    const uint64_t a_desc0 = make_smem_desc(0, a_lbo, a_sbo), b_desc0 = make_smem_desc(0, b_lbo, b_sbo);
    IssueConsts ic;
    ic.idesc = make_idesc_bf16(p.nsub, PAIR ? 256 : 128);
    ic.idesc2 = make_idesc_bf16(2 * p.nsub, PAIR ? 256 : 128);
    ic.a_kstep = (2 * a_lbo) >> 4;   // descriptor start-address increments per K step (16-byte units)
    ic.b_kstep = (2 * b_lbo) >> 4;
    ic.a_lo16 = (uint32_t)p.patch_bytes >> 4;
    ic.b_lo16 = nb >> 4;
    ic.ksteps = p.kc >> 4;
    MmaBars mb;
    mb.a_full = a_full; mb.a_empty = a_empty; mb.b_full = b_full; mb.b_empty = b_empty;
    mb.tm_full = tm_full; mb.tm_empty = tm_empty;
    mb.a_smem16 = smem_u32(a_smem) >> 4;
    mb.b_region16 = smem_u32(b_smem) >> 4;
    mb.a_slot16 = (uint32_t)p.a_slot_bytes >> 4;
    mb.b_slot16 = (uint32_t)p.b_slot_bytes >> 4;
    mb.tm_base = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int mma_mode = p.planes == 2 ? (p.stack ? 2 : 1) : 0;
    if (PAIR) {
      // only the leader CTA issues (its MMAs drive both SMs); pairs are never combined with the stacked mode
      if (rank == 0) {
        if (p.resident) {
          if (mma_mode == 1) mma_dispatch<1, true, true, SPLITK>(p, mb, ic, a_desc0, b_desc0);
          else mma_dispatch<0, true, true, SPLITK>(p, mb, ic, a_desc0, b_desc0);
        } else {
          if (mma_mode == 1) mma_dispatch<1, false, true, SPLITK>(p, mb, ic, a_desc0, b_desc0);
          else mma_dispatch<0, false, true, SPLITK>(p, mb, ic, a_desc0, b_desc0);
        }
      }
    } else if (p.resident) {
      if (mma_mode == 2) mma_dispatch<2, true, false, SPLITK>(p, mb, ic, a_desc0, b_desc0);
      else if (mma_mode == 1) mma_dispatch<1, true, false, SPLITK>(p, mb, ic, a_desc0, b_desc0);
      else mma_dispatch<0, true, false, SPLITK>(p, mb, ic, a_desc0, b_desc0);
    } else {
      if (mma_mode == 2) mma_dispatch<2, false, false, SPLITK>(p, mb, ic, a_desc0, b_desc0);
      else if (mma_mode == 1) mma_dispatch<1, false, false, SPLITK>(p, mb, ic, a_desc0, b_desc0);
      else mma_dispatch<0, false, false, SPLITK>(p, mb, ic, a_desc0, b_desc0);
    }
  } else {
    // =================================== epilogue ========================================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int m = quarter * 32 + lane;
    const int ty = m >> 3, tx = m & 7;
    const int et = (int)threadIdx.x - 64;   // index among the epilogue threads
    const int wg = (warp - 2) >> 2;         // 0 / 1: which half of the tile's units this warp handles
    int cur_n = -1, cur_co0 = -1;
    uint32_t acc = 0, acc_ph = 0;
    [[maybe_unused]] long long rt_full = 0, rt_proc = 0;
    ROLE_T0();
    const int nsq = p.msub * p.np;                  // (M block, phase) pairs of a tile, each cw columns wide
    const bool split_sq = nsq >= 2;
    const int c_step = split_sq ? 16 : 32, sq_step = split_sq ? 2 : 1, c_start = split_sq ? 0 : 16 * wg;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // The per-pixel noise values are fetched one tile ahead: their global-load latency (L2 / HBM) would otherwise sit
    // on the per-tile critical path of the epilogue warps.
    float nz_next[4] = {0.f, 0.f, 0.f, 0.f};
    auto fetch_noise = [&](int unit_idx) {
      bool live_n;
      int cu0, cu1;
      const TileCoord tn = decode_tile(p, sched_tile<PAIR>(p, sched_split<SPLITK>(p, unit_idx, cu0, cu1), rank, live_n));
#pragma unroll
      for (int sb = 0; sb < 4; ++sb) {
        nz_next[sb] = 0.f;
        if (sb < p.msub) {
          const int y = tn.y0 + p.sb_y[sb] + ty, x = tn.x0 + p.sb_x[sb] + tx;
          if (y < p.dom_h && x < p.dom_w)
            nz_next[sb] = __ldg(p.noise + (size_t)tn.n * p.noise_bstride + (size_t)y * p.W + x);
        }
      }
    };
    const bool want_noise = EPI != 1 && !SPLITK && p.noise;
    if (want_noise && t_first < p.sched_units) fetch_noise(t_first);
    const uint32_t tm_empty_leader = PAIR ? mapa_shared(smem_u32(&tm_empty[0]), 0) : 0u;
    for (int unit = t_first; unit < p.sched_units; unit += t_step) {
      bool live;
      int cu0, cu1;
      const TileCoord t = decode_tile(p, sched_tile<PAIR>(p, sched_split<SPLITK>(p, unit, cu0, cu1), rank, live));
      if (EPI != 1 && !SPLITK && (t.n != cur_n || t.co0 != cur_co0)) {
        // per-(sample, N tile) parameter table: one cooperative reload when the sample changes (tiles are visited
        // in sample order), then every per-channel parameter is a broadcast LDS.128 instead of a global load
        epi_bar_sync();
        epi_fill_table(p, epi_tab, t.n, t.co0, et);
        epi_bar_sync();
        cur_n = t.n;
        cur_co0 = t.co0;
      }
      float nz_cur[4];
#pragma unroll
      for (int sb = 0; sb < 4; ++sb) nz_cur[sb] = nz_next[sb];
      if (want_noise && unit + t_step < p.sched_units) fetch_noise(unit + t_step);
      ROLE_ACC(rt_proc);
      mbar_wait_relaxed(&tm_full[acc], acc_ph);
      ROLE_ACC(rt_full);
      tc_fence_after();
      const uint32_t taddr = lane_base + acc * p.tm_stride;
      // Units = (M block / phase pair sq, 16-column group c), walked in TMEM column order.  The TMEM load of unit u+1 is
      // in flight while unit u is processed (also across M blocks: a 16-channel layer has ONE group per block, and the
      // exposed tcgen05.ld latency was a third of its tile time).  Stacked mode: the lo-product half `sx` is folded
      // into the group right after its wait, then reloaded.
      EpiPix px;
      px.n = t.n;
      px.y = 0; px.x = 0; px.valid = false; px.nz = 0.f;
      float rgb[3] = {0.f, 0.f, 0.f};
      // the two warps of a lane quarter split the units: by (M block, phase) parity when a tile has several, else by
      // column-group parity
      int u_sq = split_sq ? wg : 0, u_c = split_sq ? 0 : 16 * wg;
      int cur_sb = -1;
      uint32_t r0[16], r1[16], sx[16];
      const bool any = u_sq < nsq && u_c < p.cw;
      // same-resolution residual blend (E.py:81-84 with the pooled conv_2 output): its 2 x 32 bytes per unit are
      // requested ONE UNIT AHEAD into the (otherwise unused) `sx` registers -- fetched at the point of use, their
      // L2/HBM latency was the critical path of the 1x1 residual convs
      constexpr bool blend_pf = EPI == 2;   // (own instantiation: the prefetch registers must not weigh on the other paths)
      auto blend_fetch = [&](int sq, int c) {
        const int sbn = sq >> p.np_shift;
        const int y = t.y0 + p.sb_y[sbn] + ty, x = t.x0 + p.sb_x[sbn] + tx;
        if (live && y < p.dom_h && x < p.dom_w) {
          const int C8 = p.Cout >> 3;
          float tmp[16];
          load8_f32b(p.blend_src, f32b_idx32(t.n, ((t.co0 + c) >> 3), y, x, C8, p.H, p.W), tmp);
          load8_f32b(p.blend_src, f32b_idx32(t.n, ((t.co0 + c) >> 3) + 1, y, x, C8, p.H, p.W), tmp + 8);
#pragma unroll
          for (int j = 0; j < 16; ++j) sx[j] = __float_as_uint(tmp[j]);
        }
      };
      if (any) {
        tm_ld16_issue(taddr + u_sq * p.acc_cols + u_c, r0);
        if (p.stack) tm_ld16_issue(taddr + u_sq * p.acc_cols + u_c + p.cw, sx);
        if (blend_pf) blend_fetch(u_sq, u_c);
      }
      auto step = [&](uint32_t (&bc)[16], uint32_t (&bn)[16]) -> bool {
        tm_ld16_wait(bc);
        int n_sq = u_sq, n_c = u_c + c_step;
        if (n_c >= p.cw) { n_c = c_start; n_sq += sq_step; }
        const bool n_valid = n_sq < nsq;
        const uint32_t n_addr = taddr + n_sq * p.acc_cols + n_c;
        if (n_valid) tm_ld16_issue(n_addr, bn);
        if (p.stack) {
          tm_ld16_wait(sx);   // (waits for every outstanding load; bn is simply complete early)
          epi_fold16(bc, sx);
          if (n_valid) tm_ld16_issue(n_addr + p.cw, sx);
        }
        const int q = u_sq & (p.np - 1);
        const int sb = u_sq >> p.np_shift;
        if (sb != cur_sb) {                 // first unit of an M block: its pixel
          cur_sb = sb;
          px.y = t.y0 + p.sb_y[sb] + ty;
          px.x = t.x0 + p.sb_x[sb] + tx;
          px.valid = live && (px.y < p.dom_h) && (px.x < p.dom_w);
          px.nz = sb == 0 ? nz_cur[0] : (sb == 1 ? nz_cur[1] : (sb == 2 ? nz_cur[2] : nz_cur[3]));
        }
        if (blend_pf) {
          float bl[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) bl[j] = __uint_as_float(sx[j]);
          if (n_valid) blend_fetch(n_sq, n_c);
          epi_group16<EPI, SPLITK>(p, epi_tab, px, t.co0, t.p0, q, u_c, bc, rgb, bl);
        } else {
          epi_group16<EPI, SPLITK>(p, epi_tab, px, t.co0, t.p0, q, u_c, bc, rgb);
        }
        if (EPI != 1 && p.rgb_w && (!n_valid || (n_sq >> p.np_shift) != sb)) {   // last unit of an M block
          if (px.valid) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
              atomicAdd(p.rgb_out + (((size_t)px.n * 3 + ch) * p.H + px.y) * p.W + px.x, rgb[ch]);
          }
          rgb[0] = rgb[1] = rgb[2] = 0.f;
          __syncwarp();
        }
        u_sq = n_sq;
        u_c = n_c;
        return n_valid;
      };
      while (any) {
        if (!step(r0, r1)) break;
        if (!step(r1, r0)) break;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(tm_empty_leader + acc * 8u); else mbar_arrive(&tm_empty[acc]);
      }
      if (++acc == (uint32_t)p.acc_stages) { acc = 0; acc_ph ^= 1; }
    }
    ROLE_ACC(rt_proc);
    if (warp == 2) {
      ROLE_FLUSH(5, rt_full);
      ROLE_FLUSH(6, rt_proc);
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                   : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// split-K finish: the complete (generic) pointwise epilogue over the summed scratch map.  One thread per
// (sample, pixel, 16-channel group); with out_pool one thread per POOLED pixel (it evaluates the four source pixels).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_splitk_finish_kernel(const __grid_constant__ ConvKParams p) {
  const int C16 = p.Cout >> 4, C8 = p.Cout >> 3;
  const int Ho = p.out_pool ? p.H >> 1 : p.H, Wo = p.out_pool ? p.W >> 1 : p.W;
  const size_t total = (size_t)p.N * C16 * Ho * Wo;
  const size_t HW = (size_t)p.H * p.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    size_t tt = i / Wo;
    const int y = (int)(tt % Ho);
    tt /= Ho;
    const int g16 = (int)(tt % C16);
    const int n = (int)(tt / C16);
    const int c0 = g16 * 16;
    auto load_px = [&](int yy, int xx, float* v, PixelCtx& px) {
      px.n = n; px.y = yy; px.x = xx; px.valid = true;
      px.nz = p.noise ? __ldg(p.noise + (size_t)n * p.noise_bstride + (size_t)yy * p.W + xx) : 0.f;
      load8_f32b(p.ws, ((size_t)n * C8 + (c0 >> 3)) * HW + (size_t)yy * p.W + xx, v);
      load8_f32b(p.ws, ((size_t)n * C8 + (c0 >> 3) + 1) * HW + (size_t)yy * p.W + xx, v + 8);
    };
    if (p.out_pool) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          float v[16];
          PixelCtx px;
          load_px(2 * y + dy, 2 * x + dx, v, px);
          epi_value16(p, px, c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += v[j];
        }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] *= 0.25f;
      const size_t HWp = (size_t)Ho * Wo;
      const size_t o = ((size_t)n * C8 + (c0 >> 3)) * HWp + (size_t)y * Wo + x;
      store8_f32b(p.out_f32b, o, acc);
      store8_f32b(p.out_f32b, o + HWp, acc + 8);
    } else {
      float v[16], rgb[3] = {0.f, 0.f, 0.f};
      PixelCtx px;
      load_px(y, x, v, px);
      epi_pointwise16(p, px, c0, v, rgb);
      if (p.rgb_w) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) atomicAdd(p.rgb_out + (((size_t)n * 3 + ch) * p.H + y) * p.W + x, rgb[ch]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// checker kernel: same tiling / packing / epilogue, CUDA-core FMAs straight from global memory.
// Test-only (DGE_CONV_FLAG_CHECKER): isolates TMA/UMMA-descriptor bugs from packing/epilogue bugs.
// ---------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(128) conv_checker_kernel(const __grid_constant__ ConvKParams p) {
  const int m = threadIdx.x, ty = m >> 3, tx = m & 7;
  const int C8in = p.Cin >> 3;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const TileCoord t = decode_tile(p, tile);
    PixelCtx px;
    px.n = t.n;
    px.y = t.y0 + ty;
    px.x = t.x0 + tx;
    px.valid = (px.y < p.dom_h) && (px.x < p.dom_w);
    px.nz = 0.f;
    if (EPI != 1 && p.noise && px.valid)
      px.nz = __ldg(p.noise + (size_t)px.n * p.noise_bstride + (size_t)px.y * p.W + px.x);
    float rgb[3] = {0.f, 0.f, 0.f};
    for (int q = 0; q < p.np; ++q) {
      for (int c = 0; c < p.cw; c += 16) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
        for (int e = 0; e < p.ntaps; ++e) {
          if (p.taps[e].phase != t.p0 + q) continue;
          const int dy = p.taps[e].a_off / p.pw, dx = p.taps[e].a_off % p.pw;
          const int iy = px.y - 1 + dy, ix = px.x - 1 + dx;
          if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;
          const int C8w = p.cin_w >> 3;
          for (int gw = 0; gw < C8w; ++gw) {
            const int g = p.taps[e].in_phase < 0 ? gw : p.taps[e].in_phase * C8w + gw;
            float ah[8], al[8];
            const uint4* xa = reinterpret_cast<const uint4*>(p.x);
            unpack8(__ldg(xa + act_idx16(px.n, g, 0, iy, ix, C8in, p.planes, p.H, p.W)), ah);
            if (p.planes == 2) unpack8(__ldg(xa + act_idx16(px.n, g, 1, iy, ix, C8in, p.planes, p.H, p.W)), al);
            for (int j = 0; j < 16; ++j) {
              const int co = t.co0 + c + j;
              const uint4* wb = reinterpret_cast<const uint4*>(p.wpk);
              // WPK [tap][Cin/8][planes][Cout][8]
              const size_t wi = (((size_t)p.taps[e].w_tap * C8w + gw) * p.planes) * p.Cout + co;
              float wh[8], wl[8];
              unpack8(__ldg(wb + wi), wh);
              float acc = v[j];
#pragma unroll
              for (int k = 0; k < 8; ++k) acc = fmaf(ah[k], wh[k], acc);
              if (p.planes == 2) {
                unpack8(__ldg(wb + wi + p.Cout), wl);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = fmaf(ah[k], wl[k], fmaf(al[k], wh[k], acc));
              }
              v[j] = acc;
            }
          }
        }
        if (EPI != 1)
          epi_pointwise16(p, px, t.co0 + c, v, rgb);
        else
          epi_rawup16<false>(p, px.n, px.y, px.x, px.valid, t.p0 + q, t.co0 + c, v);
      }
    }
    if (EPI != 1 && p.rgb_w && px.valid) {
      for (int ch = 0; ch < 3; ++ch)
        atomicAdd(p.rgb_out + (((size_t)px.n * 3 + ch) * p.H + px.y) * p.W + px.x, rgb[ch]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// Tensor maps view the 16-byte channel groups as pairs of uint64 so that a whole pixel row of a patch is ONE
// contiguous inner-dimension run (box inner extent = 2*pixels elements, <= 256).
int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return DGE_ERR_CUDA;
  }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu %llu %llu box %u %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0], box[1],
              box[2]);
    return DGE_ERR_CUDA;
  }
  return DGE_OK;
}

static int largest_div(int value, int cap, int step) {
  for (int c = (cap / step) * step; c >= step; c -= step)
    if (value % c == 0) return c;
  return 0;
}

static int g_num_sms = 0;

// Split-K is used for the small maps (<= 16x16 at batch 8): so few output tiles that most SMs would idle while each
// CTA walks the full serial K chain (9 taps x 512 channels = 288 K steps, ~26 us).  The K chunks are divided over up to 8
// CTAs per tile; partial sums are added to a zeroed fp32 scratch map and a small finish kernel applies the epilogue.
// (The transposed conv's raw map is linear in the accumulator, so its parts are added in place -- no scratch.)
static bool splitk_wanted(const dge_conv_args* a, int cw_full, int num_sms) {
  static int off = -1;
  if (off < 0) off = getenv("DGE_NO_SPLITK") ? 1 : 0;
  if (off || (a->flags & DGE_CONV_FLAG_CHECKER)) return false;
  const bool up = a->kind == DGE_CONV_UP3X3;
  const int dom_h = up ? a->h + 1 : a->h, dom_w = up ? a->w + 1 : a->w;
  const long long mt = (long long)a->n * ((dom_h + BH - 1) / BH) * ((dom_w + BW - 1) / BW);
  const long long ctas = mt * (a->cout / cw_full);
  const int cin_w = a->kind == DGE_CONV_DOWN4X4S2 ? a->cin / 4 : a->cin;
  const int taps = a->kind == DGE_CONV_1X1 ? 1 : (a->kind == DGE_CONV_DOWN4X4S2 ? 4 : 9);
  return ctas * 2 <= num_sms && (long long)taps * (cin_w / 16) >= 64 && a->cin >= 256;
}
static int conv_cw_full(const dge_conv_args* a) {
  if (a->kind == DGE_CONV_UP3X3) return largest_div(a->cout, 128, 16);
  return (a->cout % 256 == 0) ? 256 : largest_div(a->cout, 128, 16);
}

int conv_forward(const dge_conv_args* a, cudaStream_t stream) {
  DGE_REQUIRE(a != nullptr, "conv: null args");
  DGE_REQUIRE(a->kind >= 0 && a->kind <= 3, "conv: bad kind %d", a->kind);
  DGE_REQUIRE(a->n > 0 && a->h > 0 && a->w > 0, "conv: bad dims n=%d h=%d w=%d", a->n, a->h, a->w);
  DGE_REQUIRE(a->cin >= 16 && a->cin % 16 == 0, "conv: cin=%d must be a positive multiple of 16", a->cin);
  DGE_REQUIRE(a->cout >= 16 && a->cout % 16 == 0, "conv: cout=%d must be a positive multiple of 16", a->cout);
  DGE_REQUIRE(a->planes == 1 || a->planes == 2, "conv: planes=%d must be 1 or 2", a->planes);
  DGE_REQUIRE(a->x && a->wpk, "conv: null x / wpk");
  DGE_REQUIRE((a->in_h == 0 || a->in_h >= a->h) && (a->in_w == 0 || a->in_w >= a->w) &&
                  ((a->in_h == 0 && a->in_w == 0) || (a->kind != DGE_CONV_UP3X3 && !(a->flags & DGE_CONV_FLAG_CHECKER))),
              "conv: in_h/in_w (%d, %d) must cover the output domain (tcgen05 path, not the transposed conv)", a->in_h,
              a->in_w);
  const bool up = a->kind == DGE_CONV_UP3X3;
  if (up) {
    DGE_REQUIRE(a->out_raw_up != nullptr, "conv: UP3X3 needs out_raw_up");
  } else {
    DGE_REQUIRE(a->out_act || a->out_f32b || a->out_nchw || a->rgb_out, "conv: no output given");
    DGE_REQUIRE(!a->out_act || a->out_planes == 1 || a->out_planes == 2, "conv: out_planes=%d", a->out_planes);
    DGE_REQUIRE(!a->rgb_w == !a->rgb_out, "conv: rgb_w and rgb_out must be given together");
    DGE_REQUIRE(!a->noise_w || a->noise, "conv: noise_w without noise");
    DGE_REQUIRE(!a->out_pool || (a->out_f32b && !a->out_act && !a->out_nchw && !a->rgb_out && !a->blend_src &&
                                 a->h % 2 == 0 && a->w % 2 == 0 && !(a->flags & DGE_CONV_FLAG_CHECKER)),
                "conv: out_pool needs even h/w and out_f32b as the only output (tcgen05 path)");
    DGE_REQUIRE(!a->preact_add || ((a->preact_c == 0 || (a->preact_c >= a->cout && a->preact_c % 8 == 0)) &&
                                   (a->preact_up != 2 || (a->h % 2 == 0 && a->w % 2 == 0))),
                "conv: bad preact residual shape (preact_c=%d preact_up=%d)", a->preact_c, a->preact_up);
  }

  ConvKParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->n; p.H = a->h; p.W = a->w;
  for (int i = 0; i < 16; ++i) p.taps[i].in_phase = -1;
  p.Cin = a->cin; p.Cout = a->cout;
  p.cin_w = a->cin;
  p.planes = a->planes;
  p.P = up ? 4 : 1;
  p.dom_h = up ? a->h + 1 : a->h;
  p.dom_w = up ? a->w + 1 : a->w;
  p.tiles_x = (p.dom_w + BW - 1) / BW;   // (base 16x8 blocks; re-derived once the tile shape is chosen)
  p.tiles_y = (p.dom_h + BH - 1) / BH;
  // N tiling
  const int ntot = p.P * p.Cout;
  if (up) {
    // all 4 output phases of a tile share one A patch: <=128 channels x 4 phases = <=512 TMEM columns
    p.cw = largest_div(p.Cout, 128, 16);
    p.np = 4;
    p.ntile = 4 * p.cw;
  } else {
    // one N=256 MMA per K-step when possible: A is then fetched from smem once per 256 columns
    p.cw = (p.Cout % 256 == 0) ? 256 : largest_div(p.Cout, 128, 16);
    p.ntile = p.cw;
    p.np = 1;
  }
  bool want_sk = false;
  // small problems (<= 32x32 maps): shrink the N tile until the grid covers the SMs -- a 128x256 tile over K = 9*512 is
  // ~100 us of MMA time on ONE SM, so 16..32 such tiles would leave most of the chip idle
  {
    if (g_num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
      if (g_num_sms <= 0) g_num_sms = 148;
    }
    const long long mtiles = (long long)p.N * p.tiles_x * p.tiles_y;
    want_sk = (up || a->splitk_ws != nullptr) && splitk_wanted(a, p.cw, g_num_sms);
    if (!want_sk)
      while (mtiles * (p.Cout / p.cw) * 4 < (long long)g_num_sms * 3 && p.cw >= 64 && (p.cw / 2) % 16 == 0) p.cw /= 2;
    p.ntile = p.np * p.cw;
  }
  p.nsub = p.cw;   // one MMA covers the whole column block (N <= 256)
  DGE_REQUIRE(p.cw > 0 && p.nsub > 0, "conv: cannot tile cout=%d", p.Cout);
  p.n_ntiles = ntot / p.ntile;
  DGE_REQUIRE(p.n_ntiles * p.ntile == ntot, "conv: cannot tile cout=%d", p.Cout);
  DGE_REQUIRE(p.np * (p.cw / p.nsub) <= 8 && p.cw / p.nsub <= 2, "conv: internal sub-tile bookkeeping overflow");
  // hi|lo stacking on N pays while the MMA is bound by its A-operand read (2*cw <= 96 columns)
  {
    static int stack_max = -1;
    if (stack_max < 0) stack_max = getenv("DGE_STACK_MAX") ? atoi(getenv("DGE_STACK_MAX")) : 64;   // A/B switch
    const int lim = p.np == 1 ? stack_max : 48;   // (4-phase tiles: 2*cw*4 columns must leave room for two stages)
    p.stack = (p.planes == 2 && p.cw <= lim && p.np * 2 * p.cw <= 512) ? 1 : 0;
    // the 1x1 residual convs are bound by their blend loads, not by MMAs: keep `sx` free for the blend prefetch
    if (a->kind == DGE_CONV_1X1 && a->blend_src && !a->blend_pool) p.stack = 0;
  }
  p.acc_cols = p.cw * (p.stack ? 2 : 1);
  p.sb_cols = p.np * p.acc_cols;
  p.np_shift = p.np == 4 ? 2 : 0;
  const int bar_bytes = (BAR_COUNT + 2) * 8 + EPI_TAB_BYTES;   // barriers + TMEM slot, epilogue parameter table
  const int smem_cap = 224 * 1024;
  const int taps_per_chunk = (a->kind == DGE_CONV_DOWN4X4S2) ? 4 : (a->kind == DGE_CONV_1X1 ? 1 : 9);
  const long long b_all = (long long)taps_per_chunk * (p.Cin / 8) * p.planes * p.cw * 16;   // every (tap, channel) pair in use
  // tile shape: 4 or 2 M blocks per tile when the weights stay resident next to two (larger) patch slots, both
  // accumulator stages still fit TMEM, and there are enough tiles to balance the SMs
  p.msub = 1;
  if (!(a->flags & DGE_CONV_FLAG_CHECKER) && p.Cout == p.cw) {
    const int cands[2] = {4, 2};
    for (int ci = 0; ci < 2 && p.msub == 1; ++ci) {
      const int ms = cands[ci], th = ms == 4 ? 2 * BH : BH, tw = 2 * BW;
      if (p.dom_h < th || p.dom_w < tw) continue;
      if (2 * ms * p.sb_cols > 512) continue;
      const long long tiles = (long long)p.N * ((p.dom_h + th - 1) / th) * ((p.dom_w + tw - 1) / tw);
      if (tiles < 4ll * g_num_sms) continue;
      const long long patch = (long long)(th + 2) * (tw + 2) * 16;
      const int kcs[3] = {64, 32, 16};
      for (int i = 0; i < 3; ++i) {
        const int kc = largest_div(a->kind == DGE_CONV_DOWN4X4S2 ? p.Cin / 4 : p.Cin, kcs[i], 16);
        if (kc > 0 && b_all + 2 * (kc / 8) * p.planes * patch + bar_bytes <= smem_cap) {
          p.msub = ms;
          break;
        }
      }
    }
  }
  p.tile_h = p.msub == 4 ? 2 * BH : BH;
  p.tile_w = p.msub >= 2 ? 2 * BW : BW;
  p.pw = p.tile_w + 2;
#ifdef DGE_X_ALIGNED   // timing experiment only (wrong results): 128-byte aligned core matrices for every tap
  p.pw = p.tile_w + 8;
#endif
  p.ph = p.tile_h + 2;
  p.patch_bytes = p.ph * p.pw * 16;
  for (int sb = 0; sb < 4; ++sb) {
    const int sby = p.msub == 4 ? (sb >> 1) : 0, sbx = p.msub == 4 ? (sb & 1) : sb;
    p.sb_y[sb] = sby * BH;
    p.sb_x[sb] = sbx * BW;
    p.sb_off[sb] = sby * BH * p.pw + sbx * BW;
  }
  p.tiles_x = (p.dom_w + p.tile_w - 1) / p.tile_w;
  p.tiles_y = (p.dom_h + p.tile_h - 1) / p.tile_h;
  // CTA pairs (cta_group::2) for the wide column blocks: each CTA keeps half of the weight rows, so the weight operand
  // costs half the shared-memory bandwidth per SM (the limiter of the N >= 128 MMAs) and half the L2->smem traffic
  p.mtiles = p.N * p.tiles_x * p.tiles_y;
  {
    static int no_pair = -1;
    if (no_pair < 0) no_pair = getenv("DGE_NO_PAIR") ? 1 : 0;   // A/B switch for experiments
    static int pair_min = -1;
    if (pair_min < 0) pair_min = getenv("DGE_PAIR_MIN") ? atoi(getenv("DGE_PAIR_MIN")) : 64;
    // (measured: pairs lose on the narrow stacked / multi-block tiles -- E 16->16 @1024^2 0.29 -> 0.49 ms -- so they
    //  are used only where the weight operand is a large share of the shared-memory traffic)
    p.pair = (!no_pair && !(a->flags & DGE_CONV_FLAG_CHECKER) && !p.stack && p.msub == 1 && p.cw >= pair_min &&
              p.mtiles >= 2 && g_num_sms >= 2) ? 1 : 0;
  }
  p.nsub_local = p.pair ? p.cw / 2 : p.cw;
  const long long b_all_cta = p.pair ? b_all / 2 : b_all;   // weight bytes one CTA holds when resident
  // taps (patch offsets need the patch width of the chosen tile shape)
  if (a->kind == DGE_CONV_3X3) {
    p.ntaps = 9;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        TapEntry& t = p.taps[ky * 3 + kx];
        t.a_off = (int16_t)(ky * p.pw + kx);
#ifdef DGE_X_ALIGNED
        t.a_off = (int16_t)(ky * p.pw);
#endif
        t.w_tap = (int16_t)(ky * 3 + kx);
        t.phase = 0;
      }
  } else if (a->kind == DGE_CONV_1X1) {
    p.ntaps = 1;
    p.taps[0].a_off = (int16_t)(1 * p.pw + 1);
    p.taps[0].w_tap = 0;
    p.taps[0].phase = 0;
  } else if (a->kind == DGE_CONV_DOWN4X4S2) {
    // out[Y][X] = sum_{ky,kx<4} W4[ky][kx] * xin[2Y+ky-1][2X+kx-1]; xin is given space-to-depth: channel block
    // ph = 2*(row parity)+(col parity) holds xin[2y+py][2x+px].  Row 2Y+ky-1 -> parity (ky+1)%2, offset floor((ky-1)/2).
    DGE_REQUIRE(a->cin % 64 == 0, "conv: DOWN4X4S2 needs cin (=4*C) with C %% 16 == 0");
    p.cin_w = a->cin / 4;
    p.ntaps = 16;
    static const int off[4] = {-1, 0, 0, 1};
    for (int ky = 0; ky < 4; ++ky)
      for (int kx = 0; kx < 4; ++kx) {
        TapEntry& t = p.taps[ky * 4 + kx];
        t.a_off = (int16_t)((1 + off[ky]) * p.pw + (1 + off[kx]));
        t.w_tap = (int16_t)(ky * 4 + kx);
        t.phase = 0;
        t.in_phase = (int16_t)(2 * ((ky + 1) % 2) + ((kx + 1) % 2));
      }
  } else {
    // t[2Y+ky'][2X+kx'] += x[Y - a][X - b] * Wf[ky][kx],  ky = ky' + 2a  (ky' = ky%2, a = ky/2)
    p.ntaps = 9;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        TapEntry& t = p.taps[ky * 3 + kx];
        const int dy = 1 - ky / 2, dx = 1 - kx / 2;  // patch offset of input pixel (Y - ky/2, X - kx/2)
        t.a_off = (int16_t)(dy * p.pw + dx);
        t.w_tap = (int16_t)(ky * 3 + kx);
        t.phase = (int16_t)(2 * (ky % 2) + (kx % 2));
      }
  }
  p.tm_stride = p.msub * p.sb_cols;
  p.acc_stages = (2 * p.tm_stride <= 512) ? 2 : 1;   // single-buffered accumulator when the tile fills TMEM
  int cols = 32;
  while (cols < p.acc_stages * p.tm_stride) cols *= 2;
  DGE_REQUIRE(cols <= 512, "conv: TMEM overflow");
  p.tmem_cols = cols;
  const int tmem_occ = 512 / p.tmem_cols;
  // K chunking + shared-memory plan
  DGE_REQUIRE(taps_per_chunk == ((a->kind == DGE_CONV_DOWN4X4S2) ? 4 : p.ntaps), "conv: internal tap count mismatch");
  p.tpc = taps_per_chunk;
  // issue order of the taps of one K chunk (the TMA producer loads the weight slabs in the same order)
  p.n_cph = (a->kind == DGE_CONV_DOWN4X4S2) ? 4 : 1;
  for (int cph = 0; cph < p.n_cph; ++cph) {
    int j = 0;
    unsigned seen = 0;
    for (int e = 0; e < p.ntaps; ++e) {
      if (p.taps[e].in_phase >= 0 && p.taps[e].in_phase != cph) continue;
      IssueEnt& ie = p.ilist[cph * p.tpc + j];
      ie.a_off = p.taps[e].a_off;
      ie.dcol = (int16_t)(p.taps[e].phase * p.acc_cols);
      ie.first = !((seen >> p.taps[e].phase) & 1u) ? 1 : 0;   // (consulted only for the first chunk of a K range)
      ie.w_tap = p.taps[e].w_tap;
      seen |= 1u << p.taps[e].phase;
      ++j;
    }
    DGE_REQUIRE(j == p.tpc, "conv: internal tap list mismatch");
  }
  p.resident = 0;
  if (p.n_ntiles == 1) {
    // weights resident in smem for the whole kernel when they fit next to >= 2 patch slots
    const int kcs[3] = {64, 32, 16};
    for (int i = 0; i < 3 && !p.resident; ++i) {
      const int kc = largest_div(p.cin_w, kcs[i], 16);
      const long long a_slot = (long long)(kc / 8) * p.planes * p.patch_bytes;
      if (b_all_cta + 2 * a_slot + bar_bytes <= smem_cap) {
        p.resident = 1;
        p.kc = kc;
      }
    }
  }
  if (!p.resident) p.kc = largest_div(p.cin_w, p.cw > 128 ? 32 : 64, 16);
  p.nchunks = p.Cin / p.kc;
  p.a_slot_bytes = (p.kc / 8) * p.planes * p.patch_bytes;
  p.b_sub_bytes = (p.kc / 8) * p.planes * p.nsub_local * 16;
  p.b_slot_bytes = p.b_sub_bytes;
  size_t smem = 0;
  (void)tmem_occ;
  int max_occ = 1;   // one CTA (2 + 8 warps) per SM
  if (p.resident) {
    p.b_region_bytes = (int)b_all_cta;
    p.b_slots = 1;
    const int budget = smem_cap;
    p.a_slots = (int)((budget - b_all_cta - bar_bytes) / p.a_slot_bytes);
    if (p.a_slots > MAX_A_SLOTS) p.a_slots = MAX_A_SLOTS;
  } else {
    p.a_slots = 2;
    const int budget = 200 * 1024;
    p.b_slots = (budget - p.a_slots * p.a_slot_bytes) / p.b_slot_bytes;
    if (p.b_slots > MAX_B_SLOTS) p.b_slots = MAX_B_SLOTS;
    DGE_REQUIRE(p.b_slots >= 2, "conv: smem budget too small for this shape (b_slot=%d)", p.b_slot_bytes);
    p.b_region_bytes = p.b_slots * p.b_slot_bytes;
  }
  DGE_REQUIRE(p.a_slots >= 2, "conv: smem plan failed (a_slots=%d)", p.a_slots);
  smem = (size_t)p.a_slots * p.a_slot_bytes + (size_t)p.b_region_bytes + bar_bytes;
  if (smem > 113 * 1024) max_occ = 1;
  p.total_tiles = p.N * p.tiles_x * p.tiles_y * p.n_ntiles;
  p.sched_tiles = p.pair ? ((p.mtiles + 1) / 2) * p.n_ntiles : p.total_tiles;
  p.ksplit = 1;
  if (want_sk) {
    const int ctas = p.pair ? 2 * p.sched_tiles : p.sched_tiles;
    int ks = g_num_sms / (ctas > 0 ? ctas : 1);
    if (ks > 8) ks = 8;
    if (ks > p.nchunks / 2) ks = p.nchunks / 2;
    if (ks >= 2) p.ksplit = ks;
  }
  p.sched_units = p.sched_tiles * p.ksplit;
  p.div_ksplit = make_fast_div((uint32_t)p.ksplit);
  p.ws = a->splitk_ws;
  DGE_REQUIRE((long long)p.N * p.tiles_x * p.tiles_y * p.n_ntiles < (1ll << 31), "conv: too many tiles");
  p.div_ntiles = make_fast_div((uint32_t)p.n_ntiles);
  p.div_per_img = make_fast_div((uint32_t)(p.tiles_x * p.tiles_y));
  p.div_tiles_x = make_fast_div((uint32_t)p.tiles_x);
  p.act_mode = (a->slope == 1.f) ? 0 : ((a->slope >= 0.f && a->slope <= 1.f) ? 1 : 2);

  p.demod = a->demod; p.noise = a->noise; p.noise_bstride = a->noise_bstride; p.noise_w = a->noise_w;
  p.noise_scalar = a->noise_scalar; p.bias = a->bias; p.slope = a->slope; p.gain = a->gain;
  p.preact_add = a->preact_add;
  p.preact_c = a->preact_c > 0 ? a->preact_c : a->cout;
  p.preact_up = a->preact_up == 2 ? 2 : 1;
  p.blend_src = a->blend_src; p.blend_pool = a->blend_pool; p.blend_a = a->blend_a; p.blend_b = a->blend_b;
  p.out_act = a->out_act; p.out_planes = a->out_planes; p.out_scale = a->out_scale; p.out_f32b = a->out_f32b;
  p.out_nchw = a->out_nchw; p.rgb_w = a->rgb_w; p.rgb_out = a->rgb_out; p.out_raw_up = a->out_raw_up;
  p.out_pool = a->out_pool ? 1 : 0;
  p.x = a->x; p.wpk = a->wpk;

  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }

  if (a->flags & DGE_CONV_FLAG_CHECKER) {
    int grid = p.total_tiles < 8 * g_num_sms ? p.total_tiles : 8 * g_num_sms;
    if (up)
      conv_checker_kernel<1><<<grid, 128, 0, stream>>>(p);
    else
      conv_checker_kernel<0><<<grid, 128, 0, stream>>>(p);
    count_launch();
    return check_launch("conv_checker_kernel");
  }

  // tensor maps
  CUtensorMap tmA, tmB;
  {
    const uint64_t c8p = (uint64_t)(p.Cin / 8) * p.planes;
    // the stored map may extend beyond the output domain (in_h / in_w): the halo then reads real rows / columns there
    const uint64_t ih = a->in_h ? a->in_h : p.H, iw = a->in_w ? a->in_w : p.W;
    uint64_t dims[4] = {2 * iw, ih, c8p, (uint64_t)p.N};
    uint64_t strides[3] = {iw * 16, ih * iw * 16, c8p * ih * iw * 16};
    uint32_t box[4] = {(uint32_t)(2 * p.pw), (uint32_t)p.ph, (uint32_t)((p.kc / 8) * p.planes), 1};
    int r = make_tmap(&tmA, a->x, 4, dims, strides, box);
    if (r) return r;
  }
  {
    // WPK [taps][Cin/8][planes][Cout][8] viewed as uint64 [taps][c8p][Cout/rb][2*rb]: a box of (cw/rb) row blocks
    // lands in smem as [k-group][cw rows][16 B] -- the canonical K-major operand with SBO = 128 B.
    const uint64_t c8p = (uint64_t)(p.cin_w / 8) * p.planes;
    const int wtaps = (a->kind == DGE_CONV_1X1) ? 1 : (a->kind == DGE_CONV_DOWN4X4S2 ? 16 : 9);
    const int rb = p.nsub_local > 128 ? 128 : p.nsub_local;   // (pair mode: each CTA loads its half of the rows)
    uint64_t dims[4] = {(uint64_t)2 * rb, (uint64_t)(p.Cout / rb), c8p, (uint64_t)wtaps};
    uint64_t strides[3] = {(uint64_t)rb * 16, (uint64_t)p.Cout * 16, c8p * p.Cout * 16};
    uint32_t box[4] = {(uint32_t)(2 * rb), (uint32_t)(p.nsub_local / rb), (uint32_t)((p.kc / 8) * p.planes), 1};
    int r = make_tmap(&tmB, a->wpk, 4, dims, strides, box);
    if (r) return r;
    p.b_rb = rb;
  }

  // TMEM is 512 columns per SM: keep co-resident CTAs * tmem_cols <= 512 by padding the smem request.
  const size_t min_smem = (227 * 1024) / (max_occ + 1) + 1024;
  if (smem < min_smem) smem = min_smem;
  static bool attr_set[14] = {false, false, false, false, false, false, false, false, false, false, false, false, false,
                              false};
  // EPI 2: 1x1 residual convs with a same-resolution blend (their blend loads are prefetched one unit ahead);
  // EPI 3: no residual / blend / pooled / NCHW output -- a leaner instantiation for the most common launches
  int epi = up ? 1 : ((a->kind == DGE_CONV_1X1 && a->blend_src && !a->blend_pool && !p.stack && p.ksplit == 1) ? 2 : 0);
  if (epi == 0 && p.ksplit == 1 && !a->preact_add && !a->blend_src && !a->out_nchw && !a->out_pool) epi = 3;
  // EPI 4: the lean epilogue + the 2x2-mean output only (the encoder's conv_2 at every resolution): the generic EPI 0
  // instantiation spills at the 168-register limit and paced these launches, not their MMAs
  if (epi == 0 && p.ksplit == 1 && !a->preact_add && !a->blend_src && !a->out_nchw && a->out_pool && !a->rgb_w) epi = 4;
  // non-split: index 2*epi + pair (epi 0..4); split-K: 10 + 2*epi + pair (epi 0 / 1)
  const int ei = p.ksplit > 1 ? 10 + 2 * epi + (p.pair ? 1 : 0) : 2 * epi + (p.pair ? 1 : 0);
  const void* ktab[14] = {(const void*)conv_mma_kernel<0, false, false>, (const void*)conv_mma_kernel<0, true, false>,
                          (const void*)conv_mma_kernel<1, false, false>, (const void*)conv_mma_kernel<1, true, false>,
                          (const void*)conv_mma_kernel<2, false, false>, (const void*)conv_mma_kernel<2, true, false>,
                          (const void*)conv_mma_kernel<3, false, false>, (const void*)conv_mma_kernel<3, true, false>,
                          (const void*)conv_mma_kernel<4, false, false>, (const void*)conv_mma_kernel<4, true, false>,
                          (const void*)conv_mma_kernel<0, false, true>,  (const void*)conv_mma_kernel<0, true, true>,
                          (const void*)conv_mma_kernel<1, false, true>,  (const void*)conv_mma_kernel<1, true, true>};
  const void* kfn = ktab[ei];
  if (!attr_set[ei]) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
      return DGE_ERR_CUDA;
    }
    attr_set[ei] = true;
  }
  int grid = g_num_sms * max_occ;
  if (p.pair) {
    grid &= ~1;
    if (grid > 2 * p.sched_units) grid = 2 * p.sched_units;
  } else if (grid > p.sched_units) {
    grid = p.sched_units;
  }
  if (p.ksplit > 1) {   // the partial sums are accumulated into a zeroed map
    void* z = up ? (void*)a->out_raw_up : (void*)a->splitk_ws;
    const size_t zb = up ? (size_t)a->n * a->cout * (2 * a->h + 1) * (2 * a->w + 1) * 4
                         : (size_t)a->n * a->cout * a->h * a->w * 4;
    cudaError_t me = cudaMemsetAsync(z, 0, zb, stream);
    if (me != cudaSuccess) {
      set_error("conv: split-K memset failed: %s", cudaGetErrorString(me));
      return DGE_ERR_CUDA;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  void* kargs[3] = {(void*)&tmA, (void*)&tmB, (void*)&p};
  cudaError_t le = cudaLaunchKernelExC(&cfg, kfn, kargs);
  if (le != cudaSuccess) {
    set_error("conv_mma_kernel launch failed: %s", cudaGetErrorString(le));
    return DGE_ERR_CUDA;
  }
  count_launch();
  int lr = check_launch("conv_mma_kernel");
  if (lr || p.ksplit == 1 || up) return lr;
  const size_t work = (size_t)p.N * (p.Cout / 16) * (p.out_pool ? (p.H / 2) * (p.W / 2) : p.H * p.W);
  int fgrid = (int)((work + 255) / 256);
  if (fgrid > 8 * g_num_sms) fgrid = 8 * g_num_sms;
  conv_splitk_finish_kernel<<<fgrid, 256, 0, stream>>>(p);
  count_launch();
  return check_launch("conv_splitk_finish_kernel");
}

size_t conv_splitk_ws_bytes(const dge_conv_args* a) {
  if (!a || a->kind == DGE_CONV_UP3X3 || a->cout % 16 || a->cin % 16) return 0;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return splitk_wanted(a, conv_cw_full(a), g_num_sms) ? (size_t)a->n * a->cout * a->h * a->w * 4 : 0;
}

}  // namespace dge

#ifdef DGE_ROLE_TIMING
// slots: 0 producer wait a_empty, 1 producer work, 2 MMA wait tm_empty, 3 MMA wait a_full, 4 MMA issue,
//        5 epilogue (warp 2) wait tm_full, 6 epilogue work   -- cycles summed over CTAs
extern "C" int dge_exp_role_cycles(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  if (out16) cudaMemcpyFromSymbol(out16, dge::g_role_cycles, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(dge::g_role_cycles, z, sizeof(z));
  }
  return 0;
}
#endif

extern "C" size_t dge_conv_splitk_ws_bytes(const dge_conv_args* a) { return dge::conv_splitk_ws_bytes(a); }

extern "C" int dge_conv_forward(const dge_conv_args* a, void* stream) {
  return dge::conv_forward(a, static_cast<cudaStream_t>(stream));
}
