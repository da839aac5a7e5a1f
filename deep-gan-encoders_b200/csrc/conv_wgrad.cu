// conv_wgrad.cu -- weight gradient of the 3x3 / 1x1 stride-1 convolutions on tcgen05 tensor cores (sm_100a).
//
// First backward building block of SURVEY 8(f)-1 (encoder training: E_align_s2.py:205-233 calls loss.backward() and
// LREQAdam.step on every ln.Conv2d weight of model/E/E.py:27-36).  Computes, for y = conv2d(x, w, padding=k/2):
//     dW[co][ci][ky][kx] = sum_{n, y, x} dy[n, co, y, x] * x[n, ci, y + ky - pad, x + kx - pad]
// which is a GEMM whose contraction index is the PIXEL.  Both operands already live in the ACT layout
// ([N][C/8][planes][H][W][8] bf16): per pixel, 8 channels are 16 contiguous bytes.  A pixel tile staged by TMA is
// therefore a canonical *MN-major*, no-swizzle UMMA operand (K = pixels at a 16-byte stride, 8 K rows x 16 bytes form
// a 128-byte core matrix; LBO = 128 B between 8-pixel groups, SBO = bytes between 8-channel groups) -- no transpose
// pass, and a filter tap is again nothing but a start-address offset (kx pixels) into the one resident x patch.
// (tools/probe_mnmajor.cu checks this descriptor reading and the 16-byte-aligned tap shifts on the device.)
//
//   grid.x : pixel-tile chunks (split-K over the contraction; fp32 atomics merge the partial sums into dW)
//   grid.y : (filter row ky) x (128-row block of Cout) x (<=128-column block of Cin)
//   CTA    : warp 0 = TMA producer (multi-stage mbarrier ring), warp 1 = MMA issuer + TMEM owner,
//            all 4 warps = final epilogue (one TMEM lane = one output channel per thread).
//   TMEM   : taps_x accumulators of [128 x nblk] fp32 side by side (<= 384 columns), live for the whole CTA.
//   planes == 2 => split precision as in the forward kernel: dy = dh+dl, x = xh+xl, acc += dh*xh + dh*xl + dl*xh.
//
// Thin layers (Cin <= 32: the 1024^2 / 512^2 blocks of the encoder), "cross-stacked" mode: an MMA costs its 4 KB A fetch
// (~64 clk) however few of its 128 rows and N columns carry data, so both MMA dimensions are filled with filter taps:
//     dW[o][i][ky][kx] = sum_q dy[o][q - (0, kx-1)] * x[i][q + (ky-1, 0)]
//   A rows    = (kx, o):  three copies of the dy tile shifted by one column each (TMA zero-fills outside the map),
//   B columns = (ky, i-group, hi | lo): three copies of the x tile shifted by one row each, the lo plane next to the hi plane,
// and a K step of 16 pixels is TWO MMAs (dh x [xh | xl] -> D1, dl x xh -> D2) instead of six: ~160 clk per 16 pixels per
// SM, which at 148 SMs consumes the operands at HBM speed.  The epilogue adds the three partial products.
#include <stdlib.h>
#include <string.h>

#include "dge_common.cuh"
#include "tma_ptx.cuh"

namespace dge {
namespace {

constexpr int WG_THREADS = 128;
constexpr int WG_MAX_STAGES = 4;

struct WgradParams {
  int N, H, W, Cout, Cin, ksize, planes;
  int taps_x, pad;
  int th, tw, pw;                  // pixel tile (th x tw) and x patch width (tw + 2*pad)
  int tiles_x, tiles_y, tiles_total, chunks;
  int nblk, cin_blocks, cout_blocks;   // MMA N (Cin block padded to 16), block counts
  int a_groups, b_groups;          // 8-channel groups per TMA box (x planes)
  uint32_t dy_tile_bytes, x_tile_bytes, dy_bytes, x_bytes, stage_bytes;
  int stages, tmem_cols;
  int stacked;                     // all 9 taps side by side in N (Cin <= 32): x is staged as 9 shifted tiles
  int n1, n2;                      // stacked: N of the first / second MMA of a K step (n2 = 0: one MMA)
  uint32_t slot_bytes;             // stacked: bytes of one tap's x tile set
  int xstack;                      // cross-stacked mode (see the header): rows = (kx, o), columns = (ky, i, plane)
  int R;                           // xstack: output channels per CTA (3 * R <= 128 rows)
  uint32_t a_slot_bytes;           // xstack: bytes of one shifted dy tile set
  int rowpair;                     // maps <= 8 pixels wide: a K step is 8 pixels of row r + 8 pixels of row r+1
  int atomic, accumulate;          // chunks > 1: partial sums merge with atomics; else plain store / add
  float* dw;
};

__device__ __forceinline__ void wg_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wg_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// MN-major, no-swizzle operand: element (mn, k) at start + (mn/8)*SBO + (k/8)*LBO + (k%8)*16 + (mn%8)*2
__device__ __forceinline__ uint64_t wg_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::f16: D = f32, A = B = bf16, both MN-major (bits 15, 16), M = 128, N = n
__device__ __forceinline__ uint32_t wg_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void wg_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void wg_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                  const WgradParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[WG_MAX_STAGES], empty_bar[WG_MAX_STAGES], done_bar;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // block coordinates
  const int cib = blockIdx.y % p.cin_blocks;
  const int cob = (blockIdx.y / p.cin_blocks) % p.cout_blocks;
  const int ky = blockIdx.y / (p.cin_blocks * p.cout_blocks);   // 0 in stacked mode (grid.y = cout blocks)
  const int t0 = (int)(((long long)blockIdx.x * p.tiles_total) / p.chunks);
  const int t1 = (int)(((long long)(blockIdx.x + 1) * p.tiles_total) / p.chunks);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  wg_fence_before();
  __syncthreads();
  wg_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // =================================== TMA producer ====================================
    if (lane == 0) {
      const int tiles_per_img = p.tiles_x * p.tiles_y;
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int n = t / tiles_per_img, rem = t - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int y0 = ty * p.th, x0 = tx * p.tw;
        uint8_t* st = smem + (size_t)s * p.stage_bytes;
        uint8_t* sx = st + (p.xstack ? (size_t)3 * p.a_slot_bytes : (size_t)16 * p.planes * p.dy_tile_bytes);
        if (p.xstack) {
          mbar_arrive_expect_tx(&full_bar[s], 3u * p.dy_bytes + 3u * p.x_bytes);
#pragma unroll 1
          for (int k = 0; k < 3; ++k) {
            tma_load_4d(st + (size_t)k * p.a_slot_bytes, &tm_dy, &full_bar[s], 2 * (x0 + 1 - k), y0, cob * p.a_groups, n);
            tma_load_4d(sx + (size_t)k * p.slot_bytes, &tm_x, &full_bar[s], 2 * x0, y0 + k - 1, 0, n);
          }
        } else if (p.stacked) {
          mbar_arrive_expect_tx(&full_bar[s], p.dy_bytes + 9u * p.x_bytes);
          tma_load_4d(st, &tm_dy, &full_bar[s], 2 * x0, y0, cob * 16 * p.planes, n);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap)
            tma_load_4d(sx + (size_t)tap * p.slot_bytes, &tm_x, &full_bar[s], 2 * (x0 + tap % 3 - 1), y0 + tap / 3 - 1,
                        0, n);
        } else {
          mbar_arrive_expect_tx(&full_bar[s], p.dy_bytes + p.x_bytes);
          tma_load_4d(st, &tm_dy, &full_bar[s], 2 * x0, y0, cob * 16 * p.planes, n);
          tma_load_4d(sx, &tm_x, &full_bar[s], 2 * (x0 - p.pad), y0 + ky - p.pad, cib * (p.nblk / 8) * p.planes, n);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =================================== MMA issuer ======================================
    if (lane == 0) {
      const uint32_t idesc = wg_idesc(p.nblk), idesc1 = wg_idesc(p.n1), idesc2 = wg_idesc(p.n2 ? p.n2 : 16);
      const uint32_t a_sbo = p.planes * p.dy_tile_bytes, b_sbo = p.planes * p.x_tile_bytes;
      // the two 8-pixel K groups of a K step: neighbours in a row (128 B apart), or the same 8 columns of two rows
      const uint32_t a_lbo = p.rowpair ? (uint32_t)p.tw * 16u : 128u, b_lbo = p.rowpair ? (uint32_t)p.pw * 16u : 128u;
      const uint64_t a_desc0 = wg_desc(0, a_lbo, a_sbo), b_desc0 = wg_desc(0, b_lbo, b_sbo);
      const uint64_t bx_desc0 = wg_desc(0, b_lbo, p.x_tile_bytes);   // xstack: (hi, lo) planes as neighbouring N groups
      const uint32_t a_lo = p.dy_tile_bytes >> 4, b_lo = p.x_tile_bytes >> 4;   // plane 1 (lo) offsets, 16-byte units
      const int segs = p.rowpair ? 1 : p.tw / 16, rstep = p.rowpair ? 2 : 1;
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        wg_fence_after();
        const uint32_t a_base = smem_u32(smem + (size_t)s * p.stage_bytes);
        const uint32_t b_base = a_base + (p.xstack ? 3u * p.a_slot_bytes : 16u * p.planes * p.dy_tile_bytes);
        for (int r = 0; r < p.th; r += rstep) {
          for (int sg = 0; sg < segs; ++sg) {
            const uint64_t da = a_desc0 + (uint64_t)(((a_base >> 4) + (uint32_t)(r * p.tw + sg * 16)) & 0x3fff);
            const uint32_t acc0 = (it == 0 && r == 0 && sg == 0) ? 0u : 1u;
            if (p.xstack) {
              // D1 += dh(kx-stacked) x [xh | xl](ky-stacked): consecutive B groups alternate hi / lo (stride = one tile);
              // D2 += dl x xh: every other group (stride = planes tiles)
              const uint32_t boff = ((b_base >> 4) + (uint32_t)(r * p.tw + sg * 16)) & 0x3fff;
              wg_mma(tmem, da, bx_desc0 + (uint64_t)boff, idesc1, acc0);
              if (p.planes == 2) wg_mma(tmem + (uint32_t)p.n1, da + a_lo, b_desc0 + (uint64_t)boff, idesc2, acc0);
            } else if (p.stacked) {
              // one (or two) wide MMAs cover all 9 taps: N group g = tap * (Cin/8) + channel group, uniform stride
              const uint64_t db = b_desc0 + (uint64_t)(((b_base >> 4) + (uint32_t)(r * p.tw + sg * 16)) & 0x3fff);
              wg_mma(tmem, da, db, idesc1, acc0);
              if (p.planes == 2) {
                wg_mma(tmem, da, db + b_lo, idesc1, 1u);
                wg_mma(tmem, da + a_lo, db, idesc1, 1u);
              }
              if (p.n2) {
                const uint64_t db2 = db + (uint64_t)((5u * p.slot_bytes) >> 4);
                const uint32_t d2 = tmem + (uint32_t)p.n1;
                wg_mma(d2, da, db2, idesc2, acc0);
                if (p.planes == 2) {
                  wg_mma(d2, da, db2 + b_lo, idesc2, 1u);
                  wg_mma(d2, da + a_lo, db2, idesc2, 1u);
                }
              }
            } else {
              for (int kx = 0; kx < p.taps_x; ++kx) {
                const uint64_t db =
                    b_desc0 + (uint64_t)(((b_base >> 4) + (uint32_t)(r * p.pw + sg * 16 + kx)) & 0x3fff);
                const uint32_t d = tmem + (uint32_t)(kx * p.nblk);
                wg_mma(d, da, db, idesc, acc0);
                if (p.planes == 2) {
                  wg_mma(d, da, db + b_lo, idesc, 1u);
                  wg_mma(d, da + a_lo, db, idesc, 1u);
                }
              }
            }
          }
        }
        wg_commit(&empty_bar[s]);   // frees the stage once these MMAs have read it
      }
      wg_commit(&done_bar);
    }
    __syncwarp();
  }

  // =================================== epilogue (all warps) ================================
  mbar_wait_relaxed(&done_bar, 0);
  wg_fence_after();
  if (p.xstack) {
    // lane = kx * R + o; D1 columns ((ky * G + g) * planes + plane) * 8 + c, D2 columns (ky * G + g) * 8 + c  (G = Cin / 8)
    if (warp * 32 < 3 * p.R && t1 > t0) {
      const int row = warp * 32 + lane;
      const int kx = row / p.R, co = cob * p.R + (row - kx * p.R);
      const int G = p.Cin / 8;
      for (int kg = 0; kg < 3 * G; ++kg) {
        uint32_t r1[16], r2[8];
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        if (p.planes == 2) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
              "%15}, [%16];"
              : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]),
                "=r"(r1[8]), "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]),
                "=r"(r1[15])
              : "r"(lane_base + (uint32_t)(kg * 16)));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]),
                         "=r"(r2[7])
                       : "r"(lane_base + (uint32_t)(p.n1 + kg * 8)));
        } else {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]),
                         "=r"(r1[7])
                       : "r"(lane_base + (uint32_t)(kg * 8)));
#pragma unroll
          for (int i = 0; i < 8; ++i) r1[8 + i] = r2[i] = 0u;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int ky = kg / G, ci0 = (kg - ky * G) * 8;
        if (row < 3 * p.R) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float* q = p.dw + ((size_t)co * p.Cin + ci0 + i) * 9 + ky * 3 + kx;
            const float v = __uint_as_float(r1[i]) + __uint_as_float(r1[8 + i]) + __uint_as_float(r2[i]);
            if (p.atomic) atomicAdd(q, v);
            else *q = p.accumulate ? *q + v : v;
          }
        }
      }
    }
  }
  const int co = cob * 128 + warp * 32 + lane;
  if (!p.xstack && cob * 128 + warp * 32 < p.Cout && t1 > t0) {
    const int ncols = (p.stacked ? 9 : p.taps_x) * p.nblk;
    const int tap_base = p.stacked ? 0 : ky * p.ksize;
    const int kk = p.ksize * p.ksize;
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
          "%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int tl = c0 / p.nblk;            // nblk is a multiple of 16: a 16-column group never straddles taps
      const int ci0 = cib * p.nblk + (c0 - tl * p.nblk);
      const int tap = tap_base + tl;
      if (co < p.Cout) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int ci = ci0 + i;
          if (ci < p.Cin) {
            float* q = p.dw + ((size_t)co * p.Cin + ci) * kk + tap;
            const float v = __uint_as_float(r[i]);
            if (p.atomic) atomicAdd(q, v);
            else *q = p.accumulate ? *q + v : v;     // this CTA is the only writer of its dW block
          }
        }
      }
    }
  }
  wg_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

int g_wg_sms = 0;

}  // namespace
}  // namespace dge

using namespace dge;

extern "C" int dge_conv_wgrad(const void* dy_act, const void* x_act, float* dw, int n, int cout, int cin, int h,
                              int w, int ksize, int planes, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DGE_REQUIRE(dy_act && x_act && dw, "conv_wgrad: null pointer");
  DGE_REQUIRE(ksize == 1 || ksize == 3, "conv_wgrad: ksize must be 1 or 3 (got %d)", ksize);
  DGE_REQUIRE(planes == 1 || planes == 2, "conv_wgrad: planes must be 1 or 2 (got %d)", planes);
  DGE_REQUIRE(n > 0 && h > 0 && w > 0, "conv_wgrad: empty input (n=%d h=%d w=%d)", n, h, w);
  DGE_REQUIRE(cout > 0 && cin > 0 && cout % 8 == 0 && cin % 8 == 0,
              "conv_wgrad: channel counts must be positive multiples of 8 (cout=%d cin=%d)", cout, cin);
  if (!g_wg_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_wg_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_wg_sms <= 0) g_wg_sms = 148;
  }

  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = n; p.H = h; p.W = w; p.Cout = cout; p.Cin = cin; p.ksize = ksize; p.planes = planes;
  p.taps_x = ksize;
  p.pad = ksize / 2;
  if (w <= 8) {
    p.rowpair = 1;
    p.tw = 8;
    p.th = (h + 1) & ~1;
    if (p.th > 8) p.th = 8;
  } else {
    int segs = (w + 15) / 16;
    if (segs > 4) segs = 4;
    p.tw = 16 * segs;
    p.th = 64 / p.tw;
    if (p.th > h) p.th = h;
  }
  // tap-stacked mode for the thin layers of the encoder (E.py: 16 and 32 input channels at 1024^2 / 512^2): with
  // N = Cin per tap an MMA is bound by the 4 KB A fetch (~64 clk) whatever N is, so nine N=16 MMAs cost 9x one N=144
  // MMA.  x is staged as 9 shifted tiles (9x a small tile) so that (tap, channel group) is ONE uniform N stride.
  static int no_stack = -1;
  if (no_stack < 0) no_stack = getenv("DGE_WGRAD_NO_STACK") ? 1 : 0;
  p.stacked = (ksize == 3 && cin % 16 == 0 && cin <= 32 && !no_stack) ? 1 : 0;
  // cross-stacked mode (header): R output channels per CTA, the largest of 40 / 32 / 24 / 16 / 8 that divides Cout
  static int no_xstack = -1;
  if (no_xstack < 0) no_xstack = getenv("DGE_WGRAD_NO_XSTACK") ? 1 : 0;
  if (p.stacked && !no_xstack) {
    // ... as long as that is at most two CTAs' worth: every Cout block repeats the MMAs over the same x tiles, and from three
    // blocks on round 1's layout (128 output channels per CTA, taps on N only) issues fewer of them
    for (int r = 40; r >= 8 && !p.xstack; r -= 8)
      if (cout % r == 0 && cout / r <= 2) {
        p.xstack = 1;
        p.R = r;
      }
    if (p.xstack) p.stacked = 0;
  }
  p.pw = (p.stacked || p.xstack) ? p.tw : p.tw + 2 * p.pad;
  p.tiles_x = (w + p.tw - 1) / p.tw;
  p.tiles_y = (h + p.th - 1) / p.th;
  const long long tiles = (long long)n * p.tiles_x * p.tiles_y;
  DGE_REQUIRE(tiles < (1ll << 30), "conv_wgrad: too many pixel tiles (%lld)", tiles);
  p.tiles_total = (int)tiles;
  const int cin16 = (cin + 15) / 16 * 16;
  p.nblk = cin16 < 128 ? cin16 : 128;
  // Cost model (measured, tools/probe_wgrad.py): a CTA's MMA loop (A fetch or math per MMA, whichever is longer), its
  // epilogue -- 128 rows x taps_x*nblk columns leave as scattered 4-byte stores / atomics, ~45 us per 384 columns + ~7 us
  // fixed, which is what small maps pay for -- and, when the contraction is split over c chunks, one more pass of fp32
  // atomics over dW per chunk (~200 G atomics/s) plus the memset.
  const int ksteps_tile = p.rowpair ? p.th / 2 : p.th * (p.tw / 16);
  const double pass_us = (double)cout * cin * ksize * ksize / 200e3;
  auto predict = [&](int mma_n, double mmas_kstep, int blocks_y, int ncols, int* chunks_out) {
    const double clk_mma = mma_n / 2 > 64 ? mma_n / 2 : 64;
    const double cta_us = (double)p.tiles_total * ksteps_tile * mmas_kstep * clk_mma / 1900.0;
    const double epi_us = 7.0 + 45.0 * ncols / 384.0;
    int max_chunks = (2 * g_wg_sms + blocks_y - 1) / blocks_y;
    if (max_chunks > p.tiles_total) max_chunks = p.tiles_total;
    if (max_chunks < 1) max_chunks = 1;
    int chunks = 1;
    double best = (cta_us + epi_us) * ((blocks_y + g_wg_sms - 1) / g_wg_sms);
    for (int c = 2; c <= max_chunks; ++c) {
      const int waves = (blocks_y * c + g_wg_sms - 1) / g_wg_sms;
      const double t = (cta_us / c + epi_us) * waves + pass_us * c + 3.0;
      if (t < best) {
        best = t;
        chunks = c;
      }
    }
    *chunks_out = chunks;
    return best;
  };
  const double split_mmas = planes == 2 ? 3.0 : 1.0;
  if (!p.stacked && p.nblk == 128) {
    // narrower Cin blocks double the CTAs that share the epilogue: pays when the map is small (the MMA loop is short)
    static int nblk_env = -1;
    if (nblk_env < 0) nblk_env = getenv("DGE_WGRAD_NBLK") ? atoi(getenv("DGE_WGRAD_NBLK")) : 0;      // A/B switch
    const int cob_n = (cout + 127) / 128;
    int c128, c64;
    const double t128 = predict(128, split_mmas * ksize, ksize * cob_n * ((cin + 127) / 128), ksize * 128, &c128);
    const double t64 = predict(64, split_mmas * ksize, ksize * cob_n * ((cin + 63) / 64), ksize * 64, &c64);
    if (nblk_env == 64 || nblk_env == 128) p.nblk = nblk_env;
    else if (t64 < t128) p.nblk = 64;
  }
  p.cin_blocks = (cin + p.nblk - 1) / p.nblk;
  if (p.stacked) {
    p.n1 = 9 * cin <= 256 ? 9 * cin : 5 * cin;
    p.n2 = 9 * cin - p.n1;
  }
  if (p.xstack) {
    p.n1 = 3 * cin * planes;     // D1: (ky, channel group, plane)
    p.n2 = 3 * cin;              // D2: (ky, channel group), planes == 2 only
  }
  p.cout_blocks = p.xstack ? cout / p.R : (cout + 127) / 128;
  const int a_groups_total = (cout / 8) * planes, b_groups_total = (cin / 8) * planes;
  p.a_groups = 16 * planes < a_groups_total ? 16 * planes : a_groups_total;
  if (p.xstack) p.a_groups = (p.R / 8) * planes;
  p.b_groups = (p.nblk / 8) * planes < b_groups_total ? (p.nblk / 8) * planes : b_groups_total;
  p.dy_tile_bytes = (uint32_t)(p.th * p.tw * 16);
  p.x_tile_bytes = (uint32_t)(p.th * p.pw * 16);
  p.dy_bytes = p.dy_tile_bytes * p.a_groups;
  p.x_bytes = p.x_tile_bytes * p.b_groups;
  // smem image per stage keeps room for the full 16 / nblk/8 groups so the descriptor strides do not depend on clamping
  p.slot_bytes = p.x_tile_bytes * (uint32_t)((p.nblk / 8) * planes);
  const uint32_t x_region = p.slot_bytes * (p.stacked ? 9u : (p.xstack ? 3u : 1u));
  p.a_slot_bytes = p.dy_bytes;
  // xstack: the M = 128 descriptor spans 16 row groups whatever 3 * R is (rows past 3 * R land in accumulator lanes nobody
  // reads); the span of the last stage stays inside the allocation through `tail` bytes of slack
  const uint32_t a_region = p.xstack ? 3u * p.a_slot_bytes : 16u * planes * p.dy_tile_bytes;
  const uint32_t tail = p.xstack ? 16u * planes * p.dy_tile_bytes : 0u;
  p.stage_bytes = (a_region + x_region + 127u) & ~127u;
  p.stages = (int)(((p.xstack ? 224u : 216u) * 1024u - tail) / p.stage_bytes);
  if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
  DGE_REQUIRE(p.stages >= 2, "conv_wgrad: stage of %u bytes does not fit twice in shared memory", p.stage_bytes);
  int cols = 32;
  while (cols < (p.xstack ? p.n1 + (planes == 2 ? p.n2 : 0) : (p.stacked ? 9 : p.taps_x) * p.nblk)) cols *= 2;
  p.tmem_cols = cols;
  const int blocks_y = (p.stacked || p.xstack) ? p.cout_blocks : ksize * p.cout_blocks * p.cin_blocks;
  // split of the contraction: c chunks shorten every CTA's MMA loop by c but each adds one pass of fp32 atomics over dW
  // (~200 G atomics/s measured); with one chunk the CTA owns its dW block and stores it (no memset, no atomics).
  int chunks = 1;
  if (p.xstack) {
    // one K step = MMA(N = n1) + MMA(N = n2): pass the summed clocks as one MMA of twice that many columns
    const int clk = (p.n1 / 2 > 64 ? p.n1 / 2 : 64) + (planes == 2 ? (p.n2 / 2 > 64 ? p.n2 / 2 : 64) : 0);
    predict(2 * clk, 1.0, blocks_y, 3 * cin, &chunks);
  } else {
    predict(p.stacked ? p.n1 : p.nblk, split_mmas * (p.stacked ? (p.n2 ? 2 : 1) : p.taps_x), blocks_y,
            (p.stacked ? 9 : p.taps_x) * p.nblk, &chunks);
  }
  p.chunks = chunks;
  p.atomic = chunks > 1;
  p.accumulate = accumulate;
  p.dw = dw;

  CUtensorMap tm_dy, tm_x;
  {
    const uint64_t c8p = (uint64_t)a_groups_total;
    uint64_t dims[4] = {(uint64_t)2 * w, (uint64_t)h, c8p, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)w * 16, (uint64_t)h * w * 16, c8p * h * w * 16};
    uint32_t box[4] = {(uint32_t)(2 * p.tw), (uint32_t)p.th, (uint32_t)p.a_groups, 1};
    int r = make_tmap(&tm_dy, dy_act, 4, dims, strides, box);
    if (r) return r;
  }
  {
    const uint64_t c8p = (uint64_t)b_groups_total;
    uint64_t dims[4] = {(uint64_t)2 * w, (uint64_t)h, c8p, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)w * 16, (uint64_t)h * w * 16, c8p * h * w * 16};
    uint32_t box[4] = {(uint32_t)(2 * p.pw), (uint32_t)p.th, (uint32_t)p.b_groups, 1};
    int r = make_tmap(&tm_x, x_act, 4, dims, strides, box);
    if (r) return r;
  }

  const size_t smem = (size_t)p.stages * p.stage_bytes + tail + 256;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv_wgrad: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return DGE_ERR_CUDA;
    }
    attr_smem = smem;
  }
  if (!accumulate && p.atomic) {
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)cout * cin * ksize * ksize * sizeof(float), stream);
    if (e != cudaSuccess) {
      set_error("conv_wgrad: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
      return DGE_ERR_CUDA;
    }
  }
  dim3 grid((unsigned)p.chunks, (unsigned)blocks_y, 1);
  conv_wgrad_kernel<<<grid, WG_THREADS, smem, stream>>>(tm_dy, tm_x, p);
  count_launch();
  return check_launch("conv_wgrad_kernel");
}
