// Probe: tcgen05.mma kind::f16 with MN-major (a_major = b_major = 1), no-swizzle shared-memory operands.
//
// Why: the weight gradient of a conv contracts over PIXELS, and the activation layout of this repo
// ([N][C/8][planes][H][W][8] bf16) stores 8 channels contiguously per pixel -- i.e. a pixel tile in shared memory is
// an MN-major operand for that GEMM (K = pixels at a 16-byte stride, MN = channels contiguous).  This standalone
// program checks, on the device, which descriptor field carries which stride for that layout and that a tap shift
// (start address moved by whole pixels, not 128-byte aligned) is legal -- the two facts a wgrad kernel needs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_exp/probe_mnmajor tools/probe_mnmajor.cu && \
//   timeout 60 tools/_exp/probe_mnmajor
//
// Layout under test, element (mn, k) of a [MN x 16] bf16 operand:
//   byte = (mn / 8) * MNG + (k / 8) * KG + (k % 8) * 16 + (mn % 8) * 2         (KG = 128: pixels contiguous)
// Hypothesis H1: descriptor SBO = MNG, LBO = KG.   Hypothesis H2: SBO = KG, LBO = MNG.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

struct Args {
  uint32_t a_bytes, b_bytes;        // operand images to copy into shared memory
  uint32_t a_start, b_start;        // start-address offsets inside the images (tap shift)
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t idesc;
  int n;
};

__global__ void __launch_bounds__(128, 1) probe(const uint8_t* a_img, const uint8_t* b_img, float* out, Args g) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((g.a_bytes + 127) & ~127u);
  for (uint32_t i = threadIdx.x; i < g.a_bytes; i += blockDim.x) sa[i] = a_img[i];
  for (uint32_t i = threadIdx.x; i < g.b_bytes; i += blockDim.x) sb[i] = b_img[i];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    const uint64_t da = make_desc(smem_u32(sa) + g.a_start, g.a_lbo, g.a_sbo);
    const uint64_t db = make_desc(smem_u32(sb) + g.b_start, g.b_lbo, g.b_sbo);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm),
        "l"(da), "l"(db), "r"(g.idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
  }
  // wait for the MMA
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(&bar)), "r"(0u)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < g.n; c0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[(size_t)threadIdx.x * g.n + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm));
}

static uint32_t idesc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// operand image: [MN x KTOT] integers, MN-major layout with the given strides; KTOT may exceed 16 (tap shift source)
static std::vector<uint8_t> build_mn_major(const std::vector<int>& v, int mn, int ktot, uint32_t mng, uint32_t kg) {
  size_t bytes = (size_t)(mn / 8) * mng + (size_t)((ktot + 7) / 8) * kg + 256;
  std::vector<uint8_t> img(bytes, 0);
  for (int r = 0; r < mn; ++r)
    for (int k = 0; k < ktot; ++k) {
      __nv_bfloat16 h = __float2bfloat16((float)v[(size_t)r * ktot + k]);
      size_t off = (size_t)(r / 8) * mng + (size_t)(k / 8) * kg + (size_t)(k % 8) * 16 + (size_t)(r % 8) * 2;
      memcpy(&img[off], &h, 2);
    }
  return img;
}

int main() {
  const int M = 128, N = 32, KTOT = 40;  // 40 "pixels" available, each MMA consumes 16 starting at a shift
  std::vector<int> a((size_t)M * KTOT), b((size_t)N * KTOT);
  srand(7);
  for (auto& x : a) x = rand() % 7 - 3;
  for (auto& x : b) x = rand() % 7 - 3;
  // pixels contiguous: KG = 128; channel groups: a plane of KTOT pixels (640 B) padded to 768 B
  const uint32_t KG = 128, MNG = 768;
  std::vector<uint8_t> a_img = build_mn_major(a, M, KTOT, MNG, KG), b_img = build_mn_major(b, N, KTOT, MNG, KG);
  uint8_t *da, *db;
  float* dout;
  CK(cudaMalloc(&da, a_img.size()));
  CK(cudaMalloc(&db, b_img.size()));
  CK(cudaMalloc(&dout, sizeof(float) * M * N));
  CK(cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice));
  const size_t smem = ((a_img.size() + 127) & ~(size_t)127) + b_img.size() + 128;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int fails = 0;
  struct Case {
    const char* name;
    int hyp, sa, sb;  // hypothesis, pixel shifts of A and B
  } cases[] = {{"H1 (SBO=MN-group stride, LBO=K-group stride), no shift", 1, 0, 0},
               {"H2 (SBO=K-group stride, LBO=MN-group stride), no shift", 2, 0, 0},
               {"H1, B shifted by 1 pixel", 1, 0, 1},
               {"H1, A shifted by 3, B by 5 pixels", 1, 3, 5},
               {"H1, A shifted by 8, B by 17 pixels", 1, 8, 17},
               {"H2, B shifted by 1 pixel", 2, 0, 1}};
  for (const Case& c : cases) {
    Args g;
    g.a_bytes = (uint32_t)a_img.size();
    g.b_bytes = (uint32_t)b_img.size();
    g.a_start = c.sa * 16;
    g.b_start = c.sb * 16;
    g.a_lbo = g.b_lbo = c.hyp == 1 ? KG : MNG;
    g.a_sbo = g.b_sbo = c.hyp == 1 ? MNG : KG;
    g.idesc = idesc_bf16(M, N, 1, 1);
    g.n = N;
    CK(cudaMemset(dout, 0, sizeof(float) * M * N));
    probe<<<1, 128, smem>>>(da, db, dout, g);
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)M * N);
    CK(cudaMemcpy(out.data(), dout, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        int ref = 0;
        for (int k = 0; k < 16; ++k) ref += a[(size_t)m * KTOT + c.sa + k] * b[(size_t)n * KTOT + c.sb + k];
        if (out[(size_t)m * N + n] != (float)ref) ++bad;
      }
    printf("%-62s : %s (%d / %d mismatches)\n", c.name, bad ? "MISMATCH" : "EXACT", bad, M * N);
    if (c.hyp == 1 && bad) ++fails;
  }
  printf("%s\n", fails ? "RESULT: H1 does not hold" : "RESULT: H1 holds (MN-major no-swizzle: SBO = channel-group stride, "
                                                        "LBO = 8-pixel-group stride; 16-byte-aligned tap shifts legal)");
  return fails ? 1 : 0;
}
