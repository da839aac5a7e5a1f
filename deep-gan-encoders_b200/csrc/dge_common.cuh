// Shared device/host helpers for the dge_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dge_b200.h"

namespace dge {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch();
int check_launch(const char* what);  // cudaGetLastError -> DGE_ERR_CUDA

#define DGE_REQUIRE(cond, ...)                \
  do {                                        \
    if (!(cond)) {                            \
      ::dge::set_error(__VA_ARGS__);          \
      return DGE_ERR_BAD_ARG;                 \
    }                                         \
  } while (0)

// ---------------------------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// split x into bf16 hi + bf16 lo with x ~= hi + lo (relative error ~2^-17)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// 8 floats -> one 16-byte chunk of bf16 hi and (optionally) lo
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16(v[i], h[i], l[i]);
  hi = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]),
                  pack_bf16x2(h[6], h[7]));
  lo = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]),
                  pack_bf16x2(l[6], l[7]));
}

__device__ __forceinline__ void unpack8(const uint4& q, float* v) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// ACT layout addressing: [N][C/8][planes][H][W][8] bf16, returns index in 16-byte units
__device__ __forceinline__ size_t act_idx16(int n, int c8, int plane, int y, int x, int C8, int planes, int H, int W) {
  return ((((size_t)n * C8 + c8) * planes + plane) * H + y) * (size_t)W + x;
}
// F32B layout: [N][C/8][H][W][8] fp32, returns index in 32-byte units
__device__ __forceinline__ size_t f32b_idx32(int n, int c8, int y, int x, int C8, int H, int W) {
  return (((size_t)n * C8 + c8) * H + y) * (size_t)W + x;
}

// One F32B chunk (8 floats = one 32-byte sector) moves as a single 256-bit access (LDG/STG.E.256, sm_100): half the
// LSU instructions of two float4 accesses, and a store never leaves a sector half written.
__device__ __forceinline__ void load8_f32b(const float* base, size_t idx32, float* v) {
  const float* p = base + idx32 * 8;
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void store8_f32b(float* base, size_t idx32, const float* v) {
  float* p = base + idx32 * 8;
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void store8_act(void* base, int n, int c8, int y, int x, int C8, int planes, int H, int W,
                                           const float* v) {
  uint4 hi, lo;
  split8(v, hi, lo);
  uint4* p = reinterpret_cast<uint4*>(base);
  p[act_idx16(n, c8, 0, y, x, C8, planes, H, W)] = hi;
  if (planes == 2) p[act_idx16(n, c8, 1, y, x, C8, planes, H, W)] = lo;
}
__device__ __forceinline__ void load8_act(const void* base, int n, int c8, int y, int x, int C8, int planes, int H,
                                          int W, float* v) {
  const uint4* p = reinterpret_cast<const uint4*>(base);
  unpack8(__ldg(p + act_idx16(n, c8, 0, y, x, C8, planes, H, W)), v);
  if (planes == 2) {
    float l[8];
    unpack8(__ldg(p + act_idx16(n, c8, 1, y, x, C8, planes, H, W)), l);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += l[i];
  }
}

}  // namespace dge
