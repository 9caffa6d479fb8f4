// Shared host/device helpers for the b200seg kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200seg.h"

namespace b200 {

void set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      b200::set_error(__VA_ARGS__);    \
      return B200SEG_ERR_INVALID;      \
    }                                  \
  } while (0)

#define B200_CHECK_LAUNCH(name)                                                 \
  do {                                                                          \
    cudaError_t e_ = cudaGetLastError();                                        \
    if (e_ != cudaSuccess) {                                                    \
      b200::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));   \
      return B200SEG_ERR_CUDA;                                                  \
    }                                                                           \
  } while (0)

constexpr int kNumSMs = 148;

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

// 128-bit accesses spelled as uint4: a struct-of-bfloat162 copy is scalarised by nvcc into four 32-bit LDG/STG.
__device__ __forceinline__ bf16x8 ld8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  bf16x8 v;
  *reinterpret_cast<uint4*>(&v) = u;
  return v;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const bf16x8& v) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}

// 256-bit store of 16 bf16 (sm_100: STG.E.256): one L2 request instead of two for a 32-byte piece of an output row.
// p must be 32-byte aligned.
__device__ __forceinline__ void st16(__nv_bfloat16* p, const bf16x8& a, const bf16x8& b) {
  const uint4 u = *reinterpret_cast<const uint4*>(&a), w = *reinterpret_cast<const uint4*>(&b);
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w),
               "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w)
               : "memory");
}

__device__ __forceinline__ void unpack8(const bf16x8& v, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(v.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
  bf16x8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// activation and its derivative as a function of the pre-activation value
__device__ __forceinline__ float act_fwd(float x, int act, float slope) {
  switch (act) {
    case B200SEG_ACT_RELU: return x > 0.f ? x : 0.f;
    case B200SEG_ACT_LEAKY:
    case B200SEG_ACT_PRELU: return x > 0.f ? x : slope * x;
    case B200SEG_ACT_ELU: return x > 0.f ? x : expm1f(x);
    default: return x;
  }
}
__device__ __forceinline__ float act_bwd(float x, int act, float slope) {
  switch (act) {
    case B200SEG_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case B200SEG_ACT_LEAKY:
    case B200SEG_ACT_PRELU: return x > 0.f ? 1.f : slope;
    case B200SEG_ACT_ELU: return x > 0.f ? 1.f : expf(x);
    default: return 1.f;
  }
}

inline int grid_for(int64_t work_items, int threads, int max_blocks = kNumSMs * 16) {
  int64_t b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}

}  // namespace b200
