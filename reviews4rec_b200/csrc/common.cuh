// common.cuh -- shared helpers for the r4r_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/r4r_b200.h"

void r4r_set_error(const char* fmt, ...);

#define R4R_REQUIRE(cond, code, ...)                         \
  do {                                                       \
    if (!(cond)) {                                           \
      r4r_set_error(__VA_ARGS__);                            \
      return (code);                                         \
    }                                                        \
  } while (0)

// Report launch-configuration errors of the kernel just launched (asynchronous faults surface on
// the caller's next synchronisation, as with any CUDA library).
#define R4R_CHECK_LAUNCH(name)                                                       \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      r4r_set_error("%s: %s", name, cudaGetErrorString(e__));                        \
      return (int)e__;                                                               \
    }                                                                                \
  } while (0)

#define R4R_CUDA(call)                                                               \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      r4r_set_error("%s: %s", #call, cudaGetErrorString(e__));                       \
      return (int)e__;                                                               \
    }                                                                                \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// float <-> order-preserving uint32 (for packed (value,position) atomicMax keys)
__device__ __forceinline__ uint32_t f32_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// vector reductions into global memory (sm_90+): one L2 atomic transaction for 2 / 4 consecutive floats
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(a), "f"(b) : "memory");
}

// streaming 128-bit global accesses (Guideline 13: L1::no_allocate for data touched once)
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
