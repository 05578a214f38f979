// Shared host/device helpers for libdeeprec_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "deeprec_b200.h"

#ifndef __CUDA_ARCH_LIST__
#define __CUDA_ARCH_LIST__ 1000
#endif

namespace dr {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DR_REQUIRE(cond, code, ...)              \
  do {                                           \
    if (!(cond)) {                               \
      dr::set_error(__VA_ARGS__);                \
      return (code);                             \
    }                                            \
  } while (0)

#define DR_CUDA_LAUNCH_CHECK(name)                                            \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      dr::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
      return (int)e__;                                                        \
    }                                                                         \
    dr::count_launch();                                                       \
  } while (0)

#define DR_CUDA_CALL(expr)                                                    \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      dr::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));         \
      return (int)e__;                                                        \
    }                                                                         \
  } while (0)

__host__ __device__ static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device helpers -------------------------------------------------------------------
#ifdef __CUDACC__

// 128-bit read-only load that does not allocate in L1 (random rows are touched once).
__device__ __forceinline__ float4 ldg_nc_na(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_nc_na_f32(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// L2 eviction-priority hints (experiment, knob embed_l2_hints): the random table rows are dead after one use
// (evict_first), the stacked output is re-read by the next kernel (evict_last), the backward's atomically updated lines
// should drain to HBM early instead of staying dirty in L2 until the next kernel's stores push them out.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_nc_na_hint(const float* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float ldg_nc_na_f32_hint(const float* p, uint64_t pol) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void stg4_hint(float* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void red_add_v4_hint(float* p, float4 v, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void red_add_f32_hint(float* p, float v, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
// 128-bit load through the normal (L1-allocating) path.
__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// streaming 128-bit store (write-once data that the next kernel reads from L2/HBM).
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// 128-bit vector reduction into global memory (sm_90+): one REDG.E.ADD.F32x4 per 16 B.
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int W>
__device__ __forceinline__ float group_sum(float v) {
  // butterfly over the W (power of two, <= 32) lanes of an aligned lane group
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_shfl_down(float4 v, int delta) {
  return make_float4(__shfl_down_sync(0xffffffffu, v.x, delta), __shfl_down_sync(0xffffffffu, v.y, delta),
                     __shfl_down_sync(0xffffffffu, v.z, delta), __shfl_down_sync(0xffffffffu, v.w, delta));
}

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case DR_ACT_RELU: return fmaxf(v, 0.f);
    case DR_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case DR_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
// derivative expressed through the OUTPUT y = act(z)
__device__ __forceinline__ float act_grad_from_y(float y, int act) {
  switch (act) {
    case DR_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case DR_ACT_SIGMOID: return y * (1.f - y);
    case DR_ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}
#endif  // __CUDACC__

}  // namespace dr
