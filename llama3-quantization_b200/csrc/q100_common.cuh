// q100_common.cuh -- shared device helpers: storage-dtype traits, per-op rounding that
// reproduces torch's "compute in fp32, round to the storage dtype after every op"
// semantics (SURVEY.md §4 recipe), vector load/store, error plumbing.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/q100.h"

namespace q100 {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int ensure_device();   // Q100_OK iff current device is sm_100; caches the answer
int sm_count();
int current_device();   // index of the calling thread's current device in [0, 64), 0 on error
constexpr int kMaxDevices = 64;   // per-device caches (function attributes are a per-device property)

#define Q100_CUDA(call)                                            \
  do {                                                             \
    int _rc = ::q100::check_cuda((call), #call);                   \
    if (_rc != Q100_OK) return _rc;                                \
  } while (0)

#define Q100_REQUIRE(cond, code, ...)                              \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::q100::set_error(__VA_ARGS__);                              \
      return (code);                                               \
    }                                                              \
  } while (0)

// ---------------------------------------------------------------- dtype traits
template <typename T> struct DT;

template <> struct DT<__half> {
  static constexpr int kCode = Q100_F16;
  static constexpr int kVec = 8;          // elements per 16-byte vector
  using T2 = __half2;
  __device__ static __forceinline__ float to_f(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  __device__ static __forceinline__ float rd(float v) { return __half2float(__float2half_rn(v)); }
  __device__ static __forceinline__ float2 to_f2(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
  }
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

template <> struct DT<__nv_bfloat16> {
  static constexpr int kCode = Q100_BF16;
  static constexpr int kVec = 8;
  using T2 = __nv_bfloat162;
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  __device__ static __forceinline__ float rd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
  __device__ static __forceinline__ float2 to_f2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

template <> struct DT<float> {
  static constexpr int kCode = Q100_F32;
  static constexpr int kVec = 4;
  __device__ static __forceinline__ float to_f(float v) { return v; }
  __device__ static __forceinline__ float from_f(float v) { return v; }
  __device__ static __forceinline__ float rd(float v) { return v; }
};

// ---------------------------------------------------------------- memory helpers
// Streaming 128-bit accesses: the quantize kernels touch every byte exactly once, so keep
// them out of L1 (guide G13/G14).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float nan_f() { return __int_as_float(0x7fc00000); }

// torch.clamp(x, lo, hi): NaN propagates (fminf/fmaxf would drop it).
__device__ __forceinline__ float clamp_nan(float x, float lo, float hi) {
  float c = fminf(fmaxf(x, lo), hi);
  return (x != x) ? x : c;
}

// NaN-propagating min / max (torch.amin / torch.amax semantics).
__device__ __forceinline__ float min_nan(float a, float b) {
  float m = fminf(a, b);
  return (a != a || b != b) ? nan_f() : m;
}
__device__ __forceinline__ float max_nan(float a, float b) {
  float m = fmaxf(a, b);
  return (a != a || b != b) ? nan_f() : m;
}

}  // namespace q100
