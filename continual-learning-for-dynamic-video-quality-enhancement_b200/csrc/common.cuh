// Shared device/host helpers for libnervecl (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "nervecl.h"

#define NV_API extern "C" __attribute__((visibility("default")))

namespace nv {

// Tuning / profiling knobs (NERVECL_* environment variables) are read only by libraries built with -DNERVECL_TUNING;
// the default build has no environment-dependent behaviour at all.
#ifdef NERVECL_TUNING
inline const char* tune_env(const char* name) { return getenv(name); }
#else
inline const char* tune_env(const char*) { return nullptr; }
#endif


typedef __nv_bfloat16 bf16;

static inline cudaStream_t as_stream(nervecl_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Return code after a launch: positive cudaError_t on failure.
static inline int launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the sticky-less launch error so the next call starts clean
    return (int)e;
  }
  return NERVECL_OK;
}

__host__ __device__ static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }
static inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

constexpr int kSMs = 148;  // B200

// ---- scalar element access -------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- 4-wide vector access (16 B for f32, 8 B for bf16) ---------------------------------
struct f4 { float v[4]; };

__device__ __forceinline__ f4 ld4(const float* p) {
  float4 t = *reinterpret_cast<const float4*>(p);
  return f4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ f4 ld4(const bf16* p) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return f4{{fa.x, fa.y, fb.x, fb.y}};
}
__device__ __forceinline__ void st4(float* p, const f4& x) {
  *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
}
__device__ __forceinline__ void st4(bf16* p, const f4& x) {
  __nv_bfloat162 a = __floats2bfloat162_rn(x.v[0], x.v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(x.v[2], x.v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// ---- 8-wide vector access (2 x 16 B for f32, 16 B for bf16) ----------------------------
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld8(const float* p) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  return f8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ f8 ld8(const bf16* p) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  f8 r;
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {          // bf16 -> fp32 is a 16-bit shift: one ALU op per element
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
  return r;
}
__device__ __forceinline__ void st8(float* p, const f8& x) {
  *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const f8& x) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(x.v[2 * i], x.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

// ---- 16-wide store: one 256-bit access for bf16 (sm_100; p must be 32-byte aligned), 4 x 128-bit for f32 -------
// (a thread-per-pixel kernel that writes a pixel as 16-byte pieces touches half a 32-byte sector per instruction)
struct f16v { float v[16]; };
__device__ __forceinline__ void st16(bf16* p, const f16v& x) {
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x.v[2 * i], x.v[2 * i + 1]);
    r[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void st16(float* p, const f16v& x) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(p + 4 * i) = make_float4(x.v[4 * i], x.v[4 * i + 1], x.v[4 * i + 2], x.v[4 * i + 3]);
}

// ---- reductions ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Dispatch a templated launcher on activation dtype.
#define NV_DISPATCH_DTYPE(dtype, T, ...)            \
  do {                                              \
    if ((dtype) == NERVECL_F32) {                   \
      typedef float T;                              \
      __VA_ARGS__;                                  \
    } else if ((dtype) == NERVECL_BF16) {           \
      typedef nv::bf16 T;                           \
      __VA_ARGS__;                                  \
    } else {                                        \
      return NERVECL_EDTYPE;                        \
    }                                               \
  } while (0)

}  // namespace nv
