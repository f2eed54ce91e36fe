// Fused epilogue shared by the tcgen05 convolution kernels: one thread = one output pixel (TMEM lane), 16
// consecutive output channels per chunk.  v = acc; +bias; ReLU (relu == 1); *alpha; +res; +out (accumulate); ReLU (relu == 2); ReLU-mask.
#pragma once
#include "tc_common.cuh"

namespace nv {
namespace tc {

struct EpiArgs {
  int Cout;
  int relu, accumulate, res_channels, mask_c0;
  float alpha;
  const float* bias;
  const bf16* res;  int64_t ldres;
  const bf16* mask; int64_t ldmask;
  const bf16* msub; int64_t ldmsub;
  void* out;        int64_t ldo;
  float* colsum;    // optional per-channel sum of the written values (row kernel, mask-gated lean epilogue only)
  int v256_out;     // out rows / channel offsets are 32-byte aligned: one 256-bit store per 16-channel chunk
  int v256_in;      // same for the mask / residual operand of the lean epilogues (one 256-bit load)
  int v256_gen;     // generic epilogue: out, res, mask and mask_sub all allow 256-bit accesses
};

// 16 consecutive bf16 (two 16-byte loads) -> fp32
__device__ __forceinline__ void unpack16(const uint4 (&r)[2], float (&f)[16]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&r[h]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(p[i]);
      f[h * 8 + 2 * i] = t.x;
      f[h * 8 + 2 * i + 1] = t.y;
    }
  }
}
__device__ __forceinline__ void load16(const bf16* p, uint4 (&r)[2]) {
  r[0] = *reinterpret_cast<const uint4*>(p);
  r[1] = *reinterpret_cast<const uint4*>(p + 8);
}
__device__ __forceinline__ void store16(bf16* p, const float (&f)[16]) {
  uint4 r[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&r[h]);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = __floats2bfloat162_rn(f[h * 8 + 2 * i], f[h * 8 + 2 * i + 1]);
  }
  *reinterpret_cast<uint4*>(p) = r[0];
  *reinterpret_cast<uint4*>(p + 8) = r[1];
}
// 32 bytes (16 bf16) per thread: one full 32-byte sector per instruction when the address allows it (sm_100
// 256-bit global accesses); the two 16-byte halves of a thread otherwise go out as separate half-sector writes
__device__ __forceinline__ void store16(bf16* p, const uint32_t (&r)[8], bool v256) {
  if (!v256) {
    *reinterpret_cast<uint4*>(p) = make_uint4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<uint4*>(p + 8) = make_uint4(r[4], r[5], r[6], r[7]);
    return;
  }
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void pack16(const float (&f)[16], uint32_t (&r)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    r[i] = *reinterpret_cast<uint32_t*>(&h);
  }
}
__device__ __forceinline__ void store16(bf16* p, const float (&f)[16], bool v256) {
  if (!v256) { store16(p, f); return; }
  uint32_t r[8];
  pack16(f, r);
  store16(p, r, true);
}
// Packed ReLU signs of 16 NON-NEGATIVE bf16 (eight bf16x2 words): channel j <-> bit sign_bit_pos(j).  Adding 0x7FFF to a
// non-negative half sets its bit 15 exactly when it is not zero (no carry into the upper half), so a word costs three
// integer instructions instead of a compare / select / shift per channel.
__device__ __forceinline__ uint32_t sign_word16(const uint32_t (&r)[8]) {
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc |= ((r[k] + 0x7FFF7FFFu) & 0x80008000u) >> k;
  return ((acc >> 8) & 0xFFu) | ((acc >> 16) & 0xFF00u);
}
__device__ __forceinline__ constexpr int sign_bit_pos(int j) { return (j & 1) ? 15 - (j >> 1) : 7 - (j >> 1); }
// 32 bytes of an epilogue operand (16 bf16 of one pixel).  L2::64B: the request fills the whole 64-byte DRAM burst into
// L2.  The two epilogue warps that share a pixel's 32-channel slice each ask for one 32-byte sector of it; without the
// hint DRAM was read twice per burst (ncu: 64 channel-equivalents per pixel for a 32-channel ReLU mask, 9 GB per step).
__device__ __forceinline__ void load32B(const bf16* p, uint4& lo, uint4& hi, bool v256) {
  if (v256) {
    asm volatile("ld.global.L2::64B.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
  } else {
    asm volatile("ld.global.L2::64B.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "l"(p));
    asm volatile("ld.global.L2::64B.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(p + 8));
  }
}
__device__ __forceinline__ void store16(float* p, const float (&f)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(p + 4 * i) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}
__device__ __forceinline__ void store16(float* p, const float (&f)[16], bool) { store16(p, f); }

// One 16-column chunk of the epilogue in two phases so that the TMEM load and every global load of
// (up to) two chunks are in flight together before anything is consumed.
template <typename OutT>
struct EpiChunk {
  uint32_t v[16];
  uint4 acc[2], res[2], msk[2], sub[2];
  bool vec, has_res, has_mask;

  // have_pm: the 16 mask values of this chunk were already fetched by the caller (prefetched rows ahead)
  __device__ __forceinline__ void issue(const EpiArgs& a, uint32_t taddr, int c0, bool valid, int64_t p,
                                        bool have_pm = false, uint4 pm0 = uint4{0, 0, 0, 0}, uint4 pm1 = uint4{0, 0, 0, 0}) {
    tmem_ld16(taddr + (uint32_t)c0, v);
    // fast path: the chunk lies fully inside every channel range it touches
    vec = valid && (c0 + 16 <= a.Cout) && !(a.res && c0 < a.res_channels && c0 + 16 > a.res_channels) &&
          !(a.mask && c0 < a.mask_c0 && c0 + 16 > a.mask_c0);
    has_res = a.res && c0 < a.res_channels;
    has_mask = a.mask && c0 >= a.mask_c0;
    if (vec) {
      const bool v = a.v256_gen != 0;
      if (a.accumulate) load32B(reinterpret_cast<const bf16*>(a.out) + p * a.ldo + c0, acc[0], acc[1], v);
      if (has_res) load32B(a.res + p * a.ldres + c0, res[0], res[1], v);
      if (has_mask) {
        if (have_pm) { msk[0] = pm0; msk[1] = pm1; }
        else load32B(a.mask + p * a.ldmask + c0, msk[0], msk[1], v);
        if (a.msub) load32B(a.msub + p * a.ldmsub + c0, sub[0], sub[1], v);
      }
    }
  }

  __device__ __forceinline__ void finish(const EpiArgs& a, int c0, bool valid, int64_t p) {
    if (!valid || c0 >= a.Cout) return;
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
    OutT* op = reinterpret_cast<OutT*>(a.out) + p * a.ldo + c0;
    if (vec) {
      if (a.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(a.bias + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 b = __ldg(b4 + i);
          f[4 * i] += b.x; f[4 * i + 1] += b.y; f[4 * i + 2] += b.z; f[4 * i + 3] += b.w;
        }
      }
      if (a.relu == 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] *= a.alpha;
      float t[16];
      if (has_res) {
        unpack16(res, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] += t[j];
      }
      if (a.accumulate) {
        unpack16(acc, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] += t[j];
      }
      if (a.relu == 2) {                      // ReLU AFTER the residual (ResidualBlock: relu(conv(..) + identity))
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      if (has_mask) {
        unpack16(msk, t);
        if (a.msub) {
          float u[16];
          unpack16(sub, u);
#pragma unroll
          for (int j = 0; j < 16; ++j) t[j] -= u[j];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = t[j] > 0.f ? f[j] : 0.f;
      }
      store16(op, f, a.v256_gen != 0);
    } else {
      // ragged chunk (Cout tail, or a range boundary inside the chunk): element-wise
      for (int j = 0; j < 16; ++j) {
        const int c = c0 + j;
        if (c >= a.Cout) break;
        float x = f[j];
        if (a.bias) x += __ldg(a.bias + c);
        if (a.relu == 1) x = fmaxf(x, 0.f);
        x *= a.alpha;
        if (a.res && c < a.res_channels) x += ldf(a.res + p * a.ldres + c);
        if (a.accumulate) x += ldf(reinterpret_cast<const bf16*>(a.out) + p * a.ldo + c);
        if (a.relu == 2) x = fmaxf(x, 0.f);
        if (a.mask && c >= a.mask_c0) {
          float m = ldf(a.mask + p * a.ldmask + c);
          if (a.msub) m -= ldf(a.msub + p * a.ldmsub + c);
          if (!(m > 0.f)) x = 0.f;
        }
        stf(op + j, x);
      }
    }
  }
};


}  // namespace tc
}  // namespace nv
