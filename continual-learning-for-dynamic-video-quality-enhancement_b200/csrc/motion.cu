// Motion estimation / compensation kernels: 81-displacement correlation and the flow warp.
#include "common.cuh"
#include <cstdlib>

using namespace nv;

namespace nv {   // motion_mma.cu
// tcgen05 2-D tile Gram kernel (motion_tc.cu)
bool corr_fwd_tc_supported(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* out, int64_t ldo, int H, int W,
                           int cout_pad);
int corr_fwd_tc(const void* x1, int64_t ld1, const void* x2, int64_t ld2, void* out, int64_t ldo, int N, int H, int W,
                int cout_pad, cudaStream_t s);
bool corr_bwd_tc_supported(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* g, int64_t ldg, const void* dx1,
                           int64_t lddx1, const void* dx2, int64_t lddx2, const void* ws, int64_t ws_bytes, int N, int H, int W);
int corr_bwd_tc(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* g, int64_t ldg, void* dx1, int64_t lddx1,
                int acc1, void* dx2, int64_t lddx2, int acc2, void* ws, int N, int H, int W, cudaStream_t s);
int corr_fwd_mma(const void* x1, int64_t ld1, const void* x2, int64_t ld2, void* out, int64_t ldo, int N, int H, int W,
                 int cout_pad, cudaStream_t s);
int corr_bwd_mma(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* g, int64_t ldg, void* dx1,
                 int64_t lddx1, int acc1, void* dx2, int64_t lddx2, int acc2, int N, int H, int W, cudaStream_t s);
}
namespace nv {   // motion_tiled.cu
bool corr_tiled_supported(int dtype, int C, int64_t ld1, int64_t ld2, const void* x1, const void* x2);
int corr_fwd_tiled(const void* x1, int64_t ld1, const void* x2, int64_t ld2, void* out, int64_t ldo, int N, int H, int W,
                   int cout_pad, cudaStream_t s);
int corr_bwd_tiled(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* g, int64_t ldg, void* dx1,
                   int64_t lddx1, int acc1, void* dx2, int64_t lddx2, int acc2, int N, int H, int W, cudaStream_t s);
}

namespace {

constexpr int RAD = 4, ND = 9, NDISP = 81;

// ---------------------------------------------------------------------------------------
// correlation forward.  block = 32 pixels (one row strip) x 9 displacement rows; a thread
// computes the 9 horizontal displacements of its (pixel, dy) pair, streaming channels 8 at a time.
// Results are staged in shared memory and written as full pixel rows (coalesced).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(288)
corr_fwd_kernel(const T* __restrict__ x1, int64_t ld1, const T* __restrict__ x2, int64_t ld2,
                T* __restrict__ out, int64_t ldo, int N, int H, int W, int C, int cout_pad) {
  extern __shared__ float stage[];  // [32][cout_pad]
  const int strips = (W + 31) / 32;
  const int64_t row = blockIdx.x / strips;  // n*H + y
  const int x0 = (int)(blockIdx.x % strips) * 32;
  const int y = (int)(row % H);
  const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
  const int x = x0 + lane;
  const int sy = y + i - RAD;
  float acc[ND];
#pragma unroll
  for (int j = 0; j < ND; ++j) acc[j] = 0.f;
  if (x < W && sy >= 0 && sy < H) {
    const T* a = x1 + (row * W + x) * ld1;
    const T* brow = x2 + ((row + (i - RAD)) * W) * ld2;
    for (int c = 0; c < C; c += 8) {
      f8 av = ld8(a + c);
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        int sx = x + j - RAD;
        if (sx < 0 || sx >= W) continue;
        f8 bv = ld8(brow + (int64_t)sx * ld2 + c);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[j] = fmaf(av.v[k], bv.v[k], acc[j]);
      }
    }
  }
  const float inv_c = 1.f / (float)C;
#pragma unroll
  for (int j = 0; j < ND; ++j) stage[lane * cout_pad + i * ND + j] = acc[j] * inv_c;
  for (int e = threadIdx.x; e < 32 * (cout_pad - NDISP); e += blockDim.x) {
    int p = e / (cout_pad - NDISP), c = NDISP + e % (cout_pad - NDISP);
    stage[p * cout_pad + c] = 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * cout_pad; e += blockDim.x) {
    int p = e / cout_pad, c = e % cout_pad;
    if (x0 + p < W) stf(out + (row * W + x0 + p) * ldo + c, stage[e]);
  }
}

// ---------------------------------------------------------------------------------------
// correlation backward.  thread = (pixel, 8 channels).
//   dx1[p,c] = 1/C sum_d g[p,d]   * x2[p+d,c]
//   dx2[q,c] = 1/C sum_d g[q-d,d] * x1[q-d,c]
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
corr_bwd_kernel(const T* __restrict__ x1, int64_t ld1, const T* __restrict__ x2, int64_t ld2,
                const T* __restrict__ g, int64_t ldg, T* __restrict__ dx1, int64_t lddx1, int acc1,
                T* __restrict__ dx2, int64_t lddx2, int acc2, int N, int H, int W, int C) {
  const int cg = C >> 3;
  const int64_t total = (int64_t)N * H * W * cg;
  const float inv_c = 1.f / (float)C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int c0 = (int)(t % cg) << 3;
    int64_t p = t / cg;
    int x = (int)(p % W), y = (int)((p / W) % H);
    float a1[8], a2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a1[k] = a2[k] = 0.f;
    for (int i = 0; i < ND; ++i) {
      int dy = i - RAD;
      for (int j = 0; j < ND; ++j) {
        int dx = j - RAD;
        int d = i * ND + j;
        // dx1: neighbour p+d of x2
        int sy = y + dy, sx = x + dx;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
          float gv = ldf(g + p * ldg + d);
          f8 v = ld8(x2 + (p + (int64_t)dy * W + dx) * ld2 + c0);
#pragma unroll
          for (int k = 0; k < 8; ++k) a1[k] = fmaf(gv, v.v[k], a1[k]);
        }
        // dx2: source pixel q-d of x1
        int ty = y - dy, tx = x - dx;
        if (ty >= 0 && ty < H && tx >= 0 && tx < W) {
          int64_t q = p - (int64_t)dy * W - dx;
          float gv = ldf(g + q * ldg + d);
          f8 v = ld8(x1 + q * ld1 + c0);
#pragma unroll
          for (int k = 0; k < 8; ++k) a2[k] = fmaf(gv, v.v[k], a2[k]);
        }
      }
    }
    f8 o1, o2;
    if (acc1) o1 = ld8(dx1 + p * lddx1 + c0); else { for (int k = 0; k < 8; ++k) o1.v[k] = 0.f; }
    if (acc2) o2 = ld8(dx2 + p * lddx2 + c0); else { for (int k = 0; k < 8; ++k) o2.v[k] = 0.f; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      o1.v[k] += a1[k] * inv_c;
      o2.v[k] += a2[k] * inv_c;
    }
    st8(dx1 + p * lddx1 + c0, o1);
    st8(dx2 + p * lddx2 + c0, o2);
  }
}

// ---------------------------------------------------------------------------------------
// warp.  Coordinates replay the reference's fp32 op sequence with explicit round-to-nearest
// intrinsics (no FMA contraction, no re-association) -> neighbour indices bit-identical to ATen.
// ---------------------------------------------------------------------------------------
struct WarpCoord {
  float ix, iy;     // un-normalised sample position
  float x0f, y0f;   // floor
  int x0, y0;
};

__device__ __forceinline__ float unnorm_coord(int pos, float flow, int size, float inv, int div_mode) {
  float g = __fadd_rn((float)pos, flow);               // grid + flow            (super_resolution.py:126)
  g = __fmul_rn(2.0f, g);                              // 2.0 * grid             (:129)
  g = div_mode ? __fdiv_rn(g, (float)(size - 1))       // / (W-1): ATen-CPU divides,
               : __fmul_rn(g, inv);                    //          ATen-CUDA multiplies by 1/(W-1)
  g = __fsub_rn(g, 1.0f);                              // - 1.0
  // grid_sampler_unnormalize(align_corners=True): ((coord + 1) / 2) * (size - 1)
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
}

__device__ __forceinline__ WarpCoord warp_coord(int x, int y, float fx, float fy, int W, int H, float inv_w,
                                                float inv_h, int div_mode) {
  WarpCoord c;
  c.ix = unnorm_coord(x, fx, W, inv_w, div_mode);
  c.iy = unnorm_coord(y, fy, H, inv_h, div_mode);
  c.x0f = floorf(c.ix);
  c.y0f = floorf(c.iy);
  // clamp before the int conversion so huge flows cannot overflow (they are out of bounds anyway)
  c.x0 = (int)fminf(fmaxf(c.x0f, -2.f), (float)W + 1.f);
  c.y0 = (int)fminf(fmaxf(c.y0f, -2.f), (float)H + 1.f);
  return c;
}

// eight elements as loaded (bf16: one 16-byte register quad; converted to fp32 where they are used, so that a
// prefetched pixel costs 4 registers per vector instead of 8)
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 u;
  __device__ __forceinline__ static Raw8 zero() { Raw8 r; r.u = make_uint4(0, 0, 0, 0); return r; }
  __device__ __forceinline__ f8 f() const {
    f8 r;
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r.v[2 * i] = __uint_as_float(w[i] << 16);
      r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
    return r;
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ static Raw8 zero() { Raw8 r; r.a = r.b = make_float4(0.f, 0.f, 0.f, 0.f); return r; }
  __device__ __forceinline__ f8 f() const { return f8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}}; }
};
__device__ __forceinline__ Raw8<bf16> ldraw(const bf16* p) { Raw8<bf16> r; r.u = *reinterpret_cast<const uint4*>(p); return r; }
__device__ __forceinline__ Raw8<float> ldraw(const float* p) {
  Raw8<float> r;
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
  return r;
}

constexpr int WP_W = 8, WP_H = 4;                  // pixel patch of a warp (WP_W * WP_H == 32)

// Forward warp.  The first version gave each pixel C/8 adjacent lanes that ALL replayed the ~70-instruction
// coordinate sequence: 69 warp instructions per pixel, instruction-issue bound at 0.40 of the HBM roofline
// (profiles/r01z_summary.md).  Here a warp owns 32 consecutive pixels: phase A computes one pixel per LANE
// (coalesced flow load, coordinates, corner weights, validity), phase B walks the 32 pixels 32/cg at a time with
// the lanes regrouped as (pixel, 8-channel vector) and the pixel's parameters broadcast by shuffles.
template <typename T, int CGT>                     // CGT: C / 8 at compile time (the pixel loop unrolls fully), 0 = run time
__global__ void __launch_bounds__(256)
warp_fwd_kernel(const T* __restrict__ feat, int64_t ldf_, const float* __restrict__ flow, T* __restrict__ out,
                int64_t ldo, int N, int H, int W, int C, float inv_w, float inv_h, int div_mode,
                int32_t* __restrict__ idx_out) {
  const int lane = threadIdx.x & 31;
  // a warp owns a WP_H x WP_W pixel patch (not 32 pixels of one row): the south corners of one patch row are the north
  // corners of the next, so they hit in L1 instead of travelling from L2 once per row
  const int ppx = (W + WP_W - 1) / WP_W, ppy = (H + WP_H - 1) / WP_H;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int patch_x = wid % ppx, pr = wid / ppx;
  const int n_img = pr / ppy;
  if (n_img >= N) return;                                            // (whole warps)
  const int x = patch_x * WP_W + lane % WP_W, y = (pr - n_img * ppy) * WP_H + lane / WP_W;
  const int p = (n_img * H + y) * W + x;                             // (< 2^31: checked by the launcher)
  // ---- phase A: lane = pixel ----
  float wnw = 0.f, wne = 0.f, wsw = 0.f, wse = 0.f;
  int q = 0;                                                         // pixel index of the north-west corner
  unsigned vmask = 0;                                                // bit0 nw, bit1 ne, bit2 sw, bit3 se, bit4 active
  if (x < W && y < H) {
    const int r = n_img * H + y;
    const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + p);
    const WarpCoord c = warp_coord(x, y, f.x, f.y, W, H, inv_w, inv_h, div_mode);
    if (idx_out) {
      idx_out[(int64_t)p * 2] = (int)c.x0f;
      idx_out[(int64_t)p * 2 + 1] = (int)c.y0f;
    }
    const float x1f = c.x0f + 1.f, y1f = c.y0f + 1.f;
    wnw = (x1f - c.ix) * (y1f - c.iy);
    wne = (c.ix - c.x0f) * (y1f - c.iy);
    wsw = (x1f - c.ix) * (c.iy - c.y0f);
    wse = (c.ix - c.x0f) * (c.iy - c.y0f);
    const bool xin0 = c.x0 >= 0 && c.x0 < W, xin1 = c.x0 + 1 >= 0 && c.x0 + 1 < W;
    const bool yin0 = c.y0 >= 0 && c.y0 < H, yin1 = c.y0 + 1 >= 0 && c.y0 + 1 < H;
    vmask = (yin0 && xin0 ? 1u : 0u) | (yin0 && xin1 ? 2u : 0u) | (yin1 && xin0 ? 4u : 0u) | (yin1 && xin1 ? 8u : 0u) | 16u;
    q = (r - y) * W + c.y0 * W + c.x0;                               // n*H*W + y0*W + x0 (only used where valid)
  }
  // ---- phase B: lane = (pixel of the group, 8-channel vector) ----
  // The four corner vectors of the NEXT group of pixels are requested (as raw 16-byte registers) before the current
  // group is blended and stored: two groups of loads in flight per lane instead of one.
  const int cg = CGT ? CGT : (C >> 3);                               // lanes per pixel (power of two <= 32)
  const int ppi = 32 / cg;                                           // pixels per iteration
  const int c0 = (lane % cg) << 3;
  const int sub = lane / cg;
  struct Item {
    Raw8<T> v0, v1, v2, v3;
    float a, b, c, d;
    unsigned vm;
    int pp;                                                          // the pixel's linear index
  };
  auto fetch = [&](int it, Item& I) {
    const int src = it * ppi + sub;
    I.a = __shfl_sync(0xffffffffu, wnw, src); I.b = __shfl_sync(0xffffffffu, wne, src);
    I.c = __shfl_sync(0xffffffffu, wsw, src); I.d = __shfl_sync(0xffffffffu, wse, src);
    const int qq = __shfl_sync(0xffffffffu, q, src);
    I.vm = __shfl_sync(0xffffffffu, vmask, src);
    I.pp = __shfl_sync(0xffffffffu, p, src);
    const T* base = feat + (int64_t)qq * ldf_ + c0;
    I.v0 = (I.vm & 1u) ? ldraw(base) : Raw8<T>::zero();
    I.v1 = (I.vm & 2u) ? ldraw(base + ldf_) : Raw8<T>::zero();
    I.v2 = (I.vm & 4u) ? ldraw(base + (int64_t)W * ldf_) : Raw8<T>::zero();
    I.v3 = (I.vm & 8u) ? ldraw(base + (int64_t)(W + 1) * ldf_) : Raw8<T>::zero();
  };
  Item cur;
  fetch(0, cur);
#pragma unroll
  for (int it = 0; it < cg; ++it) {
    Item nxt = cur;
    if (it + 1 < cg) fetch(it + 1, nxt);                             // (warp-uniform: the shuffles inside are full-warp)
    if (cur.vm & 16u) {
      const f8 v0 = cur.v0.f(), v1 = cur.v1.f(), v2 = cur.v2.f(), v3 = cur.v3.f();
      f8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) {                                  // (nw, ne, sw, se; fused multiply-adds from zero)
        float acc = fmaf(v0.v[k], cur.a, 0.f);
        acc = fmaf(v1.v[k], cur.b, acc);
        acc = fmaf(v2.v[k], cur.c, acc);
        acc = fmaf(v3.v[k], cur.d, acc);
        o.v[k] = acc;
      }
      st8(out + (int64_t)cur.pp * ldo + c0, o);
    }
    cur = nxt;
  }
}

// blockDim multiple of 32; the cg (= C/8, power of two <= 32) threads of one pixel are adjacent
// lanes, so the flow gradient is reduced with xor-shuffles.
// 8 consecutive fp32 accumulations as two 128-bit vector reductions (red.global.add.v4.f32, sm_90+)
__device__ __forceinline__ void atomic_add8(float* p, float w, const f8& g) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(w * g.v[0], w * g.v[1], w * g.v[2], w * g.v[3]));
  atomicAdd(reinterpret_cast<float4*>(p) + 1, make_float4(w * g.v[4], w * g.v[5], w * g.v[6], w * g.v[7]));
}

// 8 consecutive bf16 accumulations as ONE 128-bit packed reduction (red.global.add.noftz.v4.bf16x2): half the
// L2 reduction operations of the fp32 form, which is what bounds the scatter (one 16-byte RED per L2 slice clock)
__device__ __forceinline__ void atomic_add8(bf16* p, float w, const f8& g) {
  uint32_t r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(w * g.v[2 * i], w * g.v[2 * i + 1]);
    r[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}

// the same reductions for an already-weighted contribution vector
__device__ __forceinline__ void red_add8(float* p, const f8& c) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(c.v[0], c.v[1], c.v[2], c.v[3]));
  atomicAdd(reinterpret_cast<float4*>(p) + 1, make_float4(c.v[4], c.v[5], c.v[6], c.v[7]));
}
__device__ __forceinline__ void red_add8(bf16* p, const f8& c) {
  uint32_t r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(c.v[2 * i], c.v[2 * i + 1]);
    r[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  // (no "memory" clobber: the reductions go to a buffer nothing in the kernel reads, and the clobber would pin every later
  //  load behind them -- one exposed DRAM latency per pixel of the run)
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]));
}

// Backward warp, same two phases as the forward (the C/8 lanes of a pixel used to replay the coordinate sequence
// each: issue-bound at 68 % with the L2 reductions idle a third of the time): phase A one pixel per LANE, phase B the
// lanes regrouped as (pixel, 8-channel vector) with the pixel's corner weights broadcast by shuffles; the flow
// gradient is reduced over the pixel's lanes with xor-shuffles.
template <typename T, typename AT, int CGT>
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const T* __restrict__ feat, int64_t ldf_, const float* __restrict__ flow,
                const T* __restrict__ dout, int64_t lddo, AT* __restrict__ dfeat, int64_t lddf,
                float* __restrict__ dflow, int N, int H, int W, int C, float inv_w, float inv_h, int div_mode) {
  const int lane = threadIdx.x & 31;
  const int npix = N * H * W;                                        // (< 2^31: checked by the launcher)
  const int p = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32 + lane;
  // ---- phase A: lane = pixel ----
  float wx0 = 0.f, wx1 = 0.f, wy0 = 0.f, wy1 = 0.f;                 // x1f - ix, ix - x0f, y1f - iy, iy - y0f
  int q = 0;
  unsigned vmask = 0;                                                // bit0 nw, bit1 ne, bit2 sw, bit3 se, bit4 active
  if (p < npix) {
    const int x = p % W, r = p / W;
    const int y = r % H;
    const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + p);
    const WarpCoord c = warp_coord(x, y, f.x, f.y, W, H, inv_w, inv_h, div_mode);
    const float x1f = c.x0f + 1.f, y1f = c.y0f + 1.f;
    wx0 = x1f - c.ix; wx1 = c.ix - c.x0f; wy0 = y1f - c.iy; wy1 = c.iy - c.y0f;
    const bool xin0 = c.x0 >= 0 && c.x0 < W, xin1 = c.x0 + 1 >= 0 && c.x0 + 1 < W;
    const bool yin0 = c.y0 >= 0 && c.y0 < H, yin1 = c.y0 + 1 >= 0 && c.y0 + 1 < H;
    vmask = (yin0 && xin0 ? 1u : 0u) | (yin0 && xin1 ? 2u : 0u) | (yin1 && xin0 ? 4u : 0u) | (yin1 && xin1 ? 8u : 0u) | 16u;
    q = (r - y) * W + c.y0 * W + c.x0;
  }
  // ---- phase B: lane = (run of consecutive pixels, 8-channel vector) ----
  // A lane group walks cg CONSECUTIVE pixels.  With a smooth flow the north-east / south-east targets of one pixel are the
  // north-west / south-west targets of the next, so those contributions are carried in registers and leave as ONE
  // reduction per target and source row (2 per pixel and vector instead of 4: the scatter is bound by L2 reduction
  // operations, profiles/r01z_summary.md).  Carries that do not meet their successor are flushed on their own.
  const int cg = CGT ? CGT : (C >> 3);
  const int c0 = (lane % cg) << 3;
  const int sub = lane / cg;
  const int pbase = p - lane;
  // d ix / d flow_x = ((W-1)/2) * (2 * 1/(W-1)), evaluated in the reference's order
  const float mx = ((float)(W - 1) * 0.5f) * (2.0f * inv_w);
  const float my = ((float)(H - 1) * 0.5f) * (2.0f * inv_h);
  f8 ct, cb;                                                         // carried north-east / south-east contributions
  int qt = -1, qb = -1;                                              // their target pixels (-1: none)
#pragma unroll
  for (int k = 0; k < 8; ++k) ct.v[k] = cb.v[k] = 0.f;
  // One pixel's operands, requested one iteration ahead of their use (the reductions of pixel i are issued while the loads
  // of pixel i + 1 are in flight; a straight loop exposes one DRAM latency per pixel of the run).
  struct Item {
    Raw8<T> g, v0, v1, v2, v3;
    float ax0, ax1, ay0, ay1;
    int qq;
    unsigned vm;
  };
  auto fetch = [&](int it, Item& I) {
    const int src = sub * cg + it;
    I.ax0 = __shfl_sync(0xffffffffu, wx0, src); I.ax1 = __shfl_sync(0xffffffffu, wx1, src);
    I.ay0 = __shfl_sync(0xffffffffu, wy0, src); I.ay1 = __shfl_sync(0xffffffffu, wy1, src);
    I.qq = __shfl_sync(0xffffffffu, q, src);
    I.vm = __shfl_sync(0xffffffffu, vmask, src);
    const T* fb = feat + (int64_t)I.qq * ldf_ + c0;
    I.g = (I.vm & 16u) ? ldraw(dout + (int64_t)(pbase + src) * lddo + c0) : Raw8<T>::zero();
    I.v0 = (I.vm & 1u) ? ldraw(fb) : Raw8<T>::zero();
    I.v1 = (I.vm & 2u) ? ldraw(fb + ldf_) : Raw8<T>::zero();
    I.v2 = (I.vm & 4u) ? ldraw(fb + (int64_t)W * ldf_) : Raw8<T>::zero();
    I.v3 = (I.vm & 8u) ? ldraw(fb + (int64_t)(W + 1) * ldf_) : Raw8<T>::zero();
  };
  Item cur;
  fetch(0, cur);
#pragma unroll
  for (int it = 0; it < cg; ++it) {
    Item nxt = cur;
    if (it + 1 < cg) fetch(it + 1, nxt);                             // (warp-uniform: the shuffles inside are full-warp)
    const int src = sub * cg + it;
    const float ax0 = cur.ax0, ax1 = cur.ax1, ay0 = cur.ay0, ay1 = cur.ay1;
    const int qq = cur.qq;
    const unsigned vm = cur.vm;
    float gix = 0.f, giy = 0.f;
    f8 nt, nb;
    int nqt = -1, nqb = -1;
#pragma unroll
    for (int k = 0; k < 8; ++k) nt.v[k] = nb.v[k] = 0.f;
    if (vm & 16u) {
      const f8 g = cur.g.f();
      if (vm & 1u) {
        const f8 v = cur.v0.f();
        const float w = ax0 * ay0;
        const bool merge = qt == qq;
        f8 c;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          c.v[k] = w * g.v[k] + (merge ? ct.v[k] : 0.f);
          gix -= v.v[k] * ay0 * g.v[k];
          giy -= v.v[k] * ax0 * g.v[k];
        }
        if (merge) qt = -1;
        red_add8(dfeat + (int64_t)qq * lddf + c0, c);
      }
      if (vm & 2u) {
        const f8 v = cur.v1.f();
        const float w = ax1 * ay0;
        nqt = qq + 1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          nt.v[k] = w * g.v[k];
          gix += v.v[k] * ay0 * g.v[k];
          giy -= v.v[k] * ax1 * g.v[k];
        }
      }
      if (vm & 4u) {
        const f8 v = cur.v2.f();
        const float w = ax0 * ay1;
        const bool merge = qb == qq + W;
        f8 c;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          c.v[k] = w * g.v[k] + (merge ? cb.v[k] : 0.f);
          gix -= v.v[k] * ay1 * g.v[k];
          giy += v.v[k] * ax0 * g.v[k];
        }
        if (merge) qb = -1;
        red_add8(dfeat + (int64_t)(qq + W) * lddf + c0, c);
      }
      if (vm & 8u) {
        const f8 v = cur.v3.f();
        const float w = ax1 * ay1;
        nqb = qq + W + 1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          nb.v[k] = w * g.v[k];
          gix += v.v[k] * ay1 * g.v[k];
          giy += v.v[k] * ax1 * g.v[k];
        }
      }
    }
    // carries that did not meet this pixel's north-west / south-west targets
    if (qt >= 0) red_add8(dfeat + (int64_t)qt * lddf + c0, ct);
    if (qb >= 0) red_add8(dfeat + (int64_t)qb * lddf + c0, cb);
    ct = nt; cb = nb; qt = nqt; qb = nqb;
    for (int o = cg >> 1; o > 0; o >>= 1) {
      gix += __shfl_xor_sync(0xffffffffu, gix, o);
      giy += __shfl_xor_sync(0xffffffffu, giy, o);
    }
    if ((vm & 16u) && (lane % cg) == 0)
      reinterpret_cast<float2*>(dflow)[pbase + src] = make_float2(gix * mx, giy * my);
    cur = nxt;
  }
  if (qt >= 0) red_add8(dfeat + (int64_t)qt * lddf + c0, ct);
  if (qb >= 0) red_add8(dfeat + (int64_t)qb * lddf + c0, cb);
}

}  // namespace

NV_API int nervecl_corr_fwd(const void* x1, int64_t ld1, const void* x2, int64_t ld2, void* out, int64_t ldo,
                            int dtype, int N, int H, int W, int C, int cout_pad, nervecl_stream_t stream) {
  if (!x1 || !x2 || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0) return NERVECL_EINVAL;
  if (cout_pad < NDISP || ldo < cout_pad) return NERVECL_EINVAL;
  if ((C & 7) || (ld1 & 7) || (ld2 & 7) || !aligned(x1, 16) || !aligned(x2, 16)) return NERVECL_EALIGN;
  if (dtype == NERVECL_BF16 && C == 64 && !nv::tune_env("NERVECL_CORR_MMA") &&
      corr_fwd_tc_supported(x1, ld1, x2, ld2, out, ldo, H, W, cout_pad))
    return corr_fwd_tc(x1, ld1, x2, ld2, out, ldo, N, H, W, cout_pad, as_stream(stream));
  if (corr_tiled_supported(dtype, C, ld1, ld2, x1, x2) && !(cout_pad & 7) && !(ldo & 7) && aligned(out, 16) && cout_pad <= 128)
    return (nv::tune_env("NERVECL_CORR_SIMT") ? corr_fwd_tiled : corr_fwd_mma)(x1, ld1, x2, ld2, out, ldo, N, H, W, cout_pad,
                                                                            as_stream(stream));
  int64_t blocks = (int64_t)N * H * cdiv(W, 32);
  size_t smem = (size_t)32 * cout_pad * sizeof(float);
  NV_DISPATCH_DTYPE(dtype, E, (corr_fwd_kernel<E><<<(unsigned)blocks, 288, smem, as_stream(stream)>>>(
                                  (const E*)x1, ld1, (const E*)x2, ld2, (E*)out, ldo, N, H, W, C, cout_pad)));
  return launch_status();
}

NV_API int nervecl_corr_bwd(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* dout,
                            int64_t lddo, void* dx1, int64_t lddx1, int acc1, void* dx2, int64_t lddx2, int acc2,
                            int dtype, int N, int H, int W, int C, void* workspace, int64_t workspace_bytes,
                            nervecl_stream_t stream) {
  if (!x1 || !x2 || !dout || !dx1 || !dx2 || N <= 0 || H <= 0 || W <= 0 || C <= 0) return NERVECL_EINVAL;
  if ((C & 7) || (ld1 & 7) || (ld2 & 7) || (lddx1 & 7) || (lddx2 & 7)) return NERVECL_EALIGN;
  if (dtype == NERVECL_BF16 && C == 64 && !nv::tune_env("NERVECL_CORR_MMA") &&
      corr_bwd_tc_supported(x1, ld1, x2, ld2, dout, lddo, dx1, lddx1, dx2, lddx2, workspace, workspace_bytes, N, H, W))
    return corr_bwd_tc(x1, ld1, x2, ld2, dout, lddo, dx1, lddx1, acc1, dx2, lddx2, acc2, workspace, N, H, W, as_stream(stream));
  if (corr_tiled_supported(dtype, C, ld1, ld2, x1, x2) && aligned(dx1, 16) && aligned(dx2, 16))
    return (nv::tune_env("NERVECL_CORR_SIMT") ? corr_bwd_tiled : corr_bwd_mma)(x1, ld1, x2, ld2, dout, lddo, dx1, lddx1, acc1, dx2,
                                                                            lddx2, acc2, N, H, W, as_stream(stream));
  int64_t total = (int64_t)N * H * W * (C >> 3);
  int blocks = (int)imax(1, imin(cdiv(total, 256), kSMs * 16));
  NV_DISPATCH_DTYPE(dtype, E, (corr_bwd_kernel<E><<<blocks, 256, 0, as_stream(stream)>>>(
                                  (const E*)x1, ld1, (const E*)x2, ld2, (const E*)dout, lddo, (E*)dx1, lddx1, acc1,
                                  (E*)dx2, lddx2, acc2, N, H, W, C)));
  return launch_status();
}

static int warp_check(int N, int H, int W, int C) {
  if (N <= 0 || H <= 1 || W <= 1 || C <= 0) return NERVECL_EINVAL;  // H or W == 1 divides by zero in the reference too
  if (C & 7) return NERVECL_EALIGN;
  int cg = C >> 3;
  if (cg > 32 || (cg & (cg - 1))) return NERVECL_EUNSUPPORTED;
  return NERVECL_OK;
}

NV_API int nervecl_warp_fwd(const void* feat, int64_t ldf, const float* flow, void* out, int64_t ldo, int dtype,
                            int N, int H, int W, int C, int div_mode, int32_t* idx_out, nervecl_stream_t stream) {
  if (!feat || !flow || !out) return NERVECL_EINVAL;
  int rc = warp_check(N, H, W, C);
  if (rc) return rc;
  if ((ldf & 7) || (ldo & 7) || !aligned(feat, 16) || !aligned(out, 16) || !aligned(flow, 8)) return NERVECL_EALIGN;
  float inv_w = 1.0f / (float)(W - 1), inv_h = 1.0f / (float)(H - 1);
  const int ppb = 256 / (C >> 3);
  if (cdiv(W, ppb) > 65535 || (int64_t)N * H > 0x7fffffff) return NERVECL_EUNSUPPORTED;
  if ((int64_t)N * H * W >= ((int64_t)1 << 31) - 64) return NERVECL_EUNSUPPORTED;
  const unsigned blocks = (unsigned)cdiv((int64_t)N * cdiv(H, WP_H) * cdiv(W, WP_W), 8);     // one warp per pixel patch
  if (C == 64) {
    NV_DISPATCH_DTYPE(dtype, E, (warp_fwd_kernel<E, 8><<<blocks, 256, 0, as_stream(stream)>>>(
                                    (const E*)feat, ldf, flow, (E*)out, ldo, N, H, W, C, inv_w, inv_h, div_mode, idx_out)));
  } else {
    NV_DISPATCH_DTYPE(dtype, E, (warp_fwd_kernel<E, 0><<<blocks, 256, 0, as_stream(stream)>>>(
                                    (const E*)feat, ldf, flow, (E*)out, ldo, N, H, W, C, inv_w, inv_h, div_mode, idx_out)));
  }
  return launch_status();
}

static int warp_bwd_launch(const void* feat, int64_t ldf, const float* flow, const void* dout, int64_t lddo,
                           void* dfeat, int64_t lddf, int dfeat_dtype, float* dflow, int dtype, int N, int H, int W, int C,
                           int div_mode, nervecl_stream_t stream) {
  if (!feat || !flow || !dout || !dfeat || !dflow) return NERVECL_EINVAL;
  int rc = warp_check(N, H, W, C);
  if (rc) return rc;
  if (dfeat_dtype != NERVECL_F32 && !(dfeat_dtype == NERVECL_BF16 && dtype == NERVECL_BF16)) return NERVECL_EDTYPE;
  if ((ldf & 7) || (lddo & 7) || (lddf & (dfeat_dtype == NERVECL_F32 ? 3 : 7)) || !aligned(feat, 16) || !aligned(dout, 16) ||
      !aligned(dfeat, 16) || !aligned(flow, 8) || !aligned(dflow, 8))
    return NERVECL_EALIGN;
  float inv_w = 1.0f / (float)(W - 1), inv_h = 1.0f / (float)(H - 1);
  if ((int64_t)N * H * W >= ((int64_t)1 << 31) - 64) return NERVECL_EUNSUPPORTED;
  const unsigned blocks = (unsigned)cdiv((int64_t)N * H * W, 256);
  if (dfeat_dtype == NERVECL_BF16) {
    if (C == 64)
      warp_bwd_kernel<bf16, bf16, 8><<<blocks, 256, 0, as_stream(stream)>>>((const bf16*)feat, ldf, flow, (const bf16*)dout, lddo,
                                                                            (bf16*)dfeat, lddf, dflow, N, H, W, C, inv_w, inv_h,
                                                                            div_mode);
    else
      warp_bwd_kernel<bf16, bf16, 0><<<blocks, 256, 0, as_stream(stream)>>>((const bf16*)feat, ldf, flow, (const bf16*)dout, lddo,
                                                                            (bf16*)dfeat, lddf, dflow, N, H, W, C, inv_w, inv_h,
                                                                            div_mode);
    return launch_status();
  }
  NV_DISPATCH_DTYPE(dtype, E, (warp_bwd_kernel<E, float, 0><<<blocks, 256, 0, as_stream(stream)>>>(
                                  (const E*)feat, ldf, flow, (const E*)dout, lddo, (float*)dfeat, lddf, dflow, N, H, W, C,
                                  inv_w, inv_h, div_mode)));
  return launch_status();
}

NV_API int nervecl_warp_bwd(const void* feat, int64_t ldf, const float* flow, const void* dout, int64_t lddo,
                            float* dfeat, int64_t lddf, float* dflow, int dtype, int N, int H, int W, int C,
                            int div_mode, nervecl_stream_t stream) {
  return warp_bwd_launch(feat, ldf, flow, dout, lddo, dfeat, lddf, NERVECL_F32, dflow, dtype, N, H, W, C, div_mode, stream);
}

NV_API int nervecl_warp_bwd_lp(const void* feat, int64_t ldf, const float* flow, const void* dout, int64_t lddo,
                               void* dfeat, int64_t lddf, float* dflow, int dtype, int N, int H, int W, int C,
                               int div_mode, nervecl_stream_t stream) {
  return warp_bwd_launch(feat, ldf, flow, dout, lddo, dfeat, lddf, dtype, dflow, dtype, N, H, W, C, div_mode, stream);
}
