// Bandwidth-tuned feature-extractor kernels (C % 8 == 0, C/8 a power of two <= 32):
//   * depthwise 3x3 forward / data-gradient and weight-gradient: a "walker" (the C/8 threads that cover one
//     pixel's channels with 128-bit accesses) slides a 3x3 register window along a 32-pixel row segment, so
//     each output pixel costs 3 new vector loads instead of 9 and the 72 filter taps of the thread's 8
//     channels live in registers for the whole kernel (the naive kernel re-loaded them per pixel and was
//     LSU-issue-bound at ~7x the HBM time);
//   * BatchNorm+ReLU forward and backward-apply with every per-channel coefficient hoisted into registers.
#include "common.cuh"

using namespace nv;

namespace {

constexpr int kThreadsFe = 256;
constexpr int DW_TH = 8;         // output rows per tile

// 16-byte async copy global -> shared (zero-fills when !valid: src-size 0)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <typename T> struct Vec8Bytes { static constexpr int value = 8 * sizeof(T); };

// Stage the halo'd tile rows [y0-1, y0+TH], pixels [x0-1, x0+TWp] of image n into shared memory as
// [(TH+2)][(TWp+2)][C] (zero outside the image) with 16-byte cp.async copies.
template <typename T>
__device__ __forceinline__ void stage_tile(T* __restrict__ sm, const T* __restrict__ x, int64_t ldx, int n, int y0, int x0,
                                           int H, int W, int C, int TWp) {
  constexpr int EPC = 16 / sizeof(T);          // elements per 16-byte chunk
  const int cpp = C / EPC;                     // chunks per pixel (divides the block size: C/8 is a power of two)
  const int PWs = TWp + 2;
  const int ch = threadIdx.x % cpp;            // this thread always copies the same chunk of a pixel ...
  const int pstep = kThreadsFe / cpp;          // ... of every pstep-th pixel: no division in the loop
  int px = threadIdx.x / cpp, r = 0;
  while (px >= PWs) { px -= PWs; ++r; }
  const T* img = x + (int64_t)n * H * W * ldx + ch * EPC;
  T* dst = sm + ch * EPC;
  for (; r < DW_TH + 2;) {
    const int yy = y0 - 1 + r, xx = x0 - 1 + px;
    const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
    const T* src = ok ? img + ((int64_t)yy * W + xx) * ldx : x;
    cp_async16(dst + (size_t)(r * PWs + px) * C, src, ok);
    px += pstep;
    while (px >= PWs) { px -= PWs; ++r; }
  }
}

// y[p,c] (+)= sum_taps x[p+tap,c] * w[c][tap]   (flip: 180-degree rotated filter = data gradient)
// block = 256 threads = (C/8 channel groups) x TWp pixel columns; tile = DW_TH rows x TWp columns.
template <typename T>
__global__ void __launch_bounds__(kThreadsFe, 2)
dwconv_tile_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w, T* __restrict__ y, int64_t ldy,
                   int N, int H, int W, int C, int flip, int accumulate, int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const int cg = C >> 3;
  const int TWp = kThreadsFe / cg;
  const int g = threadIdx.x % cg, c0 = g << 3, tx = threadIdx.x / cg;
  float wt[3][3][8];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int tap = ky * 3 + kx;
        wt[ky][kx][k] = __ldg(w + (c0 + k) * 9 + (flip ? 8 - tap : tap));
      }
  const int PWs = TWp + 2;
  const int64_t ntiles = (int64_t)N * tiles_y * tiles_x;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int txi = (int)(tile % tiles_x);
    const int64_t r = tile / tiles_x;
    const int tyi = (int)(r % tiles_y), n = (int)(r / tiles_y);
    const int x0 = txi * TWp, y0 = tyi * DW_TH;
    __syncthreads();                       // previous tile's reads are done
    stage_tile(sm, x, ldx, n, y0, x0, H, W, C, TWp);
    cp_async_wait_all();
    __syncthreads();
    const int xx = x0 + tx;
    if (xx < W) {
      // walk the DW_TH + 2 input rows of this thread's column once: each row's three neighbour vectors are
      // loaded / converted once and feed the three output rows they belong to (rolling accumulators)
      float acc[3][8];
      const T* col = sm + tx * C + c0;                                   // this thread's column of the tile
      const int rstride = PWs * C;
      T* yp = y + (((int64_t)n * H + y0) * W + xx) * ldy + c0;
      const int64_t ystride = (int64_t)W * ldy;
#pragma unroll
      for (int ri = 0; ri < DW_TH + 2; ++ri) {
        float v[3][8];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          f8 t = ld8(col + ri * rstride + kx * C);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[kx][k] = t.v[k];
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int oi = ri - ky;                     // output row (tile-local) that uses this input row as tap row ky
          if (oi < 0 || oi >= DW_TH) continue;
          float* a = acc[oi % 3];
          if (ky == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = 0.f;
          }
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fmaf(v[kx][k], wt[ky][kx][k], a[k]);
        }
        const int od = ri - 2;                        // this output row is complete
        if (od >= 0 && y0 + od < H) {
          const float* a = acc[od % 3];
          f8 o;
          if (accumulate) {
            o = ld8(yp);
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] += a[k];
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = a[k];
          }
          st8(yp, o);
          yp += ystride;
        }
      }
    }
  }
}

// dw[c][tap] += sum_p dy[p,c] * x[p+tap,c]
template <typename T>
__global__ void __launch_bounds__(kThreadsFe, 2)
dwconv_wgrad_tile_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                         float* __restrict__ dw, int N, int H, int W, int C, int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const int cg = C >> 3;
  const int TWp = kThreadsFe / cg;
  const int g = threadIdx.x % cg, c0 = g << 3, tx = threadIdx.x / cg;
  float acc[3][3][8];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[ky][kx][k] = 0.f;
  const int PWs = TWp + 2;
  T* smd = sm + (size_t)(DW_TH + 2) * PWs * C;
  const int64_t ntiles = (int64_t)N * tiles_y * tiles_x;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int txi = (int)(tile % tiles_x);
    const int64_t r = tile / tiles_x;
    const int tyi = (int)(r % tiles_y), n = (int)(r / tiles_y);
    const int x0 = txi * TWp, y0 = tyi * DW_TH;
    __syncthreads();
    stage_tile(sm, x, ldx, n, y0, x0, H, W, C, TWp);
    // dy tile [DW_TH][TWp][C] behind the x tile
    {
      constexpr int EPC = 16 / sizeof(T);
      const int cpp = C / EPC;
      const int total = DW_TH * TWp * cpp;
      for (int e = threadIdx.x; e < total; e += kThreadsFe) {
        const int ch = e % cpp;
        const int pp = e / cpp;
        const int px = pp % TWp, rr = pp / TWp;
        const int yy = y0 + rr, xq = x0 + px;
        const bool ok = yy < H && xq < W;
        const T* src = ok ? dy + (((int64_t)n * H + yy) * W + xq) * lddy + ch * EPC : dy;
        cp_async16(smd + (size_t)pp * C + ch * EPC, src, ok);
      }
    }
    cp_async_wait_all();
    __syncthreads();
    {
      // input row ri pairs with the dy rows ri - ky (tap row ky): keep the three live dy vectors in registers
      float d[3][8];
#pragma unroll
      for (int ri = 0; ri < DW_TH + 2; ++ri) {
        if (ri < DW_TH) {
          const f8 t = ld8(smd + (ri * TWp + tx) * C + c0);
#pragma unroll
          for (int k = 0; k < 8; ++k) d[ri % 3][k] = t.v[k];
        }
        float v[3][8];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          f8 t = ld8(sm + ((ri * PWs + tx + kx) * C + c0));
#pragma unroll
          for (int k = 0; k < 8; ++k) v[kx][k] = t.v[k];
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int oi = ri - ky;
          if (oi < 0 || oi >= DW_TH) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[ky][kx][k] = fmaf(d[oi % 3][k], v[kx][k], acc[ky][kx][k]);
        }
      }
    }
  }
  // block reduction: shared fp32 atomics (one address per (channel, tap)), then one global atomic each
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_raw);   // [C*9]
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&red[(c0 + k) * 9 + ky * 3 + kx], acc[ky][kx][k]);
  __syncthreads();
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) atomicAdd(dw + i, red[i]);
}

// ---------------------------------------------------------------------------------------
// BatchNorm + ReLU, 8 channels per thread, coefficients in registers.  grid = (chunks, groups).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreadsFe)
bn_relu_fwd_fast_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ stat, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const T* __restrict__ res, int64_t ldres, T* __restrict__ y,
                        int64_t ldy, int C, int64_t npix) {
  const int cg = C >> 3;
  const int g = threadIdx.x % cg, c0 = g << 3;
  const int grp = blockIdx.y;
  const int64_t base = (int64_t)grp * npix;
  float mean[8], is[8], ga[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + k;
    mean[k] = stat[((int64_t)grp * C + c) * 2];
    is[k] = stat[((int64_t)grp * C + c) * 2 + 1];
    ga[k] = gamma[c];
    be[k] = beta[c];
  }
  const int64_t stride = (int64_t)gridDim.x * kThreadsFe / cg;
#pragma unroll 4
  for (int64_t p = ((int64_t)blockIdx.x * kThreadsFe + threadIdx.x) / cg; p < npix; p += stride) {
    f8 v = ld8(x + (base + p) * ldx + c0);
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = fmaxf(fmaf((v.v[k] - mean[k]) * is[k], ga[k], be[k]), 0.f);   // == bn_value
    if (res) {
      f8 r = ld8(res + (base + p) * ldres + c0);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] += r.v[k];
    }
    st8(y + (base + p) * ldy + c0, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreadsFe)
bn_bwd_apply_fast_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                         const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ beta,
                         const double* __restrict__ bsums, T* __restrict__ dx, int64_t lddx, float* __restrict__ dgamma,
                         float* __restrict__ dbeta, int C, int64_t npix, int groups, int training) {
  const int cg = C >> 3;
  const int g = threadIdx.x % cg, c0 = g << 3;
  const int grp = blockIdx.y;
  const int64_t base = (int64_t)grp * npix;
  const float inv_n = 1.f / (float)npix;
  float mean[8], is[8], ga[8], be[8], k0[8], k1[8], k2[8];   // dx = k0*g - k1 - xhat*k2
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + k;
    mean[k] = stat[((int64_t)grp * C + c) * 2];
    is[k] = stat[((int64_t)grp * C + c) * 2 + 1];
    ga[k] = gamma[c];
    be[k] = beta[c];
    k0[k] = ga[k] * is[k];
    if (training) {
      const float sg = (float)bsums[((int64_t)grp * C + c) * 2], sq = (float)bsums[((int64_t)grp * C + c) * 2 + 1];
      k1[k] = k0[k] * sg * inv_n;
      k2[k] = k0[k] * sq * inv_n;
    } else {
      k1[k] = 0.f;
      k2[k] = 0.f;
    }
  }
  const int64_t stride = (int64_t)gridDim.x * kThreadsFe / cg;
  for (int64_t p = ((int64_t)blockIdx.x * kThreadsFe + threadIdx.x) / cg; p < npix; p += stride) {
    f8 v = ld8(x + (base + p) * ldx + c0);
    f8 d = ld8(dy + (base + p) * lddy + c0);
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (v.v[k] - mean[k]) * is[k];
      const float gk = fmaf(xh, ga[k], be[k]) > 0.f ? d.v[k] : 0.f;
      o.v[k] = k0[k] * gk - k1[k] - xh * k2[k];
    }
    st8(dx + (base + p) * lddx + c0, o);
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double sg = 0, sq = 0;
      for (int gg = 0; gg < groups; ++gg) {
        sg += bsums[((int64_t)gg * C + c) * 2];
        sq += bsums[((int64_t)gg * C + c) * 2 + 1];
      }
      if (dbeta) dbeta[c] += (float)sg;
      if (dgamma) dgamma[c] += (float)sq;
    }
  }
}

// per-(group, channel) sums in float64: sums[g][c][0..1] += (sum a, sum b) where (a, b) = (x, x^2) for the
// statistics pass and (g, g*xhat) with g = dy * [bn(x) > 0] for the backward reduction.  grid = (chunks, groups);
// a thread owns 8 channels of every (256/cg)-th pixel; fp32 partials are flushed to fp64 every 32 pixels.
template <typename T, bool BWD>
__global__ void __launch_bounds__(kThreadsFe, 2)
bn_sums_fast_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                    const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ beta, int C,
                    int64_t npix, double* __restrict__ sums) {
  const int cg = C >> 3;
  const int g = threadIdx.x % cg, c0 = g << 3;
  const int grp = blockIdx.y;
  const int64_t base = (int64_t)grp * npix;
  float mean[8], is[8], ga[8], be[8];
  if (BWD) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + k;
      mean[k] = stat[((int64_t)grp * C + c) * 2];
      is[k] = stat[((int64_t)grp * C + c) * 2 + 1];
      ga[k] = gamma[c];
      be[k] = beta[c];
    }
  }
  double sa[8], sb[8];
  float fa[8], fb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sa[k] = sb[k] = 0.0; fa[k] = fb[k] = 0.f; }
  int cnt = 0;
  const int64_t stride = (int64_t)gridDim.x * kThreadsFe / cg;
  constexpr int U = 4;                       // independent 16-byte loads in flight per thread and operand
  for (int64_t p0 = ((int64_t)blockIdx.x * kThreadsFe + threadIdx.x) / cg; p0 < npix; p0 += U * stride) {
    f8 v[U], d[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = p0 + u * stride;
      ok[u] = p < npix;
      if (ok[u]) {
        v[u] = ld8(x + (base + p) * ldx + c0);
        if (BWD) d[u] = ld8(dy + (base + p) * lddy + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      if (BWD) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (v[u].v[k] - mean[k]) * is[k];
          const float gk = fmaf(xh, ga[k], be[k]) > 0.f ? d[u].v[k] : 0.f;
          fa[k] += gk;
          fb[k] = fmaf(gk, xh, fb[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          fa[k] += v[u].v[k];
          fb[k] = fmaf(v[u].v[k], v[u].v[k], fb[k]);
        }
      }
    }
    if (++cnt == 16) {                       // 64 pixels: flush the fp32 partials to fp64 before they lose bits
#pragma unroll
      for (int k = 0; k < 8; ++k) { sa[k] += fa[k]; sb[k] += fb[k]; fa[k] = fb[k] = 0.f; }
      cnt = 0;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { sa[k] += fa[k]; sb[k] += fb[k]; }
  extern __shared__ double dred[];  // [C][2]
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) dred[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    atomicAdd(&dred[(c0 + k) * 2 + 0], sa[k]);
    atomicAdd(&dred[(c0 + k) * 2 + 1], sb[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) atomicAdd(sums + (int64_t)grp * C * 2 + i, dred[i]);
}

inline bool cg_ok(int C) {
  if (C & 7) return false;
  const int cg = C >> 3;
  return cg <= 32 && (cg & (cg - 1)) == 0;
}

}  // namespace

namespace nv {

bool fe_fast_supported(int C, int64_t lda, int64_t ldb, const void* a, const void* b) {
  return cg_ok(C) && !(lda & 7) && !(ldb & 7) && aligned(a, 16) && aligned(b, 16);
}

int dwconv_slide(const void* x, int64_t ldx, const float* w, void* y, int64_t ldy, int dtype, int N, int H, int W, int C,
                 int flip, int accumulate, cudaStream_t s) {
  const int TWp = kThreadsFe / (C >> 3);
  const int tiles_x = (W + TWp - 1) / TWp, tiles_y = (H + DW_TH - 1) / DW_TH;
  const int64_t ntiles = (int64_t)N * tiles_x * tiles_y;
  const int blocks = (int)imax(1, imin(ntiles, kSMs * 2));
  const size_t esz = dtype == NERVECL_F32 ? 4 : 2;
  const size_t smem = (size_t)(DW_TH + 2) * (TWp + 2) * C * esz;
  cudaError_t e;
#define NV_DW_LAUNCH(E)                                                                                              \
  e = cudaFuncSetAttribute(dwconv_tile_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
  if (e != cudaSuccess) return (int)e;                                                                              \
  dwconv_tile_kernel<E><<<blocks, kThreadsFe, smem, s>>>((const E*)x, ldx, w, (E*)y, ldy, N, H, W, C, flip,         \
                                                        accumulate, tiles_x, tiles_y)
  if (dtype == NERVECL_F32) { NV_DW_LAUNCH(float); } else { NV_DW_LAUNCH(bf16); }
#undef NV_DW_LAUNCH
  return launch_status();
}

int dwconv_wgrad_slide(const void* x, int64_t ldx, const void* dy, int64_t lddy, int dtype, float* dw, int N, int H, int W,
                       int C, cudaStream_t s) {
  const int TWp = kThreadsFe / (C >> 3);
  const int tiles_x = (W + TWp - 1) / TWp, tiles_y = (H + DW_TH - 1) / DW_TH;
  const int64_t ntiles = (int64_t)N * tiles_x * tiles_y;
  const int blocks = (int)imax(1, imin(ntiles, kSMs * 2));
  const size_t esz = dtype == NERVECL_F32 ? 4 : 2;
  const size_t smem = imax((int64_t)(((size_t)(DW_TH + 2) * (TWp + 2) + (size_t)DW_TH * TWp) * C * esz), (int64_t)C * 9 * 4);
  cudaError_t e;
#define NV_DW_LAUNCH(E)                                                                                              \
  e = cudaFuncSetAttribute(dwconv_wgrad_tile_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
  if (e != cudaSuccess) return (int)e;                                                                              \
  dwconv_wgrad_tile_kernel<E><<<blocks, kThreadsFe, smem, s>>>((const E*)x, ldx, (const E*)dy, lddy, dw, N, H, W, C, \
                                                              tiles_x, tiles_y)
  if (dtype == NERVECL_F32) { NV_DW_LAUNCH(float); } else { NV_DW_LAUNCH(bf16); }
#undef NV_DW_LAUNCH
  return launch_status();
}

int bn_sums_fast(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* stat, const float* gamma,
                 const float* beta, int dtype, int C, int64_t npix, int groups, double* sums, cudaStream_t s) {
  const int chunks = (int)imax(1, imin(cdiv(npix * (C >> 3), kThreadsFe * 16), (kSMs * 2 * 2) / groups + 1));
  dim3 grid(chunks, groups);
  const size_t smem = (size_t)C * 2 * sizeof(double);
  if (dy) {
    NV_DISPATCH_DTYPE(dtype, E, (bn_sums_fast_kernel<E, true><<<grid, kThreadsFe, smem, s>>>(
                                    (const E*)x, ldx, (const E*)dy, lddy, stat, gamma, beta, C, npix, sums)));
  } else {
    NV_DISPATCH_DTYPE(dtype, E, (bn_sums_fast_kernel<E, false><<<grid, kThreadsFe, smem, s>>>(
                                    (const E*)x, ldx, nullptr, 0, nullptr, nullptr, nullptr, C, npix, sums)));
  }
  return launch_status();
}

int bn_relu_fwd_fast(const void* x, int64_t ldx, const float* stat, const float* gamma, const float* beta, const void* res,
                     int64_t ldres, void* y, int64_t ldy, int dtype, int C, int64_t npix, int groups, cudaStream_t s) {
  const int chunks = (int)imax(1, imin(cdiv(npix * (C >> 3), kThreadsFe * 4), (kSMs * 8) / groups + 1));
  dim3 grid(chunks, groups);
  NV_DISPATCH_DTYPE(dtype, E, (bn_relu_fwd_fast_kernel<E><<<grid, kThreadsFe, 0, s>>>((const E*)x, ldx, stat, gamma, beta,
                                                                                      (const E*)res, ldres, (E*)y, ldy, C,
                                                                                      npix)));
  return launch_status();
}

int bn_bwd_apply_fast(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* stat, const float* gamma,
                      const float* beta, const double* bsums, void* dx, int64_t lddx, float* dgamma, float* dbeta, int dtype,
                      int C, int64_t npix, int groups, int training, cudaStream_t s) {
  const int chunks = (int)imax(1, imin(cdiv(npix * (C >> 3), kThreadsFe * 4), (kSMs * 8) / groups + 1));
  dim3 grid(chunks, groups);
  NV_DISPATCH_DTYPE(dtype, E, (bn_bwd_apply_fast_kernel<E><<<grid, kThreadsFe, 0, s>>>(
                                  (const E*)x, ldx, (const E*)dy, lddy, stat, gamma, beta, bsums, (E*)dx, lddx, dgamma, dbeta,
                                  C, npix, groups, training)));
  return launch_status();
}

}  // namespace nv
