// Feature-extractor body kernels: depthwise 3x3 (fwd / dgrad / wgrad) and grouped BatchNorm+ReLU.
// All are HBM-bound: 128-bit (bf16) / 2x128-bit (f32) channel-vector accesses, fp32 math.
#include "common.cuh"
#include <cstdlib>

using namespace nv;

namespace nv {   // fe_fast.cu
bool fe_fast_supported(int C, int64_t lda, int64_t ldb, const void* a, const void* b);
bool dw_tiled_supported(int C, int64_t lda, int64_t ldb, const void* a, const void* b);
int dwconv_slide(const void* x, int64_t ldx, const float* w, void* y, int64_t ldy, int dtype, int N, int H, int W, int C,
                 int flip, int accumulate, cudaStream_t s, const void* add = nullptr, int64_t ldadd = 0,
                 const void* mask = nullptr, int64_t ldmask = 0);
int dwconv_wgrad_slide(const void* x, int64_t ldx, const void* dy, int64_t lddy, int dtype, float* dw, int N, int H, int W,
                       int C, cudaStream_t s);
bool bn_sums_fast_supported(int C, int64_t lda, int64_t ldb);
int bn_sums_fast(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* stat, const float* gamma,
                 const float* beta, int dtype, int C, int64_t npix, int groups, double* sums, cudaStream_t s);
int bn_relu_fwd_fast(const void* x, int64_t ldx, const float* stat, const float* gamma, const float* beta, const void* res,
                     int64_t ldres, void* y, int64_t ldy, int dtype, int C, int64_t npix, int groups, cudaStream_t s);
int bn_bwd_apply_fast(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* stat, const float* gamma,
                      const float* beta, const double* bsums, void* dx, int64_t lddx, float* dgamma, float* dbeta, int dtype,
                      int C, int64_t npix, int groups, int training, cudaStream_t s);
}

namespace {

// ---------------------------------------------------------------------------------------
// depthwise 3x3: one thread = one pixel x 8 channels; the 3x3 neighbourhood is re-read through
// L1/L2 (each line is touched by 9 neighbouring threads of the same or adjacent warps).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
dwconv_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w, T* __restrict__ y,
              int64_t ldy, int N, int H, int W, int C, int flip, int accumulate) {
  const int cg = C >> 3;  // channel groups of 8
  const int64_t total = (int64_t)N * H * W * cg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int g = (int)(i % cg);
    int64_t p = i / cg;
    int xx = (int)(p % W);
    int yy = (int)((p / W) % H);
    int c0 = g << 3;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      int sy = yy + ky - 1;
      if (sy < 0 || sy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        int sx = xx + kx - 1;
        if (sx < 0 || sx >= W) continue;
        int tap = ky * 3 + kx;
        if (flip) tap = 8 - tap;
        f8 v = ld8(x + (p + (int64_t)(ky - 1) * W + (kx - 1)) * ldx + c0);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(v.v[k], __ldg(w + (c0 + k) * 9 + tap), acc[k]);
      }
    }
    f8 o;
    if (accumulate) {
      o = ld8(y + p * ldy + c0);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] += acc[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = acc[k];
    }
    st8(y + p * ldy + c0, o);
  }
}

// dw[c][tap] += sum_p dy[p,c] * x[p+tap,c].  blockDim = (C/4) x L ; thread = 4 channels.
template <typename T>
__global__ void __launch_bounds__(256)
dwconv_wgrad_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t ldy,
                    float* __restrict__ dw, int N, int H, int W, int C) {
  const int cg = C >> 2;
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int c0 = g << 2;
  const int64_t npix = (int64_t)N * H * W;
  float acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[t][k] = 0.f;
  if (lane < lanes) {
    for (int64_t p = (int64_t)blockIdx.x * lanes + lane; p < npix; p += (int64_t)gridDim.x * lanes) {
      int xx = (int)(p % W);
      int yy = (int)((p / W) % H);
      f4 gch = ld4(dy + p * ldy + c0);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        int sy = yy + ky - 1;
        if (sy < 0 || sy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          int sx = xx + kx - 1;
          if (sx < 0 || sx >= W) continue;
          f4 v = ld4(x + (p + (int64_t)(ky - 1) * W + (kx - 1)) * ldx + c0);
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[ky * 3 + kx][k] = fmaf(gch.v[k], v.v[k], acc[ky * 3 + kx][k]);
        }
      }
    }
  }
  extern __shared__ float red[];  // [C*9]
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  if (lane < lanes) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(&red[(c0 + k) * 9 + t], acc[t][k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) atomicAdd(dw + i, red[i]);
}

// ---------------------------------------------------------------------------------------
// BatchNorm
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float bn_value(float x, float mean, float invstd, float gamma, float beta) {
  return fmaf((x - mean) * invstd, gamma, beta);
}

// grid = (chunks, groups); block = 256 = (C/4) x lanes
template <typename T>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const T* __restrict__ x, int64_t ldx, int C, int64_t npix, double* __restrict__ sums) {
  const int cg = C >> 2;
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int grp = blockIdx.y;
  const T* xb = x + (int64_t)grp * npix * ldx;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (lane < lanes) {
    float fs[4] = {0, 0, 0, 0}, fq[4] = {0, 0, 0, 0};
    int cnt = 0;
    for (int64_t p = (int64_t)blockIdx.x * lanes + lane; p < npix; p += (int64_t)gridDim.x * lanes) {
      f4 v = ld4(xb + p * ldx + (g << 2));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        fs[k] += v.v[k];
        fq[k] = fmaf(v.v[k], v.v[k], fq[k]);
      }
      if (++cnt == 64) {  // flush fp32 partials to fp64 before they lose bits
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          s[k] += fs[k]; q[k] += fq[k]; fs[k] = 0.f; fq[k] = 0.f;
        }
        cnt = 0;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[k] += fs[k]; q[k] += fq[k]; }
  }
  extern __shared__ double dred[];  // [C][2]
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) dred[i] = 0.0;
  __syncthreads();
  if (lane < lanes) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&dred[((g << 2) + k) * 2 + 0], s[k]);
      atomicAdd(&dred[((g << 2) + k) * 2 + 1], q[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) atomicAdd(sums + (int64_t)grp * C * 2 + i, dred[i]);
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, float* __restrict__ stat,
                                   float* __restrict__ rmean, float* __restrict__ rvar,
                                   int64_t* __restrict__ nbt, int C, int64_t npix, int groups, float momentum,
                                   float eps, int training) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    if (training) {
      float rm = rmean[c], rv = rvar[c];
      for (int g = 0; g < groups; ++g) {
        double s = sums[((int64_t)g * C + c) * 2], q = sums[((int64_t)g * C + c) * 2 + 1];
        double mean = s / (double)npix;
        double var = q / (double)npix - mean * mean;
        if (var < 0) var = 0;
        stat[((int64_t)g * C + c) * 2] = (float)mean;
        stat[((int64_t)g * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
        double unbiased = npix > 1 ? var * (double)npix / (double)(npix - 1) : var;
        rm = (1.f - momentum) * rm + momentum * (float)mean;
        rv = (1.f - momentum) * rv + momentum * (float)unbiased;
      }
      rmean[c] = rm;
      rvar[c] = rv;
    } else {
      float m = rmean[c], is = 1.f / sqrtf(rvar[c] + eps);
      for (int g = 0; g < groups; ++g) {
        stat[((int64_t)g * C + c) * 2] = m;
        stat[((int64_t)g * C + c) * 2 + 1] = is;
      }
    }
  }
  if (training && nbt && c == 0) *nbt += groups;
}

// grid = (chunks, groups)
template <typename T>
__global__ void __launch_bounds__(256)
bn_relu_fwd_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ stat,
                   const float* __restrict__ gamma, const float* __restrict__ beta, const T* __restrict__ res,
                   int64_t ldres, T* __restrict__ y, int64_t ldy, int C, int64_t npix) {
  const int cg = C >> 2;
  const int grp = blockIdx.y;
  const int64_t base = (int64_t)grp * npix;
  const int64_t total = npix * cg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c0 = (int)(i % cg) << 2;
    int64_t p = base + i / cg;
    f4 v = ld4(x + p * ldx + c0);
    f4 o;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int c = c0 + k;
      float m = __ldg(stat + ((int64_t)grp * C + c) * 2), is = __ldg(stat + ((int64_t)grp * C + c) * 2 + 1);
      o.v[k] = fmaxf(bn_value(v.v[k], m, is, __ldg(gamma + c), __ldg(beta + c)), 0.f);
    }
    if (res) {
      f4 r = ld4(res + p * ldres + c0);
#pragma unroll
      for (int k = 0; k < 4; ++k) o.v[k] += r.v[k];
    }
    st4(y + p * ldy + c0, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                     const float* __restrict__ stat, const float* __restrict__ gamma,
                     const float* __restrict__ beta, int C, int64_t npix, double* __restrict__ bsums) {
  const int cg = C >> 2;
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int grp = blockIdx.y;
  const int64_t base = (int64_t)grp * npix;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (lane < lanes) {
    float mean[4], is[4], ga[4], be[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int c = (g << 2) + k;
      mean[k] = stat[((int64_t)grp * C + c) * 2];
      is[k] = stat[((int64_t)grp * C + c) * 2 + 1];
      ga[k] = gamma[c];
      be[k] = beta[c];
    }
    float fs[4] = {0, 0, 0, 0}, fq[4] = {0, 0, 0, 0};
    int cnt = 0;
    for (int64_t p = (int64_t)blockIdx.x * lanes + lane; p < npix; p += (int64_t)gridDim.x * lanes) {
      f4 v = ld4(x + (base + p) * ldx + (g << 2));
      f4 d = ld4(dy + (base + p) * lddy + (g << 2));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float gk = bn_value(v.v[k], mean[k], is[k], ga[k], be[k]) > 0.f ? d.v[k] : 0.f;
        fs[k] += gk;
        fq[k] = fmaf(gk, (v.v[k] - mean[k]) * is[k], fq[k]);
      }
      if (++cnt == 64) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { s[k] += fs[k]; q[k] += fq[k]; fs[k] = 0.f; fq[k] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[k] += fs[k]; q[k] += fq[k]; }
  }
  extern __shared__ double dred[];
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) dred[i] = 0.0;
  __syncthreads();
  if (lane < lanes) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&dred[((g << 2) + k) * 2 + 0], s[k]);
      atomicAdd(&dred[((g << 2) + k) * 2 + 1], q[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) atomicAdd(bsums + (int64_t)grp * C * 2 + i, dred[i]);
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                    const float* __restrict__ stat, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const double* __restrict__ bsums, T* __restrict__ dx,
                    int64_t lddx, float* __restrict__ dgamma, float* __restrict__ dbeta, int C, int64_t npix,
                    int groups, int training) {
  const int cg = C >> 2;
  const int grp = blockIdx.y;
  const int64_t base = (int64_t)grp * npix;
  const int64_t total = npix * cg;
  const float inv_n = 1.f / (float)npix;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c0 = (int)(i % cg) << 2;
    int64_t p = base + i / cg;
    f4 v = ld4(x + p * ldx + c0);
    f4 d = ld4(dy + p * lddy + c0);
    f4 o;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int c = c0 + k;
      float m = __ldg(stat + ((int64_t)grp * C + c) * 2), is = __ldg(stat + ((int64_t)grp * C + c) * 2 + 1);
      float ga = __ldg(gamma + c);
      float gk = bn_value(v.v[k], m, is, ga, __ldg(beta + c)) > 0.f ? d.v[k] : 0.f;
      float xh = (v.v[k] - m) * is;
      if (training) {
        float sg = (float)bsums[((int64_t)grp * C + c) * 2], sq = (float)bsums[((int64_t)grp * C + c) * 2 + 1];
        o.v[k] = ga * is * (gk - sg * inv_n - xh * sq * inv_n);
      } else {
        o.v[k] = ga * is * gk;
      }
    }
    st4(dx + p * lddx + c0, o);
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double sg = 0, sq = 0;
      for (int g = 0; g < groups; ++g) {
        sg += bsums[((int64_t)g * C + c) * 2];
        sq += bsums[((int64_t)g * C + c) * 2 + 1];
      }
      if (dbeta) dbeta[c] += (float)sg;
      if (dgamma) dgamma[c] += (float)sq;
    }
  }
}

inline int ew_blocks(int64_t work) { return (int)imax(1, imin(cdiv(work, 256), kSMs * 16)); }

}  // namespace

NV_API int nervecl_dwconv3x3_fwd(const void* x, int64_t ldx, const float* w, void* y, int64_t ldy, int dtype,
                                 int N, int H, int W, int C, int flip, int accumulate, nervecl_stream_t stream) {
  if (!x || !w || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0) return NERVECL_EINVAL;
  if ((C & 7) || (ldx & 7) || (ldy & 7) || !aligned(x, 16) || !aligned(y, 16)) return NERVECL_EALIGN;
  if (dtype != NERVECL_F32 && dtype != NERVECL_BF16) return NERVECL_EDTYPE;
  if (dw_tiled_supported(C, ldx, ldy, x, y))
    return dwconv_slide(x, ldx, w, y, ldy, dtype, N, H, W, C, flip, accumulate, as_stream(stream));
  int64_t total = (int64_t)N * H * W * (C >> 3);
  NV_DISPATCH_DTYPE(dtype, E, (dwconv_kernel<E><<<ew_blocks(total), 256, 0, as_stream(stream)>>>(
                                  (const E*)x, ldx, w, (E*)y, ldy, N, H, W, C, flip, accumulate)));
  return launch_status();
}

NV_API int nervecl_dwconv3x3_fwd_masked(const void* x, int64_t ldx, const float* w, const void* add, int64_t ldadd,
                                        const void* mask, int64_t ldmask, void* y, int64_t ldy, int dtype, int N, int H,
                                        int W, int C, int flip, nervecl_stream_t stream) {
  if (!x || !w || !y || !add || !mask || N <= 0 || H <= 0 || W <= 0 || C <= 0) return NERVECL_EINVAL;
  if ((C & 7) || (ldx & 7) || (ldy & 7) || (ldadd & 7) || (ldmask & 7) || !aligned(x, 16) || !aligned(y, 16) ||
      !aligned(add, 16) || !aligned(mask, 16))
    return NERVECL_EALIGN;
  if (dtype != NERVECL_BF16) return NERVECL_EDTYPE;
  if (!dw_tiled_supported(C, ldx, ldy, x, y)) return NERVECL_EUNSUPPORTED;
  return dwconv_slide(x, ldx, w, y, ldy, dtype, N, H, W, C, flip, 0, as_stream(stream), add, ldadd, mask, ldmask);
}

NV_API int nervecl_dwconv3x3_wgrad(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype,
                                   float* dw, int N, int H, int W, int C, nervecl_stream_t stream) {
  if (!x || !dy || !dw || N <= 0 || H <= 0 || W <= 0 || C <= 0) return NERVECL_EINVAL;
  if ((C & 3) || (ldx & 3) || (ldy & 3) || C > 1024) return NERVECL_EALIGN;
  if (dtype != NERVECL_F32 && dtype != NERVECL_BF16) return NERVECL_EDTYPE;
  if (dw_tiled_supported(C, ldx, ldy, x, dy))
    return dwconv_wgrad_slide(x, ldx, dy, ldy, dtype, dw, N, H, W, C, as_stream(stream));
  int64_t npix = (int64_t)N * H * W;
  int lanes = 256 / (C >> 2);
  if (lanes < 1) return NERVECL_EUNSUPPORTED;
  int blocks = (int)imax(1, imin(cdiv(npix, lanes * 8), kSMs * 4));
  size_t smem = (size_t)C * 9 * sizeof(float);
  NV_DISPATCH_DTYPE(dtype, E, (dwconv_wgrad_kernel<E><<<blocks, 256, smem, as_stream(stream)>>>(
                                  (const E*)x, ldx, (const E*)dy, ldy, dw, N, H, W, C)));
  return launch_status();
}

NV_API int nervecl_bn_stats(const void* x, int64_t ldx, int dtype, int C, int64_t npix, int groups,
                            double* sums, nervecl_stream_t stream) {
  if (!x || !sums || C <= 0 || npix <= 0 || groups <= 0) return NERVECL_EINVAL;
  if ((C & 3) || (ldx & 3) || C > 1024) return NERVECL_EALIGN;
  if (dtype != NERVECL_F32 && dtype != NERVECL_BF16) return NERVECL_EDTYPE;
  if (!nv::tune_env("NERVECL_BN_SUMS_PLAIN") && bn_sums_fast_supported(C, ldx, ldx))
    return bn_sums_fast(x, ldx, nullptr, 0, nullptr, nullptr, nullptr, dtype, C, npix, groups, sums, as_stream(stream));
  int lanes = 256 / (C >> 2);
  int chunks = (int)imax(1, imin(cdiv(npix, lanes * 16), (kSMs * 8) / groups));
  dim3 grid(chunks, groups);
  size_t smem = (size_t)C * 2 * sizeof(double);
  NV_DISPATCH_DTYPE(dtype, E, (bn_stats_kernel<E><<<grid, 256, smem, as_stream(stream)>>>(
                                  (const E*)x, ldx, C, npix, sums)));
  return launch_status();
}

NV_API int nervecl_bn_finalize(const double* sums, float* stat, float* running_mean, float* running_var,
                               int64_t* num_batches_tracked, int C, int64_t npix, int groups, float momentum,
                               float eps, int training, nervecl_stream_t stream) {
  if (!stat || !running_mean || !running_var || C <= 0 || groups <= 0) return NERVECL_EINVAL;
  if (training && !sums) return NERVECL_EINVAL;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(
      sums, stat, running_mean, running_var, num_batches_tracked, C, npix, groups, momentum, eps, training);
  return launch_status();
}

NV_API int nervecl_bn_relu_fwd(const void* x, int64_t ldx, const float* stat, const float* gamma,
                               const float* beta, const void* res, int64_t ldres, void* y, int64_t ldy,
                               int dtype, int C, int64_t npix, int groups, nervecl_stream_t stream) {
  if (!x || !stat || !gamma || !beta || !y || C <= 0 || npix <= 0 || groups <= 0) return NERVECL_EINVAL;
  if ((C & 3) || (ldx & 3) || (ldy & 3) || (res && (ldres & 3))) return NERVECL_EALIGN;
  if (dtype != NERVECL_F32 && dtype != NERVECL_BF16) return NERVECL_EDTYPE;
  if (fe_fast_supported(C, ldx, ldy, x, y) && (!res || (!(ldres & 7) && aligned(res, 16))))
    return bn_relu_fwd_fast(x, ldx, stat, gamma, beta, res, ldres, y, ldy, dtype, C, npix, groups, as_stream(stream));
  dim3 grid(ew_blocks(npix * (C >> 2)), groups);
  NV_DISPATCH_DTYPE(dtype, E, (bn_relu_fwd_kernel<E><<<grid, 256, 0, as_stream(stream)>>>(
                                  (const E*)x, ldx, stat, gamma, beta, (const E*)res, ldres, (E*)y, ldy, C, npix)));
  return launch_status();
}

NV_API int nervecl_bn_relu_bwd_reduce(const void* x, int64_t ldx, const void* dy, int64_t lddy,
                                      const float* stat, const float* gamma, const float* beta, int dtype,
                                      int C, int64_t npix, int groups, double* bsums, nervecl_stream_t stream) {
  if (!x || !dy || !stat || !gamma || !beta || !bsums || C <= 0 || npix <= 0 || groups <= 0) return NERVECL_EINVAL;
  if ((C & 3) || (ldx & 3) || (lddy & 3) || C > 1024) return NERVECL_EALIGN;
  if (dtype != NERVECL_F32 && dtype != NERVECL_BF16) return NERVECL_EDTYPE;
  if (!nv::tune_env("NERVECL_BN_SUMS_PLAIN") && bn_sums_fast_supported(C, ldx, lddy))
    return bn_sums_fast(x, ldx, dy, lddy, stat, gamma, beta, dtype, C, npix, groups, bsums, as_stream(stream));
  int lanes = 256 / (C >> 2);
  int chunks = (int)imax(1, imin(cdiv(npix, lanes * 16), (kSMs * 8) / groups));
  dim3 grid(chunks, groups);
  size_t smem = (size_t)C * 2 * sizeof(double);
  NV_DISPATCH_DTYPE(dtype, E, (bn_bwd_reduce_kernel<E><<<grid, 256, smem, as_stream(stream)>>>(
                                  (const E*)x, ldx, (const E*)dy, lddy, stat, gamma, beta, C, npix, bsums)));
  return launch_status();
}

NV_API int nervecl_bn_relu_bwd_apply(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* stat,
                                     const float* gamma, const float* beta, const double* bsums, void* dx,
                                     int64_t lddx, float* dgamma, float* dbeta, int dtype, int C, int64_t npix,
                                     int groups, int training, nervecl_stream_t stream) {
  if (!x || !dy || !stat || !gamma || !beta || !bsums || !dx || C <= 0 || npix <= 0 || groups <= 0)
    return NERVECL_EINVAL;
  if ((C & 3) || (ldx & 3) || (lddy & 3) || (lddx & 3)) return NERVECL_EALIGN;
  if (dtype != NERVECL_F32 && dtype != NERVECL_BF16) return NERVECL_EDTYPE;
  if (fe_fast_supported(C, ldx, lddy, x, dy) && !(lddx & 7) && aligned(dx, 16))
    return bn_bwd_apply_fast(x, ldx, dy, lddy, stat, gamma, beta, bsums, dx, lddx, dgamma, dbeta, dtype, C, npix, groups,
                             training, as_stream(stream));
  dim3 grid(ew_blocks(npix * (C >> 2)), groups);
  NV_DISPATCH_DTYPE(dtype, E, (bn_bwd_apply_kernel<E><<<grid, 256, 0, as_stream(stream)>>>(
                                  (const E*)x, ldx, (const E*)dy, lddy, stat, gamma, beta, bsums, (E*)dx, lddx,
                                  dgamma, dbeta, C, npix, groups, training)));
  return launch_status();
}
