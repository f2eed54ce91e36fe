// Layout conversion, weight packing and small elementwise gradient-routing kernels.
#include "common.cuh"

using namespace nv;

NV_API int nervecl_abi_version(void) { return 7; }

NV_API const char* nervecl_error_string(int code) {
  switch (code) {
    case NERVECL_OK: return "ok";
    case NERVECL_EINVAL: return "invalid argument (shape or null pointer)";
    case NERVECL_EALIGN: return "misaligned pointer or pitch";
    case NERVECL_EDTYPE: return "unsupported dtype";
    case NERVECL_EUNSUPPORTED: return "shape not supported by the requested engine";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown nervecl error";
  }
}

// ---------------------------------------------------------------------------------------
// (B,T,C,H,W) strided fp32 -> [T][B][H][W][C]
// ---------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void pack_frames_kernel(const float* __restrict__ src, int64_t sB, int64_t sT, int64_t sC,
                                   int64_t sH, T* __restrict__ dst, int64_t ldd, int B, int Tn, int C, int H,
                                   int W) {
  int64_t total = (int64_t)Tn * B * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % W);
    int64_t r = i / W;
    int y = (int)(r % H);
    r /= H;
    int b = (int)(r % B);
    int t = (int)(r / B);
    const float* s = src + b * sB + t * sT + y * sH + x;
    T* d = dst + i * ldd;
    if (VEC) {                                              // ldd % 8 == 0, 16-byte aligned rows: vector stores
      for (int c8 = 0; c8 < (int)ldd; c8 += 8) {
        f8 v;
#pragma unroll
        for (int k = 0; k < 8; ++k) v.v[k] = c8 + k < C ? __ldg(s + (c8 + k) * sC) : 0.f;
        st8(d + c8, v);
      }
    } else {
      for (int c = 0; c < C; ++c) stf(d + c, __ldg(s + c * sC));
      for (int c = C; c < (int)ldd; ++c) stf(d + c, 0.f);   // zero the channel padding
    }
  }
}

NV_API int nervecl_pack_frames(const float* src, int64_t sB, int64_t sT, int64_t sC, int64_t sH,
                               void* dst, int64_t ldd, int dtype, int B, int T, int C, int H, int W,
                               nervecl_stream_t stream) {
  if (!src || !dst || B <= 0 || T <= 0 || C <= 0 || H <= 0 || W <= 0 || ldd < C) return NERVECL_EINVAL;
  int64_t total = (int64_t)T * B * H * W;
  int blocks = (int)imin(cdiv(total, 256), kSMs * 16);
  if (!(ldd & 7) && aligned(dst, 16)) {
    NV_DISPATCH_DTYPE(dtype, E, (pack_frames_kernel<E, true><<<blocks, 256, 0, as_stream(stream)>>>(
                                    src, sB, sT, sC, sH, (E*)dst, ldd, B, T, C, H, W)));
  } else {
    NV_DISPATCH_DTYPE(dtype, E, (pack_frames_kernel<E, false><<<blocks, 256, 0, as_stream(stream)>>>(
                                    src, sB, sT, sC, sH, (E*)dst, ldd, B, T, C, H, W)));
  }
  return launch_status();
}

// (B,T,C,H,W) strided fp32 -> [T][B][H][W][c*9 + tap] (3x3 unfold, zero padding); ldd % 8 == 0
template <typename T, bool W16>
__global__ void pack_frames_unfold3_kernel(const float* __restrict__ src, int64_t sB, int64_t sT, int64_t sC,
                                           int64_t sH, T* __restrict__ dst, int64_t ldd, int B, int Tn, int C,
                                           int H, int W) {
  int64_t total = (int64_t)Tn * B * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % W);
    int64_t r = i / W;
    int y = (int)(r % H);
    r /= H;
    int b = (int)(r % B);
    int t = (int)(r / B);
    const float* s = src + b * sB + t * sT;
    T* d = dst + i * ldd;
    if (W16) {                                              // 32-byte aligned rows: full-sector stores
      for (int c16 = 0; c16 < (int)ldd; c16 += 16) {
        f16v v;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int col = c16 + k, c = col / 9, tap = col - c * 9;
          const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
          v.v[k] = (c < C && yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(s + c * sC + yy * sH + xx) : 0.f;
        }
        st16(d + c16, v);
      }
      continue;
    }
    for (int c8 = 0; c8 < (int)ldd; c8 += 8) {
      f8 v;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int col = c8 + k, c = col / 9, tap = col - c * 9;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        v.v[k] = (c < C && yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(s + c * sC + yy * sH + xx) : 0.f;
      }
      st8(d + c8, v);
    }
  }
}

// The head conv's shape (C == 3, 32-channel rows, 32-byte aligned): every column index is a compile-time constant and the
// pixel coordinates are 32-bit (the generic kernel divides col / 9 and tap / 3 at run time for all 32 columns: 0.52 ms at
// the cfg-2 shape against 0.13 ms of HBM time).
template <typename T>
__global__ void __launch_bounds__(256)
pack_frames_unfold3_c3_kernel(const float* __restrict__ src, int64_t sB, int64_t sT, int64_t sC, int64_t sH,
                              T* __restrict__ dst, int B, int H, int W, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = i % W, r = i / W;
  const int y = r % H, r2 = r / H;
  const int b = r2 % B, t = r2 / B;
  const float* s = src + b * sB + t * sT + (int64_t)y * sH + x;
  float v[27];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      const bool ok = y + dy >= 0 && y + dy < H && x + dx >= 0 && x + dx < W;
      v[c * 9 + tap] = ok ? __ldg(s + c * sC + dy * sH + dx) : 0.f;
    }
  T* d = dst + (int64_t)i * 32;
#pragma unroll
  for (int c16 = 0; c16 < 32; c16 += 16) {
    f16v o;
#pragma unroll
    for (int k = 0; k < 16; ++k) o.v[k] = (c16 + k < 27) ? v[(c16 + k < 27) ? c16 + k : 0] : 0.f;
    st16(d + c16, o);
  }
}

NV_API int nervecl_pack_frames_unfold3(const float* src, int64_t sB, int64_t sT, int64_t sC, int64_t sH,
                                       void* dst, int64_t ldd, int dtype, int B, int T, int C, int H, int W,
                                       nervecl_stream_t stream) {
  if (!src || !dst || B <= 0 || T <= 0 || C <= 0 || H <= 0 || W <= 0 || ldd < 9 * C) return NERVECL_EINVAL;
  if ((ldd & 7) || !aligned(dst, 16)) return NERVECL_EALIGN;
  int64_t total = (int64_t)T * B * H * W;
  int blocks = (int)imin(cdiv(total, 256), kSMs * 16);
  if (C == 3 && ldd == 32 && aligned(dst, 32) && total < ((int64_t)1 << 31)) {
    NV_DISPATCH_DTYPE(dtype, E, (pack_frames_unfold3_c3_kernel<E><<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(
                                    src, sB, sT, sC, sH, (E*)dst, B, H, W, (int)total)));
    return launch_status();
  }
  if (!(ldd & 15) && aligned(dst, 32)) {
    NV_DISPATCH_DTYPE(dtype, E, (pack_frames_unfold3_kernel<E, true><<<blocks, 256, 0, as_stream(stream)>>>(
                                    src, sB, sT, sC, sH, (E*)dst, ldd, B, T, C, H, W)));
  } else {
    NV_DISPATCH_DTYPE(dtype, E, (pack_frames_unfold3_kernel<E, false><<<blocks, 256, 0, as_stream(stream)>>>(
                                    src, sB, sT, sC, sH, (E*)dst, ldd, B, T, C, H, W)));
  }
  return launch_status();
}

// dst[p][o*9 + tap] = src[p - tap offset][o]: gradient-side unfold of a narrow tensor (C <= 3), ldd % 8 == 0
template <typename TS, typename TD, bool VEC>
__global__ void unfold3_grad_kernel(const TS* __restrict__ src, int64_t lds, int C, TD* __restrict__ dst, int64_t ldd,
                                    int N, int H, int W) {
  int64_t total = (int64_t)N * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int64_t r = i / W;
    const int y = (int)(r % H);
    float v[3][9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y - (tap / 3 - 1), xx = x - (tap % 3 - 1);
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const TS* sp = src + (i + (int64_t)(yy - y) * W + (xx - x)) * lds;
      if (VEC) {                                           // bf16, 8-byte aligned pixels: one load per neighbour
        const f4 q = ok ? ld4(sp) : f4{{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int o = 0; o < 3; ++o) v[o][tap] = o < C ? q.v[o] : 0.f;
      } else {
#pragma unroll
        for (int o = 0; o < 3; ++o) v[o][tap] = (ok && o < C) ? ldf(sp + o) : 0.f;
      }
    }
    TD* d = dst + i * ldd;
    if (VEC && ldd == 32) {                                // (VEC also asserts a 32-byte aligned dst: full-sector stores)
#pragma unroll
      for (int c16 = 0; c16 < 32; c16 += 16) {
        f16v o16;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int col = c16 + k;
          o16.v[k] = col < 27 ? v[col / 9][col % 9] : 0.f;
        }
        st16(d + c16, o16);
      }
      continue;
    }
#pragma unroll
    for (int c8 = 0; c8 < 32; c8 += 8) {                   // (ldd <= 32; static columns: no dynamic register indexing)
      if (c8 >= (int)ldd) break;
      f8 o8;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int col = c8 + k;
        o8.v[k] = col < 27 ? v[col / 9][col % 9] : 0.f;
      }
      st8(d + c8, o8);
    }
  }
}

NV_API int nervecl_unfold3_grad(const void* src, int64_t lds, int src_dtype, int C, void* dst, int64_t ldd,
                                int dst_dtype, int N, int H, int W, nervecl_stream_t stream) {
  if (!src || !dst || N <= 0 || H <= 0 || W <= 0 || C <= 0 || C > 3 || lds < C || ldd < 9 * C || ldd > 32) return NERVECL_EINVAL;
  if ((ldd & 7) || !aligned(dst, 16)) return NERVECL_EALIGN;
  const int64_t total = (int64_t)N * H * W;
  const int blocks = (int)imin(cdiv(total, 256), kSMs * 16);
  cudaStream_t s = as_stream(stream);
#define NV_UNF(TS, TD, V) unfold3_grad_kernel<TS, TD, V><<<blocks, 256, 0, s>>>((const TS*)src, lds, C, (TD*)dst, ldd, N, H, W)
  if (src_dtype == NERVECL_F32 && dst_dtype == NERVECL_F32) NV_UNF(float, float, false);
  else if (src_dtype == NERVECL_F32 && dst_dtype == NERVECL_BF16) NV_UNF(float, bf16, false);
  else if (src_dtype == NERVECL_BF16 && dst_dtype == NERVECL_BF16) {
    if (lds >= 4 && !(lds & 3) && aligned(src, 8) && aligned(dst, 32)) NV_UNF(bf16, bf16, true); else NV_UNF(bf16, bf16, false);
  } else return NERVECL_EDTYPE;
#undef NV_UNF
  return launch_status();
}

// ---------------------------------------------------------------------------------------
// NHWC slice <-> NCHW fp32 via a 32x32 shared-memory transpose over (pixel, channel)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, int64_t ld, float* __restrict__ dst,
                                    int C, int64_t HW) {
  __shared__ float tile[32][33];
  int n = blockIdx.z;
  int64_t p0 = (int64_t)blockIdx.x * 32;
  int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t p = p0 + i;
    int c = c0 + threadIdx.x;
    if (p < HW && c < C) tile[i][threadIdx.x] = ldf(src + (n * HW + p) * ld + c);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    int64_t p = p0 + threadIdx.x;
    if (p < HW && c < C) dst[((int64_t)n * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t ld,
                                    int C, int64_t HW) {
  __shared__ float tile[32][33];
  int n = blockIdx.z;
  int64_t p0 = (int64_t)blockIdx.x * 32;
  int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    int64_t p = p0 + threadIdx.x;
    if (p < HW && c < C) tile[i][threadIdx.x] = __ldg(src + ((int64_t)n * C + c) * HW + p);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t p = p0 + i;
    int c = c0 + threadIdx.x;
    if (p < HW && c < C) stf(dst + (n * HW + p) * ld + c, tile[threadIdx.x][i]);
  }
}

NV_API int nervecl_nhwc_to_nchw(const void* src, int64_t ld, int dtype, float* dst, int N, int C, int H,
                                int W, nervecl_stream_t stream) {
  if (!src || !dst || N <= 0 || C <= 0 || H <= 0 || W <= 0 || ld < C) return NERVECL_EINVAL;
  int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)cdiv(HW, 32), (unsigned)cdiv(C, 32), N), block(32, 8);
  NV_DISPATCH_DTYPE(dtype, E, (nhwc_to_nchw_kernel<E><<<grid, block, 0, as_stream(stream)>>>(
                                  (const E*)src, ld, dst, C, HW)));
  return launch_status();
}

NV_API int nervecl_nchw_to_nhwc(const float* src, void* dst, int64_t ld, int dtype, int N, int C, int H,
                                int W, nervecl_stream_t stream) {
  if (!src || !dst || N <= 0 || C <= 0 || H <= 0 || W <= 0 || ld < C) return NERVECL_EINVAL;
  int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)cdiv(HW, 32), (unsigned)cdiv(C, 32), N), block(32, 8);
  NV_DISPATCH_DTYPE(dtype, E, (nchw_to_nhwc_kernel<E><<<grid, block, 0, as_stream(stream)>>>(
                                  src, (E*)dst, ld, C, HW)));
  return launch_status();
}

// ---------------------------------------------------------------------------------------
// OIHW fp32 -> [tap][O][Ipad]   (or the transposed / 180-degree-rotated operator)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ dst, int O, int I, int KK,
                                   int nrows, int cols, int rows, int pad_to, int transpose_flip) {
  // dst[tap][r][c], r < rows (= rows_pad), c < pad_to (= cols_pad); data where r < nrows, c < cols
  int64_t total = (int64_t)KK * rows * pad_to;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % pad_to);
    int64_t r2 = i / pad_to;
    int r = (int)(r2 % rows);
    int tap = (int)(r2 / rows);
    float v = 0.f;
    if (c < cols && r < nrows) {
      if (!transpose_flip) {
        v = __ldg(w + ((int64_t)r * I + c) * KK + tap);  // o = r, i = c
      } else {
        v = __ldg(w + ((int64_t)c * I + r) * KK + (KK - 1 - tap));  // o = c, i = r, rotated tap
      }
    }
    stf(dst + i, v);
  }
}

NV_API int nervecl_pack_conv_weight(const float* w_oihw, void* dst, int dtype, int O, int I, int KH, int KW,
                                    int rows_pad, int cols_pad, int transpose_flip, nervecl_stream_t stream) {
  if (!w_oihw || !dst || O <= 0 || I <= 0 || KH <= 0 || KW != KH) return NERVECL_EINVAL;
  int nrows = transpose_flip ? I : O;
  int cols = transpose_flip ? O : I;
  if (cols_pad < cols || rows_pad < nrows) return NERVECL_EINVAL;
  int KK = KH * KW;
  int rows = rows_pad, pad_to = cols_pad;
  int64_t total = (int64_t)KK * rows * pad_to;
  int blocks = (int)imin(cdiv(total, 256), kSMs * 8);
  NV_DISPATCH_DTYPE(dtype, E, (pack_weight_kernel<E><<<blocks, 256, 0, as_stream(stream)>>>(
                                  w_oihw, (E*)dst, O, I, KK, nrows, cols, rows, pad_to, transpose_flip)));
  return launch_status();
}

// Batched form: up to kPackBatch weights per launch (descriptor table in the kernel parameters), block
// (blockIdx.y = entry).  Replaces ~130 tiny launches per training step by 3-4.
constexpr int kPackBatch = 48;
struct PackEntry { const float* w; void* dst; int O, I, KK, rows, cols_pad, flip; };
struct PackTable { PackEntry e[kPackBatch]; };

template <typename T>
__global__ void pack_weight_batched_kernel(const __grid_constant__ PackTable tab) {
  const PackEntry& q = tab.e[blockIdx.y];
  const int nrows = q.flip ? q.I : q.O, cols = q.flip ? q.O : q.I;
  const int64_t total = (int64_t)q.KK * q.rows * q.cols_pad;
  T* dst = reinterpret_cast<T*>(q.dst);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % q.cols_pad);
    const int64_t r2 = i / q.cols_pad;
    const int r = (int)(r2 % q.rows), tap = (int)(r2 / q.rows);
    float v = 0.f;
    if (c < cols && r < nrows)
      v = q.flip ? __ldg(q.w + ((int64_t)c * q.I + r) * q.KK + (q.KK - 1 - tap)) : __ldg(q.w + ((int64_t)r * q.I + c) * q.KK + tap);
    stf(dst + i, v);
  }
}

NV_API int nervecl_pack_conv_weights_batched(int n, const float* const* w_host, void* const* dst_host, const int32_t* O_host,
                                             const int32_t* I_host, const int32_t* K_host, const int32_t* rows_pad_host,
                                             const int32_t* cols_pad_host, const int32_t* flip_host, int dtype,
                                             nervecl_stream_t stream) {
  if (n < 0 || (n && (!w_host || !dst_host || !O_host || !I_host || !K_host || !rows_pad_host || !cols_pad_host || !flip_host)))
    return NERVECL_EINVAL;
  if (dtype != NERVECL_F32 && dtype != NERVECL_BF16) return NERVECL_EDTYPE;
  for (int base = 0; base < n; base += kPackBatch) {
    const int m = (int)imin(kPackBatch, n - base);
    PackTable tab;
    int64_t biggest = 0;
    for (int j = 0; j < m; ++j) {
      const int i = base + j;
      const int nrows = flip_host[i] ? I_host[i] : O_host[i], cols = flip_host[i] ? O_host[i] : I_host[i];
      if (!w_host[i] || !dst_host[i] || O_host[i] <= 0 || I_host[i] <= 0 || K_host[i] <= 0 || cols_pad_host[i] < cols ||
          rows_pad_host[i] < nrows)
        return NERVECL_EINVAL;
      tab.e[j] = PackEntry{w_host[i], dst_host[i], O_host[i], I_host[i], K_host[i] * K_host[i], rows_pad_host[i],
                           cols_pad_host[i], flip_host[i]};
      biggest = imax(biggest, (int64_t)tab.e[j].KK * rows_pad_host[i] * cols_pad_host[i]);
    }
    for (int j = m; j < kPackBatch; ++j) tab.e[j] = tab.e[0];
    dim3 grid((unsigned)imax(1, imin(cdiv(biggest, 256 * 4), 64)), (unsigned)m);
    if (dtype == NERVECL_F32) pack_weight_batched_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(tab);
    else pack_weight_batched_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(tab);
    int rc = launch_status();
    if (rc) return rc;
  }
  return NERVECL_OK;
}

// ---------------------------------------------------------------------------------------
// out = (acc ? out : 0) + alpha * x
// ---------------------------------------------------------------------------------------
template <typename TX, typename TO>
__global__ void axpy_kernel(const TX* __restrict__ x, int64_t ldx, TO* __restrict__ out, int64_t ldo,
                            int64_t npix, int C, float alpha, int accumulate) {
  int c4n = C >> 2;
  int64_t total = npix * c4n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / c4n;
    int c = (int)(i % c4n) << 2;
    f4 a = ld4(x + p * ldx + c);
    f4 o;
    if (accumulate) {
      o = ld4(out + p * ldo + c);
#pragma unroll
      for (int k = 0; k < 4; ++k) o.v[k] += alpha * a.v[k];
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) o.v[k] = alpha * a.v[k];
    }
    st4(out + p * ldo + c, o);
  }
}

// 8 elements per thread and access (16 bytes of bf16; the 4-wide kernel moves 8), two independent items in flight, 32-bit
// index arithmetic: 0.70 -> ~0.9 of the copy bandwidth on the [B, H, W, 64] activations the engine adds / copies.
template <typename TX, typename TO>
__global__ void __launch_bounds__(256)
axpy8_kernel(const TX* __restrict__ x, int64_t ldx, TO* __restrict__ out, int64_t ldo, uint32_t total, int C,
             float alpha, int accumulate) {
  const uint32_t c8n = (uint32_t)C >> 3;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    const uint32_t j = i + stride;
    const bool two = j < total;
    const uint32_t p0 = i / c8n, p1 = two ? j / c8n : p0;
    const uint32_t c0 = (i - p0 * c8n) << 3, c1 = two ? (j - p1 * c8n) << 3 : c0;
    const TX* x0 = x + (int64_t)p0 * ldx + c0;
    const TX* x1 = x + (int64_t)p1 * ldx + c1;
    TO* o0 = out + (int64_t)p0 * ldo + c0;
    TO* o1 = out + (int64_t)p1 * ldo + c1;
    const f8 a0 = ld8(x0), a1 = ld8(x1);
    f8 r0, r1;
    if (accumulate) {
      r0 = ld8(const_cast<const TO*>(o0));
      r1 = ld8(const_cast<const TO*>(o1));
#pragma unroll
      for (int k = 0; k < 8; ++k) { r0.v[k] += alpha * a0.v[k]; r1.v[k] += alpha * a1.v[k]; }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) { r0.v[k] = alpha * a0.v[k]; r1.v[k] = alpha * a1.v[k]; }
    }
    st8(o0, r0);
    if (two) st8(o1, r1);
  }
}

template <typename TX, typename TO>
__global__ void axpy_scalar_kernel(const TX* __restrict__ x, int64_t ldx, TO* __restrict__ out, int64_t ldo,
                                   int64_t npix, int C, float alpha, int accumulate) {
  int64_t total = npix * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / C;
    int c = (int)(i % C);
    float v = alpha * ldf(x + p * ldx + c);
    if (accumulate) v += ldf(const_cast<const TO*>(out) + p * ldo + c);
    stf(out + p * ldo + c, v);
  }
}

NV_API int nervecl_axpy(const void* x, int64_t ldx, int x_dtype, void* out, int64_t ldo, int out_dtype,
                        int64_t npix, int C, float alpha, int accumulate, nervecl_stream_t stream) {
  if (!x || !out || npix <= 0 || C <= 0) return NERVECL_EINVAL;
  const bool vec = !((C & 3) || (ldx & 3) || (ldo & 3) || !aligned(x, 16) || !aligned(out, 16));
  const bool vec8 = vec && !((C & 7) || (ldx & 7) || (ldo & 7)) && npix * (C >> 3) < ((int64_t)1 << 31);
  int64_t total = vec ? npix * (C >> 2) : npix * C;
  int blocks = (int)imin(cdiv(total, 256), kSMs * 16);
  const int blocks8 = (int)imin(cdiv(npix * (C >> 3), 512), kSMs * 8);
  cudaStream_t s = as_stream(stream);
#define LAUNCH(TX, TO)                                                                                          \
  do {                                                                                                          \
    if (vec8)                                                                                                   \
      axpy8_kernel<TX, TO><<<blocks8, 256, 0, s>>>((const TX*)x, ldx, (TO*)out, ldo, (uint32_t)(npix * (C >> 3)), C, alpha, \
                                                   accumulate);                                                 \
    else if (vec)                                                                                               \
      axpy_kernel<TX, TO><<<blocks, 256, 0, s>>>((const TX*)x, ldx, (TO*)out, ldo, npix, C, alpha, accumulate); \
    else                                                                                                        \
      axpy_scalar_kernel<TX, TO><<<blocks, 256, 0, s>>>((const TX*)x, ldx, (TO*)out, ldo, npix, C, alpha,       \
                                                        accumulate);                                            \
  } while (0)
  if (x_dtype == NERVECL_F32 && out_dtype == NERVECL_F32) LAUNCH(float, float);
  else if (x_dtype == NERVECL_F32 && out_dtype == NERVECL_BF16) LAUNCH(float, bf16);
  else if (x_dtype == NERVECL_BF16 && out_dtype == NERVECL_F32) LAUNCH(bf16, float);
  else if (x_dtype == NERVECL_BF16 && out_dtype == NERVECL_BF16) LAUNCH(bf16, bf16);
  else return NERVECL_EDTYPE;
#undef LAUNCH
  return launch_status();
}

// ---------------------------------------------------------------------------------------
// out = dy * [(y - y_sub) > 0]
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ dy, int64_t lddy, const T* __restrict__ y, int64_t ldy,
                                const T* __restrict__ ysub, int64_t ldys, T* __restrict__ out, int64_t ldo,
                                int64_t npix, int C) {
  int c4n = C >> 2;
  int64_t total = npix * c4n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / c4n;
    int c = (int)(i % c4n) << 2;
    f4 g = ld4(dy + p * lddy + c);
    f4 v = ld4(y + p * ldy + c);
    if (ysub) {
      f4 s = ld4(ysub + p * ldys + c);
#pragma unroll
      for (int k = 0; k < 4; ++k) v.v[k] -= s.v[k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) g.v[k] = v.v[k] > 0.f ? g.v[k] : 0.f;
    st4(out + p * ldo + c, g);
  }
}

NV_API int nervecl_relu_bwd(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* y_sub,
                            int64_t ldys, void* out, int64_t ldo, int dtype, int64_t npix, int C,
                            nervecl_stream_t stream) {
  if (!dy || !y || !out || npix <= 0 || C <= 0) return NERVECL_EINVAL;
  if ((C & 3) || (lddy & 3) || (ldy & 3) || (ldo & 3) || (y_sub && (ldys & 3))) return NERVECL_EALIGN;
  int64_t total = npix * (C >> 2);
  int blocks = (int)imin(cdiv(total, 256), kSMs * 16);
  NV_DISPATCH_DTYPE(dtype, E, (relu_bwd_kernel<E><<<blocks, 256, 0, as_stream(stream)>>>(
                                  (const E*)dy, lddy, (const E*)y, ldy, (const E*)y_sub, ldys, (E*)out, ldo,
                                  npix, C)));
  return launch_status();
}

NV_API int nervecl_fill_zero(void* p, size_t bytes, nervecl_stream_t stream) {
  if (!p) return NERVECL_EINVAL;
  if (bytes == 0) return NERVECL_OK;
  cudaError_t e = cudaMemsetAsync(p, 0, bytes, as_stream(stream));
  return e == cudaSuccess ? NERVECL_OK : (int)e;
}

// ---------------------------------------------------------------------------------------
// MSE forward + backward in one pass
// ---------------------------------------------------------------------------------------
__global__ void mse_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dgrad,
                           float* __restrict__ loss, int64_t n, float scale) {
  double acc = 0.0;
  int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = reinterpret_cast<const float4*>(a)[i];
    float4 y = reinterpret_cast<const float4*>(b)[i];
    float4 d = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
    float s = d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
    acc += (double)s;
    if (dgrad) {
      float k = 2.f * scale;
      reinterpret_cast<float4*>(dgrad)[i] = make_float4(k * d.x, k * d.y, k * d.z, k * d.w);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = n4 << 2; i < n; ++i) {
      float d = a[i] - b[i];
      acc += (double)d * d;
      if (dgrad) dgrad[i] = 2.f * scale * d;
    }
  }
  acc = warp_sum(acc);
  __shared__ double part[8];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) part[w] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    atomicAdd(loss, (float)(t * (double)scale));
  }
}

NV_API int nervecl_mse_fwd_bwd(const float* a, const float* b, float* dgrad, float* loss, int64_t n,
                               float scale, nervecl_stream_t stream) {
  if (!a || !b || !loss || n <= 0) return NERVECL_EINVAL;
  if (!aligned(a, 16) || !aligned(b, 16) || (dgrad && !aligned(dgrad, 16))) return NERVECL_EALIGN;
  int blocks = (int)imin(cdiv(n >> 2, 256) + 1, kSMs * 8);
  mse_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a, b, dgrad, loss, n, scale);
  return launch_status();
}
