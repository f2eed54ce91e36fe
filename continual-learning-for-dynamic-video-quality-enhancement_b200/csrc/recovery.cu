// Kernels of the FrameRecoveryNet trunk that the SuperResolutionNet path does not already have
// (reference nerve_cl/models/frame_recovery.py): the strided / 7x7 direct convolution of the encoder stem and
// down-sampling shortcuts, max pooling, depth-to-space (ConvTranspose2d(4,2,1) runs as a 3x3 convolution with
// 4*Cout outputs through the tcgen05 conv + this re-layout), bilinear resize, the FusionModule's attention blend
// and the final tanh / resize / mask blend.  All NHWC, pixel pitch `ld`, fp32 or bf16 activations.  The dense
// stride-1 3x3 / 1x1 convolutions of the network (94 % of its FLOPs) go through nervecl_conv2d_fwd.
#include "common.cuh"

using namespace nv;

namespace {

inline int ew_blocks(int64_t work) { return (int)imax(1, imin(cdiv(work, 256), kSMs * 16)); }

// ---------------------------------------------------------------------------------------------------------------
// Direct convolution, any odd/even kernel, stride and padding; one thread = one output pixel x 16 output channels,
// weights of an 8-input-channel chunk staged in shared memory.  Used only for the low-FLOP strided layers
// (stem 7x7/2 over 4 channels: 25 kFLOP per output pixel; 1x1/2 shortcuts).
// ---------------------------------------------------------------------------------------------------------------
constexpr int DC_CO = 16, DC_CK = 8;

template <typename TI, typename TO>
__global__ void __launch_bounds__(128)
conv_direct_kernel(const TI* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ bias,
                   TO* __restrict__ out, int64_t ldo, int N, int H, int W, int Cin, int Cout, int K, int stride, int pad,
                   int OH, int OW, int relu) {
  extern __shared__ float w_s[];                     // [DC_CK][K*K][DC_CO]
  const int co0 = blockIdx.y * DC_CO;
  const int64_t npix = (int64_t)N * OH * OW;
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool live = p < npix;
  int ox = 0, oy = 0, n = 0;
  if (live) {
    ox = (int)(p % OW);
    const int64_t r = p / OW;
    oy = (int)(r % OH);
    n = (int)(r / OH);
  }
  float acc[DC_CO];
#pragma unroll
  for (int j = 0; j < DC_CO; ++j) acc[j] = 0.f;
  const int KK = K * K;
  for (int c0 = 0; c0 < Cin; c0 += DC_CK) {
    const int ck = min(DC_CK, Cin - c0);
    __syncthreads();
    for (int e = threadIdx.x; e < DC_CK * KK * DC_CO; e += blockDim.x) {
      const int co = e % DC_CO, r = e / DC_CO, tap = r % KK, ci = r / KK;
      float v = 0.f;
      if (ci < ck && co0 + co < Cout) v = __ldg(w + ((int64_t)(co0 + co) * Cin + c0 + ci) * KK + tap);   // OIHW
      w_s[e] = v;
    }
    __syncthreads();
    if (!live) continue;
    for (int ky = 0; ky < K; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < K; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= W) continue;
        const TI* xp = x + (((int64_t)n * H + iy) * W + ix) * ldx + c0;
        const int tap = ky * K + kx;
        for (int ci = 0; ci < ck; ++ci) {
          const float v = ldf(xp + ci);
          const float4* w4 = reinterpret_cast<const float4*>(w_s + (ci * KK + tap) * DC_CO);
#pragma unroll
          for (int q = 0; q < DC_CO / 4; ++q) {
            const float4 ww = w4[q];
            acc[4 * q] = fmaf(v, ww.x, acc[4 * q]);
            acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
          }
        }
      }
    }
  }
  if (!live) return;
  TO* op = out + p * ldo + co0;
#pragma unroll
  for (int j = 0; j < DC_CO; ++j) {
    if (co0 + j >= Cout) break;
    float v = acc[j] + (bias ? __ldg(bias + co0 + j) : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    stf(op + j, v);
  }
}

// max pooling k x k / stride / pad (padding never wins: -inf), 8 channels per thread
template <typename E>
__global__ void __launch_bounds__(256)
maxpool_kernel(const E* __restrict__ x, int64_t ldx, E* __restrict__ y, int64_t ldy, int N, int H, int W, int C, int k,
               int stride, int pad, int OH, int OW) {
  const int cg = C >> 3;
  const int64_t total = (int64_t)N * OH * OW * cg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    int64_t r = i / cg;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int n = (int)(r / OH);
    f8 m;
#pragma unroll
    for (int j = 0; j < 8; ++j) m.v[j] = -INFINITY;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= W) continue;
        const f8 v = ld8(x + (((int64_t)n * H + iy) * W + ix) * ldx + g * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) m.v[j] = fmaxf(m.v[j], v.v[j]);
      }
    }
    st8(y + (((int64_t)n * OH + oy) * OW + ox) * ldy + g * 8, m);
  }
}

// out[n, s*y + i, s*x + j, c] = in[n, y, x, (i*s + j)*C + c]
template <typename E>
__global__ void __launch_bounds__(256)
depth_to_space_kernel(const E* __restrict__ x, int64_t ldx, E* __restrict__ y, int64_t ldy, int N, int H, int W, int C,
                      int s) {
  const int cg = C >> 3;
  const int OH = H * s, OW = W * s;
  const int64_t total = (int64_t)N * OH * OW * cg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    int64_t r = i / cg;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int n = (int)(r / OH);
    const int yy = oy / s, ii = oy % s, xx = ox / s, jj = ox % s;
    const f8 v = ld8(x + (((int64_t)n * H + yy) * W + xx) * ldx + (ii * s + jj) * C + g * 8);
    st8(y + (((int64_t)n * OH + oy) * OW + ox) * ldy + g * 8, v);
  }
}

// ATen's bilinear source index, align_corners=False: max(scale * (dst + 0.5) - 0.5, 0)
__device__ __forceinline__ void bilinear_coord(int dst, int in_size, float scale, int& i0, int& i1, float& l1) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i0 = i0 > in_size - 1 ? in_size - 1 : i0;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l1 = l1 > 1.f ? 1.f : l1;
}

template <typename E>
__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const E* __restrict__ x, int64_t ldx, E* __restrict__ y, int64_t ldy, int N, int H, int W, int C,
                       int OH, int OW, float sh, float sw) {
  const int cg = C >> 3;
  const int64_t total = (int64_t)N * OH * OW * cg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    int64_t r = i / cg;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int n = (int)(r / OH);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_coord(oy, H, sh, y0, y1, ly);
    bilinear_coord(ox, W, sw, x0, x1, lx);
    const E* b = x + (int64_t)n * H * W * ldx + g * 8;
    const f8 a = ld8(b + ((int64_t)y0 * W + x0) * ldx), bq = ld8(b + ((int64_t)y0 * W + x1) * ldx);
    const f8 c = ld8(b + ((int64_t)y1 * W + x0) * ldx), d = ld8(b + ((int64_t)y1 * W + x1) * ldx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    f8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = hy * (hx * a.v[j] + lx * bq.v[j]) + ly * (hx * c.v[j] + lx * d.v[j]);
    st8(y + (((int64_t)n * OH + oy) * OW + ox) * ldy + g * 8, o);
  }
}

// FusionModule (frame_recovery.py:232-256): per pixel  a = softmax(logits[0:2]);
//   out[c] = aligned[c] + a0 * mean_c(spatial) + a1 * mean_c(temporal)     (the two all-ones/C 1x1 convs are channel means)
// one warp per pixel
template <typename E>
__global__ void __launch_bounds__(256)
fusion_blend_kernel(const E* __restrict__ aligned, int64_t lda, const float* __restrict__ logits, int64_t ldl,
                    const E* __restrict__ sp, int64_t lds, int Cs, const E* __restrict__ tp, int64_t ldt, int Ct,
                    E* __restrict__ out, int64_t ldo, int C, int64_t npix) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < npix; p += nwarps) {
    float ss = 0.f, st = 0.f;
    for (int c = lane * 8; c < Cs; c += 256) {
      const f8 v = ld8(sp + p * lds + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += v.v[j];
    }
    for (int c = lane * 8; c < Ct; c += 256) {
      const f8 v = ld8(tp + p * ldt + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) st += v.v[j];
    }
    ss = warp_sum(ss) / (float)Cs;
    st = warp_sum(st) / (float)Ct;
    const float l0 = __ldg(logits + p * ldl), l1 = __ldg(logits + p * ldl + 1);
    const float m = fmaxf(l0, l1);
    const float e0 = __expf(l0 - m), e1 = __expf(l1 - m);
    const float f = (e0 * ss + e1 * st) / (e0 + e1);
    for (int c = lane * 8; c < C; c += 256) {
      f8 v = ld8(aligned + p * lda + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] += f;
      st8(out + p * ldo + c, v);
    }
  }
}

// FrameRecoveryNet tail (frame_recovery.py:430-442): recovered = tanh(decoder conv) [bilinearly resized to (H, W) when
// the decoder's (OHd, OWd) differs]; out = frame * (1 - mask) + recovered * mask.  NCHW fp32 in / out, conv NHWC fp32.
__global__ void __launch_bounds__(256)
recovery_finish_kernel(const float* __restrict__ conv, int64_t ldc, int Hd, int Wd, const float* __restrict__ frame,
                       const float* __restrict__ mask, float* __restrict__ out, int N, int C, int H, int W, float sh,
                       float sw) {
  const int64_t total = (int64_t)N * C * H * W;
  const bool same = Hd == H && Wd == W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    int64_t r = i / W;
    const int y = (int)(r % H);
    r /= H;
    const int c = (int)(r % C);
    const int n = (int)(r / C);
    const float* b = conv + (int64_t)n * Hd * Wd * ldc + c;
    float v;
    if (same) {
      v = tanhf(b[((int64_t)y * Wd + x) * ldc]);
    } else {
      int y0, y1, x0, x1;
      float ly, lx;
      bilinear_coord(y, Hd, sh, y0, y1, ly);
      bilinear_coord(x, Wd, sw, x0, x1, lx);
      const float a = tanhf(b[((int64_t)y0 * Wd + x0) * ldc]), bq = tanhf(b[((int64_t)y0 * Wd + x1) * ldc]);
      const float cc = tanhf(b[((int64_t)y1 * Wd + x0) * ldc]), d = tanhf(b[((int64_t)y1 * Wd + x1) * ldc]);
      v = (1.f - ly) * ((1.f - lx) * a + lx * bq) + ly * ((1.f - lx) * cc + lx * d);
    }
    const float m = mask ? mask[((int64_t)n * H + y) * W + x] : 0.f;
    out[i] = frame[i] * (1.f - m) + v * m;
  }
}

}  // namespace

NV_API int nervecl_conv2d_direct(const void* x, int64_t ldx, int dtype, const float* w_oihw, const float* bias, void* out,
                                 int64_t ldo, int out_dtype, int N, int H, int W, int Cin, int Cout, int K, int stride,
                                 int pad, int relu, nervecl_stream_t stream) {
  if (!x || !w_oihw || !out || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || K < 1 || K > 7 || stride < 1 ||
      pad < 0 || ldx < Cin || ldo < Cout)
    return NERVECL_EINVAL;
  const int OH = (H + 2 * pad - K) / stride + 1, OW = (W + 2 * pad - K) / stride + 1;
  if (OH <= 0 || OW <= 0) return NERVECL_EINVAL;
  const int64_t npix = (int64_t)N * OH * OW;
  dim3 grid((unsigned)cdiv(npix, 128), (unsigned)cdiv(Cout, DC_CO));
  const size_t smem = (size_t)DC_CK * K * K * DC_CO * sizeof(float);
  if (dtype == NERVECL_F32 && out_dtype == NERVECL_F32)
    conv_direct_kernel<float, float><<<grid, 128, smem, as_stream(stream)>>>((const float*)x, ldx, w_oihw, bias, (float*)out, ldo,
                                                                              N, H, W, Cin, Cout, K, stride, pad, OH, OW, relu);
  else if (dtype == NERVECL_BF16 && out_dtype == NERVECL_BF16)
    conv_direct_kernel<bf16, bf16><<<grid, 128, smem, as_stream(stream)>>>((const bf16*)x, ldx, w_oihw, bias, (bf16*)out, ldo, N,
                                                                            H, W, Cin, Cout, K, stride, pad, OH, OW, relu);
  else if (dtype == NERVECL_BF16 && out_dtype == NERVECL_F32)
    conv_direct_kernel<bf16, float><<<grid, 128, smem, as_stream(stream)>>>((const bf16*)x, ldx, w_oihw, bias, (float*)out, ldo,
                                                                             N, H, W, Cin, Cout, K, stride, pad, OH, OW, relu);
  else if (dtype == NERVECL_F32 && out_dtype == NERVECL_BF16)
    conv_direct_kernel<float, bf16><<<grid, 128, smem, as_stream(stream)>>>((const float*)x, ldx, w_oihw, bias, (bf16*)out, ldo,
                                                                             N, H, W, Cin, Cout, K, stride, pad, OH, OW, relu);
  else
    return NERVECL_EDTYPE;
  return launch_status();
}

NV_API int nervecl_maxpool2d(const void* x, int64_t ldx, void* y, int64_t ldy, int dtype, int N, int H, int W, int C, int k,
                             int stride, int pad, nervecl_stream_t stream) {
  if (!x || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0 || k < 1 || stride < 1 || pad < 0 || 2 * pad > k) return NERVECL_EINVAL;
  if ((C & 7) || (ldx & 7) || (ldy & 7) || !aligned(x, 16) || !aligned(y, 16)) return NERVECL_EALIGN;
  const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
  if (OH <= 0 || OW <= 0) return NERVECL_EINVAL;
  const int64_t total = (int64_t)N * OH * OW * (C >> 3);
  NV_DISPATCH_DTYPE(dtype, E, (maxpool_kernel<E><<<ew_blocks(total), 256, 0, as_stream(stream)>>>(
                                  (const E*)x, ldx, (E*)y, ldy, N, H, W, C, k, stride, pad, OH, OW)));
  return launch_status();
}

NV_API int nervecl_depth_to_space(const void* x, int64_t ldx, void* y, int64_t ldy, int dtype, int N, int H, int W, int C,
                                  int s, nervecl_stream_t stream) {
  if (!x || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0 || s < 1 || s > 8 || ldx < (int64_t)C * s * s || ldy < C)
    return NERVECL_EINVAL;
  if ((C & 7) || (ldx & 7) || (ldy & 7) || !aligned(x, 16) || !aligned(y, 16)) return NERVECL_EALIGN;
  const int64_t total = (int64_t)N * H * s * W * s * (C >> 3);
  NV_DISPATCH_DTYPE(dtype, E, (depth_to_space_kernel<E><<<ew_blocks(total), 256, 0, as_stream(stream)>>>(
                                  (const E*)x, ldx, (E*)y, ldy, N, H, W, C, s)));
  return launch_status();
}

NV_API int nervecl_resize_bilinear(const void* x, int64_t ldx, void* y, int64_t ldy, int dtype, int N, int H, int W, int C,
                                   int OH, int OW, nervecl_stream_t stream) {
  if (!x || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0 || OH <= 0 || OW <= 0) return NERVECL_EINVAL;
  if ((C & 7) || (ldx & 7) || (ldy & 7) || !aligned(x, 16) || !aligned(y, 16)) return NERVECL_EALIGN;
  const int64_t total = (int64_t)N * OH * OW * (C >> 3);
  const float sh = (float)H / (float)OH, sw = (float)W / (float)OW;
  NV_DISPATCH_DTYPE(dtype, E, (resize_bilinear_kernel<E><<<ew_blocks(total), 256, 0, as_stream(stream)>>>(
                                  (const E*)x, ldx, (E*)y, ldy, N, H, W, C, OH, OW, sh, sw)));
  return launch_status();
}

NV_API int nervecl_fusion_blend(const void* aligned_, int64_t lda, const float* logits, int64_t ldl, const void* spatial,
                                int64_t lds, int Cs, const void* temporal, int64_t ldt, int Ct, void* out, int64_t ldo,
                                int dtype, int C, int64_t npix, nervecl_stream_t stream) {
  if (!aligned_ || !logits || !spatial || !temporal || !out || C <= 0 || Cs <= 0 || Ct <= 0 || npix <= 0 || ldl < 2)
    return NERVECL_EINVAL;
  if (((C | Cs | Ct) & 7) || ((lda | lds | ldt | ldo) & 7) || !aligned(aligned_, 16) || !aligned(spatial, 16) ||
      !aligned(temporal, 16) || !aligned(out, 16))
    return NERVECL_EALIGN;
  const int blocks = (int)imax(1, imin(cdiv(npix, 8), kSMs * 8));
  NV_DISPATCH_DTYPE(dtype, E, (fusion_blend_kernel<E><<<blocks, 256, 0, as_stream(stream)>>>(
                                  (const E*)aligned_, lda, logits, ldl, (const E*)spatial, lds, Cs, (const E*)temporal, ldt,
                                  Ct, (E*)out, ldo, C, npix)));
  return launch_status();
}

NV_API int nervecl_recovery_finish(const float* conv_out, int64_t ldc, int Hd, int Wd, const float* frame, const float* mask,
                                   float* out, int N, int C, int H, int W, nervecl_stream_t stream) {
  if (!conv_out || !frame || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0 || Hd <= 0 || Wd <= 0 || ldc < C)
    return NERVECL_EINVAL;
  const int64_t total = (int64_t)N * C * H * W;
  recovery_finish_kernel<<<ew_blocks(total), 256, 0, as_stream(stream)>>>(conv_out, ldc, Hd, Wd, frame, mask, out, N, C, H, W,
                                                                         (float)Hd / (float)H, (float)Wd / (float)W);
  return launch_status();
}
