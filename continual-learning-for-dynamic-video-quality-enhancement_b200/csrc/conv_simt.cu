// CUDA-core (fp32 accumulate) direct convolution: forward/dgrad with fused epilogue, and wgrad.
// This is the exact-precision engine (fp32 parity path) and the engine for shapes the tcgen05
// implicit GEMM does not take (Cin=3 head, Cout=2 flow, Cout=T logits, 7x7 2->1 spatial gate).
#include "common.cuh"
#include "conv_internal.cuh"

using namespace nv;

namespace {

constexpr int TH = 8, TW = 32;      // output tile (pixels)
constexpr int CO_T = 16;            // output channels per block
constexpr int CK = 8;               // input-channel chunk staged in shared memory
constexpr int CKP = CK + 1;         // padded pitch (odd -> conflict-free across pixels)
constexpr int THREADS = 128;        // each thread: 2 pixels (rows ty and ty+4) x CO_T channels

template <typename TI, typename TO, int K>
__global__ void __launch_bounds__(THREADS)
conv_fwd_kernel(const nervecl_conv_params a) {
  constexpr int R = K / 2;
  constexpr int HH = TH + K - 1, HW = TW + K - 1;
  extern __shared__ float smem[];
  float* in_s = smem;                              // [HH][HW][CKP]
  float* w_s = smem + HH * HW * CKP;               // [K*K][CK][CO_T]

  const int tiles_x = (a.W + TW - 1) / TW;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x % tiles_x;
  const int y0 = tile_y * TH, x0 = tile_x * TW;
  const int co0 = blockIdx.y * CO_T;
  const int n = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // ty in 0..3

  const TI* __restrict__ x = reinterpret_cast<const TI*>(a.x);
  const TI* __restrict__ w = reinterpret_cast<const TI*>(a.w);
  const int64_t img = (int64_t)n * a.H * a.W;

  float acc[2][CO_T];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < a.Cin; c0 += CK) {
    const int ck = min(CK, a.Cin - c0);
    // stage input halo tile
    for (int e = threadIdx.x; e < HH * HW * CK; e += THREADS) {
      int c = e % CK;
      int hp = e / CK;
      int hx = hp % HW, hy = hp / HW;
      int gy = y0 + hy - R, gx = x0 + hx - R;
      float v = 0.f;
      if (c < ck && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W)
        v = ldf(x + (img + (int64_t)gy * a.W + gx) * a.ldx + c0 + c);
      in_s[hp * CKP + c] = v;
    }
    // stage weights: w_s[tap][ci][co] = w[tap][co0+co][c0+ci]
    for (int e = threadIdx.x; e < K * K * CO_T * CK; e += THREADS) {
      int ci = e % CK;
      int r = e / CK;
      int co = r % CO_T, tap = r / CO_T;
      float v = 0.f;
      if (ci < ck && co0 + co < a.Cout) v = ldf(w + ((int64_t)tap * a.w_rows + co0 + co) * a.w_ld + c0 + ci);
      w_s[(tap * CK + ci) * CO_T + co] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ky = 0; ky < K; ++ky) {
#pragma unroll 1
      for (int kx = 0; kx < K; ++kx) {
        const float* i0 = in_s + ((ty + ky) * HW + tx + kx) * CKP;
        const float* i1 = i0 + 4 * HW * CKP;
        const float* wp = w_s + (ky * K + kx) * CK * CO_T;
#pragma unroll
        for (int ci = 0; ci < CK; ++ci) {
          float v0 = i0[ci], v1 = i1[ci];
          const float4* w4 = reinterpret_cast<const float4*>(wp + ci * CO_T);
#pragma unroll
          for (int q = 0; q < CO_T / 4; ++q) {
            float4 ww = w4[q];
            acc[0][4 * q + 0] = fmaf(v0, ww.x, acc[0][4 * q + 0]);
            acc[0][4 * q + 1] = fmaf(v0, ww.y, acc[0][4 * q + 1]);
            acc[0][4 * q + 2] = fmaf(v0, ww.z, acc[0][4 * q + 2]);
            acc[0][4 * q + 3] = fmaf(v0, ww.w, acc[0][4 * q + 3]);
            acc[1][4 * q + 0] = fmaf(v1, ww.x, acc[1][4 * q + 0]);
            acc[1][4 * q + 1] = fmaf(v1, ww.y, acc[1][4 * q + 1]);
            acc[1][4 * q + 2] = fmaf(v1, ww.z, acc[1][4 * q + 2]);
            acc[1][4 * q + 3] = fmaf(v1, ww.w, acc[1][4 * q + 3]);
          }
        }
      }
    }
    __syncthreads();
  }

  // fused epilogue (see nervecl.h)
  const TI* __restrict__ res = reinterpret_cast<const TI*>(a.res);
  const TI* __restrict__ mask = reinterpret_cast<const TI*>(a.mask);
  const TI* __restrict__ msub = reinterpret_cast<const TI*>(a.mask_sub);
  TO* __restrict__ out = reinterpret_cast<TO*>(a.out);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int gy = y0 + ty + 4 * i, gx = x0 + tx;
    if (gy >= a.H || gx >= a.W) continue;
    int64_t p = img + (int64_t)gy * a.W + gx;
#pragma unroll
    for (int j = 0; j < CO_T; ++j) {
      int co = co0 + j;
      if (co >= a.Cout) break;
      float v = acc[i][j];
      if (a.bias) v += __ldg(a.bias + co);
      if (a.relu == 1) v = fmaxf(v, 0.f);
      v *= a.alpha;
      if (res && co < a.res_channels) v += ldf(res + p * a.ldres + co);
      if (a.accumulate) v += ldf(const_cast<const TO*>(out) + p * a.ldo + co);
      if (a.relu == 2) v = fmaxf(v, 0.f);
      if (mask && co >= a.mask_c0) {
        float m = ldf(mask + p * a.ldmask + co);
        if (msub) m -= ldf(msub + p * a.ldmask_sub + co);
        if (!(m > 0.f)) v = 0.f;
      }
      stf(out + p * a.ldo + co, v);
    }
  }
}

template <typename TI, typename TO, int K>
int launch_fwd(const nervecl_conv_params& a, cudaStream_t s) {
  constexpr int HH = TH + K - 1, HW = TW + K - 1;
  size_t smem = (size_t)(HH * HW * CKP + K * K * CK * CO_T) * sizeof(float);
  auto kern = conv_fwd_kernel<TI, TO, K>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)(cdiv(a.H, TH) * cdiv(a.W, TW)), (unsigned)cdiv(a.Cout, CO_T), (unsigned)a.N);
  kern<<<grid, THREADS, smem, s>>>(a);
  return launch_status();
}

template <typename TI, typename TO>
int launch_fwd_k(const nervecl_conv_params& a, cudaStream_t s) {
  switch (a.K) {
    case 1: return launch_fwd<TI, TO, 1>(a, s);
    case 3: return launch_fwd<TI, TO, 3>(a, s);
    case 7: return launch_fwd<TI, TO, 7>(a, s);
    default: return NERVECL_EUNSUPPORTED;
  }
}

// ---------------------------------------------------------------------------------------
// wgrad: each thread owns (ci, co and co+16) x K*K taps; blocks stride over pixel tiles.
// ---------------------------------------------------------------------------------------
constexpr int WG_CI = 16, WG_CO = 32, WG_THREADS = 256;

template <typename T, int K>
__global__ void __launch_bounds__(WG_THREADS)
conv_wgrad_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t ldy,
                  float* __restrict__ dw, float* __restrict__ db, int N, int H, int W, int Cin, int Cout,
                  float scale) {
  constexpr int R = K / 2;
  constexpr int HH = TH + K - 1, HWd = TW + K - 1;
  constexpr int XP = WG_CI + 1;
  extern __shared__ float smem[];
  float* x_s = smem;                         // [HH][HWd][XP]
  float* dy_s = smem + HH * HWd * XP;        // [TH*TW][WG_CO]

  const int ci0 = blockIdx.x * WG_CI, co0 = blockIdx.y * WG_CO;
  const int ci = threadIdx.x & 15, cot = threadIdx.x >> 4;  // cot in 0..15
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const int64_t ntiles = (int64_t)N * tiles_x * tiles_y;

  float acc0[K * K], acc1[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc0[t] = acc1[t] = 0.f;
  float bsum0 = 0.f, bsum1 = 0.f;

  for (int64_t tile = blockIdx.z; tile < ntiles; tile += gridDim.z) {
    int n = (int)(tile / (tiles_x * tiles_y));
    int tr = (int)(tile % (tiles_x * tiles_y));
    int y0 = (tr / tiles_x) * TH, x0 = (tr % tiles_x) * TW;
    int64_t img = (int64_t)n * H * W;
    for (int e = threadIdx.x; e < HH * HWd * WG_CI; e += WG_THREADS) {
      int c = e % WG_CI;
      int hp = e / WG_CI;
      int hx = hp % HWd, hy = hp / HWd;
      int gy = y0 + hy - R, gx = x0 + hx - R;
      float v = 0.f;
      if (ci0 + c < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = ldf(x + (img + (int64_t)gy * W + gx) * ldx + ci0 + c);
      x_s[hp * XP + c] = v;
    }
    for (int e = threadIdx.x; e < TH * TW * WG_CO; e += WG_THREADS) {
      int c = e % WG_CO;
      int p = e / WG_CO;
      int gy = y0 + p / TW, gx = x0 + p % TW;
      float v = 0.f;
      if (co0 + c < Cout && gy < H && gx < W) v = ldf(dy + (img + (int64_t)gy * W + gx) * ldy + co0 + c);
      dy_s[p * WG_CO + c] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int p = 0; p < TH * TW; ++p) {
      float g0 = dy_s[p * WG_CO + cot], g1 = dy_s[p * WG_CO + cot + 16];
      bsum0 += g0;
      bsum1 += g1;
      const float* xp = x_s + ((p / TW) * HWd + (p % TW)) * XP + ci;
#pragma unroll
      for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          float xv = xp[(ky * HWd + kx) * XP];
          acc0[ky * K + kx] = fmaf(g0, xv, acc0[ky * K + kx]);
          acc1[ky * K + kx] = fmaf(g1, xv, acc1[ky * K + kx]);
        }
    }
    __syncthreads();
  }
  const int gci = ci0 + ci;
  if (gci < Cin) {
    int co_a = co0 + cot, co_b = co0 + cot + 16;
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
      if (co_a < Cout) atomicAdd(dw + ((int64_t)co_a * Cin + gci) * (K * K) + t, scale * acc0[t]);
      if (co_b < Cout) atomicAdd(dw + ((int64_t)co_b * Cin + gci) * (K * K) + t, scale * acc1[t]);
    }
  }
  if (db && blockIdx.x == 0 && ci == 0) {
    if (co0 + cot < Cout) atomicAdd(db + co0 + cot, scale * bsum0);
    if (co0 + cot + 16 < Cout) atomicAdd(db + co0 + cot + 16, scale * bsum1);
  }
}

template <typename T, int K>
int launch_wgrad(const void* x, int64_t ldx, const void* dy, int64_t ldy, float* dw, float* db, int N, int H,
                 int W, int Cin, int Cout, float scale, cudaStream_t s) {
  constexpr int HH = TH + K - 1, HWd = TW + K - 1;
  size_t smem = (size_t)(HH * HWd * (WG_CI + 1) + TH * TW * WG_CO) * sizeof(float);
  auto kern = conv_wgrad_kernel<T, K>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int gx = (int)cdiv(Cin, WG_CI), gy = (int)cdiv(Cout, WG_CO);
  int64_t ntiles = (int64_t)N * cdiv(H, TH) * cdiv(W, TW);
  int64_t want = cdiv((int64_t)kSMs * 4, (int64_t)gx * gy);
  int gz = (int)imax(1, imin(ntiles, want));
  dim3 grid(gx, gy, gz);
  kern<<<grid, WG_THREADS, smem, s>>>((const T*)x, ldx, (const T*)dy, ldy, dw, db, N, H, W, Cin, Cout, scale);
  return launch_status();
}

}  // namespace

namespace nv {

int conv_simt_fwd(const nervecl_conv_params& a, cudaStream_t s) {
  if (a.dtype == NERVECL_F32 && a.out_dtype == NERVECL_F32) return launch_fwd_k<float, float>(a, s);
  if (a.dtype == NERVECL_BF16 && a.out_dtype == NERVECL_BF16) return launch_fwd_k<bf16, bf16>(a, s);
  if (a.dtype == NERVECL_BF16 && a.out_dtype == NERVECL_F32) return launch_fwd_k<bf16, float>(a, s);
  return NERVECL_EDTYPE;
}

int conv_simt_wgrad(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype, float* dw, float* db,
                    int N, int H, int W, int Cin, int Cout, int K, float scale, cudaStream_t s) {
#define WG(T, KK) return launch_wgrad<T, KK>(x, ldx, dy, ldy, dw, db, N, H, W, Cin, Cout, scale, s)
  if (dtype == NERVECL_F32) {
    if (K == 1) WG(float, 1);
    if (K == 3) WG(float, 3);
    if (K == 7) WG(float, 7);
  } else if (dtype == NERVECL_BF16) {
    if (K == 1) WG(bf16, 1);
    if (K == 3) WG(bf16, 3);
    if (K == 7) WG(bf16, 7);
  } else {
    return NERVECL_EDTYPE;
  }
#undef WG
  return NERVECL_EUNSUPPORTED;
}

}  // namespace nv
