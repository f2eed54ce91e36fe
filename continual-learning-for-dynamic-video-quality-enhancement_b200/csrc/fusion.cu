// Temporal fusion (softmax over T + weighted sum), CBAM, and the output stage
// (PixelShuffle + bicubic skip + clamp).  Bandwidth-bound; fp32 math; channel-vector accesses.
#include "common.cuh"

using namespace nv;

namespace {

constexpr int TMAX = 8;

__device__ __forceinline__ float sigmoidf(float z) { return 1.f / (1.f + expf(-z)); }

__device__ __forceinline__ float group_sum(float v, int cg) {
  for (int o = cg >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float group_max(float v, int cg) {
  for (int o = cg >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int group_min_i(int v, int cg) {
  for (int o = cg >> 1; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// (pixel, 8-channel vector) of a flat thread index; 32-bit arithmetic when the index space allows it (a 64-bit
// division costs more instructions than the rest of these kernels' loop bodies)
__device__ __forceinline__ void split_item(int64_t t, int cg, bool small, int64_t& p, int& c0) {
  if (small) {
    const uint32_t tt = (uint32_t)t, pp = tt / (uint32_t)cg;
    p = pp;
    c0 = (int)(tt - pp * (uint32_t)cg) << 3;
  } else {
    p = t / cg;
    c0 = (int)(t - p * cg) << 3;
  }
}

// thread = (pixel, 8 channels); the cg = C/8 threads of a pixel are adjacent lanes.  NT (frames) is a template parameter:
// with a run-time frame count the per-frame arrays lived in local memory and the frame loads were issued one at a time.
template <typename T, int NT>
__global__ void __launch_bounds__(256)
tfuse_fwd_kernel(const T* __restrict__ feats, int64_t ldf_, const float* __restrict__ logits,
                 float* __restrict__ attn, T* __restrict__ out, int64_t ldo, int64_t npix, int C) {
  const int cg = C >> 3;
  const int64_t total = npix * cg;
  const bool small = total < ((int64_t)1 << 31);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int c0;
    int64_t p;
    split_item(t, cg, small, p, c0);
    const T* fp = feats + p * ldf_ + c0;
    f8 v[NT];
    float a[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) v[i] = ld8(fp + i * C);            // every load of the item in flight before the softmax
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NT; ++i) { a[i] = __ldg(logits + p * NT + i); m = fmaxf(m, a[i]); }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NT; ++i) { a[i] = expf(a[i] - m); s += a[i]; }
    const float inv = 1.f / s;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      a[i] *= inv;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(a[i], v[i].v[k], acc[k]);
    }
    if (c0 == 0) {
#pragma unroll
      for (int i = 0; i < NT; ++i) attn[p * NT + i] = a[i];
    }
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = acc[k];
    st8(out + p * ldo + c0, o);
  }
}

template <typename T, int NT>
__global__ void __launch_bounds__(256)
tfuse_bwd_kernel(const T* __restrict__ feats, int64_t ldf_, const float* __restrict__ attn,
                 const T* __restrict__ dout, int64_t lddo, const float* __restrict__ nc_bias, int64_t pix_per_image,
                 T* __restrict__ dfeats, int64_t lddf, float* __restrict__ dlogits, int64_t npix, int C) {
  const int cg = C >> 3;
  const int64_t total = npix * cg;
  const int64_t total_pad = cdiv(total, 32) * 32;
  const bool small = total_pad < ((int64_t)1 << 31);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total_pad;
       t += (int64_t)gridDim.x * blockDim.x) {
    const bool active = t < total;
    int c0;
    int64_t p;
    split_item(t, cg, small, p, c0);
    float da[NT], a[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) da[i] = a[i] = 0.f;
    if (active) {
      const T* fp = feats + p * ldf_ + c0;
      f8 v[NT];
#pragma unroll
      for (int i = 0; i < NT; ++i) v[i] = ld8(fp + i * C);
      f8 d = ld8(dout + p * lddo + c0);
#pragma unroll
      for (int i = 0; i < NT; ++i) a[i] = __ldg(attn + p * NT + i);
      if (nc_bias) {
        const float* b = nc_bias + (p / pix_per_image) * C + c0;
#pragma unroll
        for (int k = 0; k < 8; ++k) d.v[k] += __ldg(b + k);
      }
      T* dp = dfeats + p * lddf + c0;
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        f8 o;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          o.v[k] = a[i] * d.v[k];
          s = fmaf(d.v[k], v[i].v[k], s);
        }
        da[i] = s;
        st8(dp + i * C, o);
      }
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) da[i] = group_sum(da[i], cg);
    if (active && c0 == 0) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < NT; ++i) dot = fmaf(a[i], da[i], dot);
#pragma unroll
      for (int i = 0; i < NT; ++i) dlogits[p * NT + i] = a[i] * (da[i] - dot);
    }
  }
}

// dispatch on the frame count (1 .. TMAX)
#define NV_DISPATCH_FRAMES(T_, NT, ...)                                              \
  switch (T_) {                                                                      \
    case 1: { constexpr int NT = 1; __VA_ARGS__; } break;                            \
    case 2: { constexpr int NT = 2; __VA_ARGS__; } break;                            \
    case 3: { constexpr int NT = 3; __VA_ARGS__; } break;                            \
    case 4: { constexpr int NT = 4; __VA_ARGS__; } break;                            \
    case 5: { constexpr int NT = 5; __VA_ARGS__; } break;                            \
    case 6: { constexpr int NT = 6; __VA_ARGS__; } break;                            \
    case 7: { constexpr int NT = 7; __VA_ARGS__; } break;                            \
    default: { constexpr int NT = 8; __VA_ARGS__; } break;                           \
  }

// out[n][c] += scale * sum_p x[n,p,c];  grid = (chunks, N), block = (C/4) x lanes
template <typename T>
__global__ void __launch_bounds__(256)
chan_sum_kernel(const T* __restrict__ x, int64_t ldx, int64_t pix, int C, float scale, float* __restrict__ out) {
  const int cg = C >> 2;
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int n = blockIdx.y;
  const T* xb = x + (int64_t)n * pix * ldx;
  double s[4] = {0, 0, 0, 0};
  if (lane < lanes) {
    float fs[4] = {0, 0, 0, 0};
    int cnt = 0;
    for (int64_t p = (int64_t)blockIdx.x * lanes + lane; p < pix; p += (int64_t)gridDim.x * lanes) {
      f4 v = ld4(xb + p * ldx + (g << 2));
#pragma unroll
      for (int k = 0; k < 4; ++k) fs[k] += v.v[k];
      if (++cnt == 64) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { s[k] += fs[k]; fs[k] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] += fs[k];
  }
  extern __shared__ double dred[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) dred[i] = 0.0;
  __syncthreads();
  if (lane < lanes) {
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(&dred[(g << 2) + k], s[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + (int64_t)n * C + i, (float)(dred[i] * (double)scale));
}

// one block per image
__global__ void ca_gate_fwd_kernel(const float* __restrict__ pool, const float* __restrict__ w1,
                                   const float* __restrict__ w2, float* __restrict__ hidden,
                                   float* __restrict__ gate, int C, int R) {
  extern __shared__ float sh[];  // [R]
  const int n = blockIdx.x;
  const float* pl = pool + (int64_t)n * C;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < R; r += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(pl[c], w1[(int64_t)r * C + c], s);
    s = warp_sum(s);
    if (lane == 0) { s = fmaxf(s, 0.f); sh[r] = s; hidden[(int64_t)n * R + r] = s; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s = fmaf(sh[r], w2[(int64_t)c * R + r], s);
    gate[(int64_t)n * C + c] = sigmoidf(s);
  }
}

__global__ void ca_gate_bwd_kernel(const float* __restrict__ pool, const float* __restrict__ w1,
                                   const float* __restrict__ w2, const float* __restrict__ hidden,
                                   const float* __restrict__ gate, const float* __restrict__ dgate,
                                   float* __restrict__ dpool, float* __restrict__ dw1, float* __restrict__ dw2,
                                   int C, int R) {
  extern __shared__ float sh[];  // ds[C], dh[R]
  float* ds = sh;
  float* dh = sh + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float gt = gate[(int64_t)n * C + c];
    ds[c] = dgate[(int64_t)n * C + c] * gt * (1.f - gt);
  }
  __syncthreads();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < R; r += nw) {
    float h = hidden[(int64_t)n * R + r];
    float s = 0.f;
    for (int c = lane; c < C; c += 32) {
      s = fmaf(ds[c], w2[(int64_t)c * R + r], s);
      atomicAdd(dw2 + (int64_t)c * R + r, ds[c] * h);
    }
    s = warp_sum(s);
    if (lane == 0) dh[r] = h > 0.f ? s : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float pc = pool[(int64_t)n * C + c];
    float s = 0.f;
    for (int r = 0; r < R; ++r) {
      s = fmaf(dh[r], w1[(int64_t)r * C + c], s);
      atomicAdd(dw1 + (int64_t)r * C + c, dh[r] * pc);
    }
    dpool[(int64_t)n * C + c] = s;
  }
}

// grid = (chunks, N): the image -- and with it the thread's eight channel-gate values -- is fixed per block, so the gates
// sit in registers (the flat-index form re-derived the image with a 64-bit division and re-loaded eight gate values per
// item: ncu 82 % issue-slot utilisation at 31 % of the DRAM bandwidth).  Same arithmetic, same order.
template <typename T>
__global__ void __launch_bounds__(256)
cbam_stats_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ gate, float* __restrict__ stats,
                  int64_t pix, int C) {
  const int cg = C >> 3, sh = __ffs(cg) - 1;         // power of two <= 32 (host: group_ok)
  const int n = blockIdx.y;
  const int c0 = (threadIdx.x & (cg - 1)) << 3;      // fixed per thread: block size and stride are multiples of cg
  float gt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) gt[k] = __ldg(gate + (int64_t)n * C + c0 + k);
  const T* xb = x + (int64_t)n * pix * ldx + c0;
  float2* sb = reinterpret_cast<float2*>(stats) + (int64_t)n * pix;
  const int64_t total = pix << sh;
  const int64_t total_pad = cdiv(total, 32) * 32;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total_pad;
       t += (int64_t)gridDim.x * blockDim.x) {
    const bool active = t < total;
    const int64_t p = t >> sh;
    float s = 0.f, m = -INFINITY;
    if (active) {
      const f8 v = ld8(xb + p * ldx);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xs = v.v[k] * gt[k];
        s += xs;
        m = fmaxf(m, xs);
      }
    }
    s = group_sum(s, cg);
    m = group_max(m, cg);
    if (active && c0 == 0) sb[p] = make_float2(s / (float)C, m);
  }
}

// Spatial attention + output, one block per 8 x 32 pixel tile.  (The first version spread a pixel's 49 taps over its C/8
// lanes -- 64-bit index arithmetic and scattered float2 loads per lane: 0.69 ms at the cfg-2 shape against 0.15 ms of
// HBM time.)  Phase A: the tile's (8+6) x (32+6) statistics halo goes to shared memory (zero outside the image = the
// conv's zero padding), one thread per pixel runs the 7x7 conv over it and keeps the sigmoid in shared memory; phase B:
// all threads stream the tile's channels as 16-byte vectors, four consecutive pixels per 8-lane group row.
constexpr int CB_TH = 8, CB_TW = 32;
template <typename T>
__global__ void __launch_bounds__(256)
cbam_apply_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ gate,
                  const float* __restrict__ stats, const float* __restrict__ w7, float* __restrict__ sgate,
                  T* __restrict__ out, int64_t ldo, int H, int W, int C) {
  constexpr int HH = CB_TH + 6, HW = CB_TW + 6;
  __shared__ float2 st_s[HH][HW + 1];
  __shared__ float ws[98];
  __shared__ float sg_s[CB_TH * CB_TW];
  const int n = blockIdx.z, y0 = blockIdx.y * CB_TH, x0 = blockIdx.x * CB_TW;
  const int64_t img = (int64_t)n * H * W;
  for (int i = threadIdx.x; i < 98; i += 256) ws[i] = w7[i];
  for (int e = threadIdx.x; e < HH * HW; e += 256) {
    const int hy = e / HW, hx = e - hy * HW;
    const int gy = y0 + hy - 3, gx = x0 + hx - 3;
    float2 v = make_float2(0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(reinterpret_cast<const float2*>(stats) + img + (int64_t)gy * W + gx);
    st_s[hy][hx] = v;
  }
  __syncthreads();
  {
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    float z = 0.f;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float2 v = st_s[ty + ky][tx + kx];
        z = fmaf(ws[ky * 7 + kx], v.x, z);
        z = fmaf(ws[49 + ky * 7 + kx], v.y, z);
      }
    const float sg = sigmoidf(z);
    sg_s[threadIdx.x] = sg;
    if (y0 + ty < H && x0 + tx < W) sgate[img + (int64_t)(y0 + ty) * W + x0 + tx] = sg;
  }
  __syncthreads();
  const int cg = C >> 3;                             // power of two <= 32 (host: group_ok)
  const int sh = __ffs(cg) - 1;
  const int c0 = (threadIdx.x & (cg - 1)) << 3;      // fixed per thread: 256 is a multiple of cg
  float gt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) gt[k] = __ldg(gate + (int64_t)n * C + c0 + k);
  const int items = CB_TH * CB_TW * cg;
#pragma unroll 2
  for (int it = threadIdx.x; it < items; it += 256) {
    const int pl = it >> sh;
    const int ty = pl >> 5, tx = pl & 31;
    if (y0 + ty >= H || x0 + tx >= W) continue;
    const int64_t p = img + (int64_t)(y0 + ty) * W + x0 + tx;
    const float sg = sg_s[pl];
    const f8 v = ld8(x + p * ldx + c0);
    f8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = v.v[k] * gt[k] * sg;
    st8(out + p * ldo + c0, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
cbam_bwd_dz_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ gate,
                   const float* __restrict__ sgate, const T* __restrict__ dy, int64_t lddy, float* __restrict__ dz,
                   int64_t pix, int C) {
  const int cg = C >> 3, sh = __ffs(cg) - 1;         // (grid = (chunks, N), gates in registers: see cbam_stats_kernel)
  const int n = blockIdx.y;
  const int c0 = (threadIdx.x & (cg - 1)) << 3;
  float gt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) gt[k] = __ldg(gate + (int64_t)n * C + c0 + k);
  const int64_t img = (int64_t)n * pix;
  const T* xb = x + img * ldx + c0;
  const T* db = dy + img * lddy + c0;
  const int64_t total = pix << sh;
  const int64_t total_pad = cdiv(total, 32) * 32;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total_pad;
       t += (int64_t)gridDim.x * blockDim.x) {
    const bool active = t < total;
    const int64_t p = t >> sh;
    float s = 0.f;
    if (active) {
      const f8 v = ld8(xb + p * ldx);
      const f8 d = ld8(db + p * lddy);
#pragma unroll
      for (int k = 0; k < 8; ++k) s = fmaf(d.v[k], v.v[k] * gt[k], s);
    }
    s = group_sum(s, cg);
    if (active && c0 == 0) {
      const float sg = sgate[img + p];
      dz[img + p] = s * sg * (1.f - sg);
    }
  }
}

// 16x16 pixel tiles, blocks loop over them.  dstats = transposed 7x7 conv of dz;  dw7 += sum dz * shifted stats.
//
// dw7[ch][ky][kx] = sum_{ty,tx} dz[ty][tx] * stats[ch][ty + ky][tx + kx]: thread = (ch, ky, tile row ty); its seven kx
// outputs slide over ONE 22-value statistics row (1 shared-memory load per 7 multiply-adds) and stay in registers across
// ALL the block's tiles; one shuffle reduction over the 16 rows and 98 coalesced atomics per BLOCK at the end.  (The first
// version gave each of the 98 outputs a thread that walked the whole tile -- 512 loads each -- and issued 98 atomics per
// tile: 1.4 M atomic operations on four cache lines at the cfg-2 shape.)
__global__ void __launch_bounds__(256)
cbam_bwd_spatial_kernel(const float* __restrict__ dz, const float* __restrict__ stats, const float* __restrict__ w7,
                        float* __restrict__ dstats, float* __restrict__ dw7, int H, int W, int tiles_x, int tiles_y,
                        int ntiles) {
  constexpr int TS = 16, HS = TS + 6;
  __shared__ float dz_s[HS][HS];
  __shared__ float st_s[2][HS][HS];
  __shared__ float ws[98];
  __shared__ float red[98];
  for (int i = threadIdx.x; i < 98; i += blockDim.x) ws[i] = w7[i];
  const int rty = threadIdx.x & (TS - 1), gq = threadIdx.x >> 4;        // weight-gradient role: gq = ch * 7 + ky (< 14)
  const int rch = gq / 7, rky = gq - rch * 7;
  float acc[7];
#pragma unroll
  for (int kx = 0; kx < 7; ++kx) acc[kx] = 0.f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int bx = tile % tiles_x, r = tile / tiles_x;
    const int by = r % tiles_y, n = r / tiles_y;
    const int y0 = by * TS, x0 = bx * TS;
    const int64_t img = (int64_t)n * H * W;
    __syncthreads();                                  // the previous tile's readers are done (and ws is visible)
    for (int e = threadIdx.x; e < HS * HS; e += blockDim.x) {
      int hy = e / HS, hx = e % HS;
      int gy = y0 + hy - 3, gx = x0 + hx - 3;
      float d = 0.f, s0 = 0.f, s1 = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        int64_t q = img + (int64_t)gy * W + gx;
        d = dz[q];
        float2 st = reinterpret_cast<const float2*>(stats)[q];
        s0 = st.x;
        s1 = st.y;
      }
      dz_s[hy][hx] = d;
      st_s[0][hy][hx] = s0;
      st_s[1][hy][hx] = s1;
    }
    __syncthreads();
    {
      int ty = threadIdx.x / TS, tx = threadIdx.x % TS;
      int gy = y0 + ty, gx = x0 + tx;
      if (gy < H && gx < W) {
        // dstats[q][ch] = sum_k w[ch][k] * dz[q - (k - 3)]
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky)
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
            float d = dz_s[ty + 3 - (ky - 3)][tx + 3 - (kx - 3)];
            a0 = fmaf(ws[ky * 7 + kx], d, a0);
            a1 = fmaf(ws[49 + ky * 7 + kx], d, a1);
          }
        reinterpret_cast<float2*>(dstats)[img + (int64_t)gy * W + gx] = make_float2(a0, a1);
      }
    }
    if (gq < 14) {
      // pixels of the tile outside the image: dz_s is zero there (only the INTERIOR [3, 3 + TS) is read as dz)
      const bool row_in = y0 + rty < H;
#pragma unroll
      for (int tx = 0; tx < TS; ++tx) {
        const float d = (row_in && x0 + tx < W) ? dz_s[rty + 3][tx + 3] : 0.f;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) acc[kx] = fmaf(d, st_s[rch][rty + rky][tx + kx], acc[kx]);
      }
    }
  }
  if (gq < 14) {
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
#pragma unroll
      for (int o = TS / 2; o > 0; o >>= 1) acc[kx] += __shfl_xor_sync(0xffffffffu, acc[kx], o);
      if (rty == 0) red[gq * 7 + kx] = acc[kx];
    }
  }
  __syncthreads();
  if (threadIdx.x < 98) atomicAdd(dw7 + threadIdx.x, red[threadIdx.x]);
}

// grid = (chunks, N)
template <typename T>
__global__ void __launch_bounds__(256)
cbam_bwd_dx_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ gate,
                   const float* __restrict__ sgate, const float* __restrict__ stats,
                   const float* __restrict__ dstats, const T* __restrict__ dy, int64_t lddy, T* __restrict__ dx,
                   int64_t lddx, float* __restrict__ dgate, int64_t pix, int C) {
  const int cg = C >> 3;
  const int n = blockIdx.y;
  const int64_t total = pix * cg;
  const int64_t total_pad = cdiv(total, 32) * 32;
  const float inv_c = 1.f / (float)C;
  float gacc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) gacc[k] = 0.f;
  const int c0 = (int)(threadIdx.x % cg) << 3;  // fixed per thread: blockDim and stride are multiples of cg
  float gt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) gt[k] = gate[(int64_t)n * C + c0 + k];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total_pad;
       t += (int64_t)gridDim.x * blockDim.x) {
    const bool active = t < total;
    int64_t p = (int64_t)n * pix + t / cg;
    int first = 1 << 30;
    f8 v, d;
    float2 st = make_float2(0.f, 0.f), dst = make_float2(0.f, 0.f);
    float sg = 0.f;
    if (active) {
      v = ld8(x + p * ldx + c0);
      d = ld8(dy + p * lddy + c0);
      st = reinterpret_cast<const float2*>(stats)[p];
      dst = reinterpret_cast<const float2*>(dstats)[p];
      sg = sgate[p];
#pragma unroll
      for (int k = 7; k >= 0; --k)
        if (v.v[k] * gt[k] == st.y) first = c0 + k;
    }
    first = group_min_i(first, cg);
    if (active) {
      f8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float dxs = d.v[k] * sg + dst.x * inv_c + ((c0 + k) == first ? dst.y : 0.f);
        o.v[k] = dxs * gt[k];
        gacc[k] = fmaf(dxs, v.v[k], gacc[k]);
      }
      st8(dx + p * lddx + c0, o);
    }
  }
  extern __shared__ float red[];  // [C]
  for (int i = threadIdx.x; i < C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) atomicAdd(&red[c0 + k], gacc[k]);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dgate + (int64_t)n * C + i, red[i]);
}

// ---------------------------------------------------------------------------------------
// output stage
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float cubic1(float x) {  // |x| <= 1, A = -0.75
  const float A = -0.75f;
  return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
}
__device__ __forceinline__ float cubic2(float x) {  // 1 < |x| < 2
  const float A = -0.75f;
  return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

// ATen upsample_bicubic2d, align_corners=False, scale_factor given (UpSample.h:289-311,413)
__device__ __forceinline__ float bicubic_sample(const float* __restrict__ img, int64_t sH, int H, int W, int Y,
                                                int X, float rscale) {
  float ry = rscale * ((float)Y + 0.5f) - 0.5f;
  float rx = rscale * ((float)X + 0.5f) - 0.5f;
  float fy = floorf(ry), fx = floorf(rx);
  int iy = (int)fy, ix = (int)fx;
  float ty = ry - fy, tx = rx - fx;
  float cx[4] = {cubic2(tx + 1.f), cubic1(tx), cubic1(1.f - tx), cubic2(2.f - tx)};
  float cy[4] = {cubic2(ty + 1.f), cubic1(ty), cubic1(1.f - ty), cubic2(2.f - ty)};
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int yy = min(max(iy - 1 + i, 0), H - 1);
    const float* row = img + (int64_t)yy * sH;
    float r = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int xx = min(max(ix - 1 + j, 0), W - 1);
      r += __ldg(row + xx) * cx[j];
    }
    acc += r * cy[i];
  }
  return acc;
}

__global__ void __launch_bounds__(256)
upfinish_fwd_kernel(const float* __restrict__ conv, const float* __restrict__ lr, int64_t sN, int64_t sC,
                    int64_t sH, float* __restrict__ out, int N, int C, int H, int W, int s, float rscale) {
  const int HO = H * s, WO = W * s, CS = C * s * s;
  const int64_t total = (int64_t)N * C * HO * WO;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int X = (int)(i % WO);
    int64_t r = i / WO;
    int Y = (int)(r % HO);
    r /= HO;
    int c = (int)(r % C);
    int n = (int)(r / C);
    int y = Y / s, ii = Y % s, x = X / s, jj = X % s;
    float v = conv[(((int64_t)n * H + y) * W + x) * CS + c * s * s + ii * s + jj];
    v += bicubic_sample(lr + n * sN + c * sC, sH, H, W, Y, X, rscale);
    out[i] = fminf(fmaxf(v, 0.f), 1.f);
  }
}

__global__ void __launch_bounds__(256)
upfinish_bwd_kernel(const float* __restrict__ conv, const float* __restrict__ lr, int64_t sN, int64_t sC,
                    int64_t sH, const float* __restrict__ dout, float* __restrict__ dconv, int N, int C, int H,
                    int W, int s, float rscale) {
  const int HO = H * s, WO = W * s, CS = C * s * s;
  const int64_t total = (int64_t)N * C * HO * WO;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int X = (int)(i % WO);
    int64_t r = i / WO;
    int Y = (int)(r % HO);
    r /= HO;
    int c = (int)(r % C);
    int n = (int)(r / C);
    int y = Y / s, ii = Y % s, x = X / s, jj = X % s;
    int64_t ci = (((int64_t)n * H + y) * W + x) * CS + c * s * s + ii * s + jj;
    float v = conv[ci] + bicubic_sample(lr + n * sN + c * sC, sH, H, W, Y, X, rscale);
    dconv[ci] = (v >= 0.f && v <= 1.f) ? dout[i] : 0.f;
  }
}

// One thread per LR pixel for integer scales 2 / 3 / 4 (the element-per-thread kernels above were instruction bound at
// ~200 instructions per output: 64-bit index arithmetic and the cubic coefficients recomputed for every channel).  The S
// row phases and S column phases of the pixel's S x S outputs share their coefficient sets and clamped tap indices across
// the phases and channels; every output keeps bicubic_sample's arithmetic (same taps, same order).  The thread's C * S * S
// conv values are one contiguous record; its outputs are S adjacent floats per channel and row.
template <int S, bool BWD>
__global__ void __launch_bounds__(128, S == 2 ? 8 : 4)
upfinish_px_kernel(const float* __restrict__ conv, const float* __restrict__ lr, int64_t sN, int64_t sC, int64_t sH,
                   float* __restrict__ out, const float* __restrict__ dout, float* __restrict__ dconv, int N, int C, int H,
                   int W, float rscale) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;               // (N * H * W < 2^31: checked by the launcher)
  if (p >= N * H * W) return;
  const int x = p % W, r = p / W;
  const int y = r % H, n = r / H;
  // column phases: coefficient sets and clamped tap columns of the S outputs of a row
  float cx[S][4];
  int xo[S][4];
#pragma unroll
  for (int q = 0; q < S; ++q) {
    const float rx = rscale * ((float)(x * S + q) + 0.5f) - 0.5f;
    const float fx = floorf(rx);
    const int ix = (int)fx;
    const float tx = rx - fx;
    cx[q][0] = cubic2(tx + 1.f); cx[q][1] = cubic1(tx); cx[q][2] = cubic1(1.f - tx); cx[q][3] = cubic2(2.f - tx);
#pragma unroll
    for (int j = 0; j < 4; ++j) xo[q][j] = min(max(ix - 1 + j, 0), W - 1);
  }
  const int HO = H * S, WO = W * S;
  const float* rec = conv + (int64_t)p * (C * S * S);
  float* drec = BWD ? dconv + (int64_t)p * (C * S * S) : nullptr;
  const int ib0 = n * (int)sN;
#pragma unroll 1
  for (int qy = 0; qy < S; ++qy) {
    // row phase
    const float ry = rscale * ((float)(y * S + qy) + 0.5f) - 0.5f;
    const float fy = floorf(ry);
    const int iy = (int)fy;
    const float ty = ry - fy;
    const float cy[4] = {cubic2(ty + 1.f), cubic1(ty), cubic1(1.f - ty), cubic2(2.f - ty)};
    int yo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) yo[i] = min(max(iy - 1 + i, 0), H - 1) * (int)sH;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
      const int ib = ib0 + c * (int)sC;
      const int64_t obase = (((int64_t)n * C + c) * HO + (int64_t)(y * S + qy)) * WO + (int64_t)x * S;
      // the S conv values of this (channel, row phase): one 8 / 16-byte load for S = 2 / 4 (the record is 16-byte aligned)
      float v[S];
      const float* rp = rec + c * S * S + qy * S;
      if (S == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(rp));
        v[0] = t.x; v[1 % S] = t.y;
      } else if (S == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(rp));
        v[0] = t.x; v[1 % S] = t.y; v[2 % S] = t.z; v[3 % S] = t.w;
      } else {
#pragma unroll
        for (int qx = 0; qx < S; ++qx) v[qx] = __ldg(rp + qx);
      }
#pragma unroll
      for (int qx = 0; qx < S; ++qx) {
        float acc = 0.f;                                             // (bicubic_sample's taps, in its order)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ro = ib + yo[i];                                 // (32-bit offsets: checked by the launcher)
          float rr = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) rr += __ldg(lr + (ro + xo[qx][j])) * cx[qx][j];
          acc += rr * cy[i];
        }
        v[qx] += acc;
      }
      if (BWD) {
        float g[S];
        const float* gp = dout + obase;
        if (S == 2) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(gp));
          g[0] = t.x; g[1 % S] = t.y;
        } else if (S == 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(gp));
          g[0] = t.x; g[1 % S] = t.y; g[2 % S] = t.z; g[3 % S] = t.w;
        } else {
#pragma unroll
          for (int qx = 0; qx < S; ++qx) g[qx] = __ldg(gp + qx);
        }
#pragma unroll
        for (int qx = 0; qx < S; ++qx) g[qx] = (v[qx] >= 0.f && v[qx] <= 1.f) ? g[qx] : 0.f;
        float* dp = drec + c * S * S + qy * S;
        if (S == 2) *reinterpret_cast<float2*>(dp) = make_float2(g[0], g[1 % S]);
        else if (S == 4) *reinterpret_cast<float4*>(dp) = make_float4(g[0], g[1 % S], g[2 % S], g[3 % S]);
        else {
#pragma unroll
          for (int qx = 0; qx < S; ++qx) dp[qx] = g[qx];
        }
      } else {
        float* o = out + obase;
#pragma unroll
        for (int qx = 0; qx < S; ++qx) v[qx] = fminf(fmaxf(v[qx], 0.f), 1.f);
        if (S == 2) *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1 % S]);
        else if (S == 4) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1 % S], v[2 % S], v[3 % S]);
        else {
#pragma unroll
          for (int qx = 0; qx < S; ++qx) o[qx] = v[qx];
        }
      }
      asm volatile("" ::: "memory");                 // keep the next channel's 16 S tap loads behind this channel's stores
    }
  }
}

// true when the per-pixel kernels take the shape (else the element-per-thread kernels run)
inline bool upfinish_px_ok(const void* conv, const void* out_or_dout, const void* dconv, int64_t sN, int64_t sC, int64_t sH,
                           int N, int C, int H, int W, int s) {
  const int64_t lim = (int64_t)1 << 31;              // every element offset into lr fits 32 bits
  return (s == 2 || s == 3 || s == 4) && (int64_t)N * H * W < lim && sN >= 0 && sC >= 0 && sH >= 0 &&
         (N - 1) * sN + (C - 1) * sC + (int64_t)H * sH < lim &&
         aligned(conv, 16) && aligned(out_or_dout, 16) && (!dconv || aligned(dconv, 16));
}

// EnhancementEngine strength blend (enhancement_engine.py:172-182): out = strength * out + (1 - strength) * bicubic(lr)
__global__ void __launch_bounds__(256)
bicubic_blend_kernel(float* __restrict__ out, const float* __restrict__ lr, int64_t sN, int64_t sC, int64_t sH, int N, int C,
                     int H, int W, int s, float rscale, float strength) {
  const int HO = H * s, WO = W * s;
  const int64_t total = (int64_t)N * C * HO * WO;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int X = (int)(i % WO);
    int64_t r = i / WO;
    int Y = (int)(r % HO);
    r /= HO;
    int c = (int)(r % C);
    int n = (int)(r / C);
    const float b = bicubic_sample(lr + n * sN + c * sC, sH, H, W, Y, X, rscale);
    out[i] = strength * out[i] + (1.f - strength) * b;
  }
}

inline int ew_blocks(int64_t work) { return (int)imax(1, imin(cdiv(work, 256), kSMs * 16)); }
inline bool group_ok(int C) {
  int cg = C >> 3;
  return !(C & 7) && cg >= 1 && cg <= 32 && !(cg & (cg - 1));
}

}  // namespace

NV_API int nervecl_tfuse_fwd(const void* feats, int64_t ldf, const float* logits, float* attn, void* out,
                             int64_t ldo, int dtype, int64_t npix, int T, int C, nervecl_stream_t stream) {
  if (!feats || !logits || !attn || !out || npix <= 0 || T <= 0 || T > TMAX || C <= 0) return NERVECL_EINVAL;
  if ((C & 7) || (ldf & 7) || (ldo & 7)) return NERVECL_EALIGN;
  NV_DISPATCH_DTYPE(dtype, E, NV_DISPATCH_FRAMES(T, NT, (tfuse_fwd_kernel<E, NT><<<ew_blocks(npix * (C >> 3)), 256, 0, as_stream(stream)>>>(
                                  (const E*)feats, ldf, logits, attn, (E*)out, ldo, npix, C))));
  return launch_status();
}

NV_API int nervecl_tfuse_bwd(const void* feats, int64_t ldf, const float* attn, const void* dout, int64_t lddo,
                             const float* nc_bias, int64_t pix_per_image, void* dfeats, int64_t lddf,
                             float* dlogits, int dtype, int64_t npix, int T, int C, nervecl_stream_t stream) {
  if (!feats || !attn || !dout || !dfeats || !dlogits || npix <= 0 || T <= 0 || T > TMAX || C <= 0)
    return NERVECL_EINVAL;
  if (nc_bias && pix_per_image <= 0) return NERVECL_EINVAL;
  if (!group_ok(C)) return NERVECL_EUNSUPPORTED;
  if ((ldf & 7) || (lddo & 7) || (lddf & 7)) return NERVECL_EALIGN;
  NV_DISPATCH_DTYPE(dtype, E, NV_DISPATCH_FRAMES(T, NT, (tfuse_bwd_kernel<E, NT><<<ew_blocks(npix * (C >> 3)), 256, 0, as_stream(stream)>>>(
                                  (const E*)feats, ldf, attn, (const E*)dout, lddo, nc_bias, pix_per_image,
                                  (E*)dfeats, lddf, dlogits, npix, C))));
  return launch_status();
}

NV_API int nervecl_chan_sum(const void* x, int64_t ldx, int dtype, int N, int64_t pix_per_image, int C,
                            float scale, float* out, nervecl_stream_t stream) {
  if (!x || !out || N <= 0 || pix_per_image <= 0 || C <= 0) return NERVECL_EINVAL;
  if ((C & 3) || (ldx & 3) || C > 1024) return NERVECL_EALIGN;
  int lanes = 256 / (C >> 2);
  int chunks = (int)imax(1, imin(cdiv(pix_per_image, lanes * 16), (kSMs * 8) / N + 1));
  dim3 grid(chunks, N);
  NV_DISPATCH_DTYPE(dtype, E, (chan_sum_kernel<E><<<grid, 256, C * sizeof(double), as_stream(stream)>>>(
                                  (const E*)x, ldx, pix_per_image, C, scale, out)));
  return launch_status();
}

NV_API int nervecl_ca_gate_fwd(const float* pool, const float* w1, const float* w2, float* hidden, float* gate,
                               int N, int C, int R, nervecl_stream_t stream) {
  if (!pool || !w1 || !w2 || !hidden || !gate || N <= 0 || C <= 0 || R <= 0) return NERVECL_EINVAL;
  ca_gate_fwd_kernel<<<N, 128, R * sizeof(float), as_stream(stream)>>>(pool, w1, w2, hidden, gate, C, R);
  return launch_status();
}

NV_API int nervecl_ca_gate_bwd(const float* pool, const float* w1, const float* w2, const float* hidden,
                               const float* gate, const float* dgate, float* dpool, float* dw1, float* dw2, int N,
                               int C, int R, nervecl_stream_t stream) {
  if (!pool || !w1 || !w2 || !hidden || !gate || !dgate || !dpool || !dw1 || !dw2 || N <= 0 || C <= 0 || R <= 0)
    return NERVECL_EINVAL;
  ca_gate_bwd_kernel<<<N, 128, (C + R) * sizeof(float), as_stream(stream)>>>(pool, w1, w2, hidden, gate, dgate,
                                                                               dpool, dw1, dw2, C, R);
  return launch_status();
}

NV_API int nervecl_cbam_stats_fwd(const void* x, int64_t ldx, const float* gate, float* stats, int dtype, int N,
                                  int64_t pix_per_image, int C, nervecl_stream_t stream) {
  if (!x || !gate || !stats || N <= 0 || pix_per_image <= 0) return NERVECL_EINVAL;
  if (!group_ok(C)) return NERVECL_EUNSUPPORTED;
  if (ldx & 7) return NERVECL_EALIGN;
  if (N > 65535) return NERVECL_EUNSUPPORTED;
  const int chunks = (int)imax(1, imin(cdiv(pix_per_image * (C >> 3), 256 * 4), (kSMs * 16) / N + 1));
  dim3 grid(chunks, N);
  NV_DISPATCH_DTYPE(dtype, E, (cbam_stats_kernel<E><<<grid, 256, 0, as_stream(stream)>>>(
                                  (const E*)x, ldx, gate, stats, pix_per_image, C)));
  return launch_status();
}

NV_API int nervecl_cbam_apply_fwd(const void* x, int64_t ldx, const float* gate, const float* stats,
                                  const float* w7, float* sgate, void* out, int64_t ldo, int dtype, int N, int H,
                                  int W, int C, nervecl_stream_t stream) {
  if (!x || !gate || !stats || !w7 || !sgate || !out || N <= 0 || H <= 0 || W <= 0) return NERVECL_EINVAL;
  if (!group_ok(C)) return NERVECL_EUNSUPPORTED;
  if ((ldx & 7) || (ldo & 7)) return NERVECL_EALIGN;
  if (N > 65535 || cdiv(H, CB_TH) > 65535) return NERVECL_EUNSUPPORTED;
  dim3 grid((unsigned)cdiv(W, CB_TW), (unsigned)cdiv(H, CB_TH), (unsigned)N);
  NV_DISPATCH_DTYPE(dtype, E, (cbam_apply_kernel<E><<<grid, 256, 0, as_stream(stream)>>>(
                                  (const E*)x, ldx, gate, stats, w7, sgate, (E*)out, ldo, H, W, C)));
  return launch_status();
}

NV_API int nervecl_cbam_bwd_dz(const void* x, int64_t ldx, const float* gate, const float* sgate, const void* dy,
                               int64_t lddy, float* dz, int dtype, int N, int64_t pix_per_image, int C,
                               nervecl_stream_t stream) {
  if (!x || !gate || !sgate || !dy || !dz || N <= 0 || pix_per_image <= 0) return NERVECL_EINVAL;
  if (!group_ok(C)) return NERVECL_EUNSUPPORTED;
  if ((ldx & 7) || (lddy & 7)) return NERVECL_EALIGN;
  if (N > 65535) return NERVECL_EUNSUPPORTED;
  const int chunks = (int)imax(1, imin(cdiv(pix_per_image * (C >> 3), 256 * 4), (kSMs * 16) / N + 1));
  dim3 grid(chunks, N);
  NV_DISPATCH_DTYPE(dtype, E, (cbam_bwd_dz_kernel<E><<<grid, 256, 0, as_stream(stream)>>>(
                                  (const E*)x, ldx, gate, sgate, (const E*)dy, lddy, dz, pix_per_image, C)));
  return launch_status();
}

NV_API int nervecl_cbam_bwd_spatial(const float* dz, const float* stats, const float* w7, float* dstats,
                                    float* dw7, int N, int H, int W, nervecl_stream_t stream) {
  if (!dz || !stats || !w7 || !dstats || !dw7 || N <= 0 || H <= 0 || W <= 0) return NERVECL_EINVAL;
  const int64_t tiles_x = cdiv(W, 16), tiles_y = cdiv(H, 16), ntiles = tiles_x * tiles_y * N;
  if (ntiles >= ((int64_t)1 << 31)) return NERVECL_EUNSUPPORTED;
  const int blocks = (int)imin(ntiles, kSMs * 8);
  cbam_bwd_spatial_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dz, stats, w7, dstats, dw7, H, W, (int)tiles_x, (int)tiles_y,
                                                               (int)ntiles);
  return launch_status();
}

NV_API int nervecl_cbam_bwd_dx(const void* x, int64_t ldx, const float* gate, const float* sgate,
                               const float* stats, const float* dstats, const void* dy, int64_t lddy, void* dx,
                               int64_t lddx, float* dgate, int dtype, int N, int64_t pix_per_image, int C,
                               nervecl_stream_t stream) {
  if (!x || !gate || !sgate || !stats || !dstats || !dy || !dx || !dgate || N <= 0 || pix_per_image <= 0)
    return NERVECL_EINVAL;
  if (!group_ok(C)) return NERVECL_EUNSUPPORTED;
  if ((ldx & 7) || (lddy & 7) || (lddx & 7)) return NERVECL_EALIGN;
  int chunks = (int)imax(1, imin(cdiv(pix_per_image * (C >> 3), 256 * 4), (kSMs * 8) / N + 1));
  dim3 grid(chunks, N);
  NV_DISPATCH_DTYPE(dtype, E, (cbam_bwd_dx_kernel<E><<<grid, 256, C * sizeof(float), as_stream(stream)>>>(
                                  (const E*)x, ldx, gate, sgate, stats, dstats, (const E*)dy, lddy, (E*)dx, lddx,
                                  dgate, pix_per_image, C)));
  return launch_status();
}

NV_API int nervecl_upfinish_fwd(const float* conv_out, const float* lr, int64_t sN, int64_t sC, int64_t sH,
                                float* out, int N, int C, int H, int W, int s, nervecl_stream_t stream) {
  if (!conv_out || !lr || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0 || s < 1 || s > 8) return NERVECL_EINVAL;
  int64_t total = (int64_t)N * C * H * s * W * s;
  float rscale = (float)(1.0 / (double)s);
  if (upfinish_px_ok(conv_out, out, nullptr, sN, sC, sH, N, C, H, W, s)) {
    const unsigned blocks = (unsigned)cdiv((int64_t)N * H * W, 128);
#define UPF(S) upfinish_px_kernel<S, false><<<blocks, 128, 0, as_stream(stream)>>>(conv_out, lr, sN, sC, sH, out, nullptr, \
                                                                                   nullptr, N, C, H, W, rscale)
    if (s == 2) UPF(2); else if (s == 3) UPF(3); else UPF(4);
#undef UPF
    return launch_status();
  }
  upfinish_fwd_kernel<<<ew_blocks(total), 256, 0, as_stream(stream)>>>(conv_out, lr, sN, sC, sH, out, N, C, H, W, s,
                                                                       rscale);
  return launch_status();
}

NV_API int nervecl_upfinish_bwd(const float* conv_out, const float* lr, int64_t sN, int64_t sC, int64_t sH,
                                const float* dout, float* dconv, int N, int C, int H, int W, int s,
                                nervecl_stream_t stream) {
  if (!conv_out || !lr || !dout || !dconv || N <= 0 || C <= 0 || H <= 0 || W <= 0 || s < 1 || s > 8)
    return NERVECL_EINVAL;
  int64_t total = (int64_t)N * C * H * s * W * s;
  float rscale = (float)(1.0 / (double)s);
  if (upfinish_px_ok(conv_out, dout, dconv, sN, sC, sH, N, C, H, W, s)) {
    const unsigned blocks = (unsigned)cdiv((int64_t)N * H * W, 128);
#define UPB(S) upfinish_px_kernel<S, true><<<blocks, 128, 0, as_stream(stream)>>>(conv_out, lr, sN, sC, sH, nullptr, dout, \
                                                                                  dconv, N, C, H, W, rscale)
    if (s == 2) UPB(2); else if (s == 3) UPB(3); else UPB(4);
#undef UPB
    return launch_status();
  }
  upfinish_bwd_kernel<<<ew_blocks(total), 256, 0, as_stream(stream)>>>(conv_out, lr, sN, sC, sH, dout, dconv, N, C,
                                                                       H, W, s, rscale);
  return launch_status();
}

NV_API int nervecl_bicubic_blend(float* out, const float* lr, int64_t sN, int64_t sC, int64_t sH, int N, int C, int H, int W,
                                 int s, float strength, nervecl_stream_t stream) {
  if (!out || !lr || N <= 0 || C <= 0 || H <= 0 || W <= 0 || s < 1 || s > 8) return NERVECL_EINVAL;
  int64_t total = (int64_t)N * C * H * s * W * s;
  float rscale = (float)(1.0 / (double)s);
  bicubic_blend_kernel<<<ew_blocks(total), 256, 0, as_stream(stream)>>>(out, lr, sN, sC, sH, N, C, H, W, s, rscale, strength);
  return launch_status();
}
