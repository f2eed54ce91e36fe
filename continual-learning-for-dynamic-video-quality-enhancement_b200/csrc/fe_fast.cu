// Bandwidth-tuned feature-extractor kernels (C % 8 == 0, C/8 a power of two <= 32):
//   * depthwise 3x3 forward / data-gradient and weight-gradient: a "walker" (the C/8 threads that cover one
//     pixel's channels with 128-bit accesses) slides a 3x3 register window along a 32-pixel row segment, so
//     each output pixel costs 3 new vector loads instead of 9 and the 72 filter taps of the thread's 8
//     channels live in registers for the whole kernel (the naive kernel re-loaded them per pixel and was
//     LSU-issue-bound at ~7x the HBM time);
//   * BatchNorm+ReLU forward and backward-apply with every per-channel coefficient hoisted into registers.
#include "common.cuh"
#include "tc_common.cuh"

using namespace nv;

namespace {

constexpr int kThreadsFe = 256;
#ifndef DW_TH_ROWS
#define DW_TH_ROWS 8
#endif
#ifndef DW_NBUF
#define DW_NBUF 2
#endif
constexpr int DW_TH = DW_TH_ROWS;         // output rows per tile
constexpr int kDwBufs = DW_NBUF;           // tile buffers per CTA (TMA loads in flight: kDwBufs - 1)

// Depthwise tiles are brought in by TMA: one 4-D box {C, TWp + 2, TH + 2, 1} per tile with out-of-image pixels
// zero-filled (= the conv padding), landing in shared memory as [(TH+2)][(TWp+2)][C].  The per-thread cp.async
// staging this replaces cost ~400 of the kernel's 2100 instructions per thread and tile, and the kernel was
// instruction-issue bound (72 % SM throughput at 49 % of the HBM roofline, profiles/r01z_summary.md).
struct TileXY { int n, y0, x0; };
__device__ __forceinline__ TileXY tile_xy(int64_t tile, int tiles_x, int tiles_y, int TWp, int TH) {
  const int txi = (int)(tile % tiles_x);
  const int64_t r = tile / tiles_x;
  return TileXY{(int)(r / tiles_y), (int)(r % tiles_y) * TH, txi * TWp};
}

// One output column of DW_TH rows: each input row's three neighbour vectors are loaded / converted once and feed
// the three output rows they belong to (rolling accumulators).  FULL: all DW_TH rows are inside the image.
// MODE 0: y = conv;  1: y += conv;  2: y = (conv + add) where mask > 0, else 0 (`pre`: the thread's add / mask vectors
// of all DW_TH rows, requested before the tile's barrier wait so that their latency hides behind it).
template <typename T> struct DwPre { uint4 add[DW_TH], mask[DW_TH]; };
__device__ __forceinline__ f8 unpack8(const uint4& t) {
  f8 r;
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
  return r;
}
template <typename T, int MODE, bool FULL>
__device__ __forceinline__ void dw_column(const T* __restrict__ col, int rstride, int C, const float (&wt)[3][3][8],
                                          T* __restrict__ yp, int64_t ystride, int nrows, const DwPre<T>* pre = nullptr) {
  float acc[3][8];
#pragma unroll
  for (int ri = 0; ri < DW_TH + 2; ++ri) {
    float v[3][8];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const f8 t = ld8(col + ri * rstride + kx * C);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[kx][k] = t.v[k];
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int oi = ri - ky;                     // output row (tile-local) that uses this input row as tap row ky
      if (oi < 0 || oi >= DW_TH) continue;
      float* a = acc[oi % 3];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = (ky == 0 && kx == 0) ? v[kx][k] * wt[ky][kx][k] : fmaf(v[kx][k], wt[ky][kx][k], a[k]);
    }
    const int od = ri - 2;                        // this output row is complete
    if (od >= 0 && (FULL || od < nrows)) {
      const float* a = acc[od % 3];
      f8 o;
      if (MODE == 1) {
        o = ld8(yp);
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] += a[k];
      } else if (MODE == 2) {
        const f8 ad = unpack8(pre->add[od]), mk = unpack8(pre->mask[od]);
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = mk.v[k] > 0.f ? a[k] + ad.v[k] : 0.f;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = a[k];
      }
      st8(yp, o);
      yp += ystride;
    }
  }
}

// y[p,c] (+)= sum_taps x[p+tap,c] * w[c][tap]   (flip: 180-degree rotated filter = data gradient)
// block = 256 threads = (C/8 channel groups) x TWp pixel columns; tile = DW_TH rows x TWp columns; two tile
// buffers so the next tile's TMA load is in flight while this one is computed.
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreadsFe, MODE == 2 ? 1 : 2)
dwconv_tile_kernel(const __grid_constant__ CUtensorMap tmx, const float* __restrict__ w, T* __restrict__ y, int64_t ldy,
                   int N, int H, int W, int C, int flip, int tiles_x, int tiles_y, const T* __restrict__ add = nullptr,
                   int64_t ldadd = 0, const T* __restrict__ mask = nullptr, int64_t ldmask = 0) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[kDwBufs];
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
  const size_t tile_elems = (size_t)(DW_TH + 2) * PWs * C;
  const uint32_t tile_bytes = (uint32_t)(tile_elems * sizeof(T));
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmx);
    for (int b = 0; b < kDwBufs; ++b) tc::mbar_init(&bars[b], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  auto load = [&](int b, int64_t tile) {
    const TileXY t0 = tile_xy(tile, tiles_x, tiles_y, TWp, DW_TH);
    tc::mbar_expect_tx(&bars[b], tile_bytes);
    tc::tma_load_4d(sm + b * tile_elems, &tmx, &bars[b], 0, t0.x0 - 1, t0.y0 - 1, t0.n);
  };
  if (threadIdx.x == 0)
    for (int b = 0; b < kDwBufs - 1; ++b)
      if ((int64_t)blockIdx.x + (int64_t)b * gridDim.x < ntiles) load(b, blockIdx.x + (int64_t)b * gridDim.x);
  int buf = 0;
  uint32_t par = 0;                        // phase parity bit of each buffer's barrier
  const int rstride = PWs * C;
  const int64_t ystride = (int64_t)W * ldy;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const TileXY t = tile_xy(tile, tiles_x, tiles_y, TWp, DW_TH);
    // refill the buffer that was read in the previous round (every read of it ended before the last barrier)
    const int nb = buf == 0 ? kDwBufs - 1 : buf - 1;
    const int64_t nt = tile + (int64_t)(kDwBufs - 1) * gridDim.x;
    if (threadIdx.x == 0 && nt < ntiles) load(nb, nt);
    const int xx = t.x0 + tx;
    DwPre<T> pre;
    if (MODE == 2 && sizeof(T) == 2 && xx < W) {
      const int64_t p0 = ((int64_t)t.n * H + t.y0) * W + xx;
#pragma unroll
      for (int r = 0; r < DW_TH; ++r) {
        const bool ok = t.y0 + r < H;
        pre.add[r] = ok ? *reinterpret_cast<const uint4*>(add + (p0 + (int64_t)r * W) * ldadd + c0) : make_uint4(0, 0, 0, 0);
        pre.mask[r] = ok ? *reinterpret_cast<const uint4*>(mask + (p0 + (int64_t)r * W) * ldmask + c0) : make_uint4(0, 0, 0, 0);
      }
    }
    tc::mbar_wait(&bars[buf], (par >> buf) & 1u);
    par ^= 1u << buf;
    if (xx < W) {
      const T* col = sm + buf * tile_elems + tx * C + c0;                // this thread's column of the tile
      T* yp = y + (((int64_t)t.n * H + t.y0) * W + xx) * ldy + c0;
      const int nrows = H - t.y0;
      if (nrows >= DW_TH) dw_column<T, MODE, true>(col, rstride, C, wt, yp, ystride, DW_TH, &pre);
      else dw_column<T, MODE, false>(col, rstride, C, wt, yp, ystride, nrows, &pre);
    }
    __syncthreads();                       // every read of this buffer is done before the next round refills it
    if (++buf == kDwBufs) buf = 0;
  }
}

// dw[c][tap] += sum_p dy[p,c] * x[p+tap,c]
// Same tiling with WG_TH-row tiles and two (x, dy) tile buffers; the 72 partial sums of the thread's 8 channels
// stay in registers across all of the block's tiles.  Out-of-image dy pixels are zero-filled by TMA.
constexpr int WG_TH = 4;

template <typename T>
__global__ void __launch_bounds__(kThreadsFe, 2)
dwconv_wgrad_tile_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmd,
                         float* __restrict__ dw, int N, int H, int W, int C, int tiles_x, int tiles_y) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
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
  const size_t x_elems = (size_t)(WG_TH + 2) * PWs * C;
  const size_t d_elems = (size_t)WG_TH * TWp * C;
  const size_t buf_elems = x_elems + d_elems;                      // [x tile | dy tile]
  const uint32_t buf_bytes = (uint32_t)(buf_elems * sizeof(T));
  const int64_t ntiles = (int64_t)N * tiles_y * tiles_x;
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmx);
    tc::prefetch_tmap(&tmd);
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  auto load = [&](int b, int64_t tile) {
    const TileXY t = tile_xy(tile, tiles_x, tiles_y, TWp, WG_TH);
    T* dst = sm + b * buf_elems;
    tc::mbar_expect_tx(&bars[b], buf_bytes);
    tc::tma_load_4d(dst, &tmx, &bars[b], 0, t.x0 - 1, t.y0 - 1, t.n);
    tc::tma_load_4d(dst + x_elems, &tmd, &bars[b], 0, t.x0, t.y0, t.n);
  };
  if (threadIdx.x == 0 && (int64_t)blockIdx.x < ntiles) load(0, blockIdx.x);
  int buf = 0;
  uint32_t par = 0;                        // phase parity bit of each buffer's barrier
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    if (threadIdx.x == 0 && tile + gridDim.x < ntiles) load(buf ^ 1, tile + gridDim.x);
    tc::mbar_wait(&bars[buf], (par >> buf) & 1u);
    par ^= 1u << buf;
    {
      const T* smx = sm + buf * buf_elems + tx * C + c0;
      const T* smd = sm + buf * buf_elems + x_elems + tx * C + c0;
      // input row ri pairs with the dy rows ri - ky (tap row ky): keep the three live dy vectors in registers
      float d[3][8];
#pragma unroll
      for (int ri = 0; ri < WG_TH + 2; ++ri) {
        if (ri < WG_TH) {
          const f8 t = ld8(smd + ri * TWp * C);
#pragma unroll
          for (int k = 0; k < 8; ++k) d[ri % 3][k] = t.v[k];
        }
        float v[3][8];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const f8 t = ld8(smx + (ri * PWs + kx) * C);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[kx][k] = t.v[k];
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int oi = ri - ky;
          if (oi < 0 || oi >= WG_TH) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[ky][kx][k] = fmaf(d[oi % 3][k], v[kx][k], acc[ky][kx][k]);
        }
      }
    }
    __syncthreads();
  }
  // block reduction: shared fp32 atomics (one address per (channel, tap)), then one global atomic each
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
__global__ void __launch_bounds__(kThreadsFe, 3)
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
__global__ void __launch_bounds__(kThreadsFe, 3)
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
// statistics pass and (g, g*xhat) with g = dy * [bn(x) > 0] for the backward reduction.
// One wave of kSMs*4 (backward: kSMs*3) blocks, flattened over (group, slab): a block streams one contiguous slab of pixels of one
// group; a thread owns 4 channels of every (256/cg)-th pixel and keeps U independent vector loads per operand in
// flight (64 KB per SM, enough to cover the HBM latency); fp32 partials go to fp64 every 16 rounds.
template <typename T, bool BWD, int U>
__global__ void __launch_bounds__(kThreadsFe, BWD ? 3 : 4)
bn_sums_fast_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                    const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ beta, int C,
                    int64_t npix, int bpg, double* __restrict__ sums) {
  const int cg = C >> 2;                     // threads per pixel: a power of two <= 256
  const int lanes = kThreadsFe / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg, c0 = g << 2;
  const int grp = blockIdx.x / bpg, slab = blockIdx.x % bpg;
  const int64_t per = cdiv(npix, bpg);
  const int64_t p_lo = (int64_t)slab * per, p_hi = p_lo + per < npix ? p_lo + per : npix;
  const T* xb = x + (int64_t)grp * npix * ldx + c0;
  const T* db = BWD ? dy + (int64_t)grp * npix * lddy + c0 : nullptr;
  float mean[4], is[4], ga[4], be[4];
  if (BWD) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + k;
      mean[k] = stat[((int64_t)grp * C + c) * 2];
      is[k] = stat[((int64_t)grp * C + c) * 2 + 1];
      ga[k] = gamma[c];
      be[k] = beta[c];
    }
  }
  double sa[4], sb[4];
  float fa[4], fb[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { sa[k] = sb[k] = 0.0; fa[k] = fb[k] = 0.f; }
  auto add = [&](const f4& v, const f4& d) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (BWD) {
        const float xh = (v.v[k] - mean[k]) * is[k];
        const float gk = fmaf(xh, ga[k], be[k]) > 0.f ? d.v[k] : 0.f;   // == bn_value(...) > 0
        fa[k] += gk;
        fb[k] = fmaf(gk, xh, fb[k]);
      } else {
        fa[k] += v.v[k];
        fb[k] = fmaf(v.v[k], v.v[k], fb[k]);
      }
    }
  };
  int64_t p = p_lo + lane;
  int cnt = 0;
  for (; p + (int64_t)(U - 1) * lanes < p_hi; p += (int64_t)U * lanes) {
    f4 v[U], d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = ld4(xb + (p + (int64_t)u * lanes) * ldx);
      if (BWD) d[u] = ld4(db + (p + (int64_t)u * lanes) * lddy);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) add(v[u], BWD ? d[u] : v[u]);
    if (++cnt == 64 / U) {                   // 64 pixels: flush the fp32 partials to fp64 before they lose bits
#pragma unroll
      for (int k = 0; k < 4; ++k) { sa[k] += fa[k]; sb[k] += fb[k]; fa[k] = fb[k] = 0.f; }
      cnt = 0;
    }
  }
  for (; p < p_hi; p += lanes) {
    const f4 v = ld4(xb + p * ldx);
    add(v, BWD ? ld4(db + p * lddy) : v);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { sa[k] += fa[k]; sb[k] += fb[k]; }
  // lanes of one warp that own the same channels, then the warps of the block, then one global atomic per value
  for (int o = cg; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sa[k] += __shfl_xor_sync(0xffffffffu, sa[k], o);
      sb[k] += __shfl_xor_sync(0xffffffffu, sb[k], o);
    }
  }
  extern __shared__ double dred[];  // [C][2]
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) dred[i] = 0.0;
  __syncthreads();
  if ((threadIdx.x & 31) < cg) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&dred[(c0 + k) * 2 + 0], sa[k]);
      atomicAdd(&dred[(c0 + k) * 2 + 1], sb[k]);
    }
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
// TMA-tiled depthwise kernels: a box holds at most 256 pixels per row and 256 channels
bool dw_tiled_supported(int C, int64_t lda, int64_t ldb, const void* a, const void* b) {
  return fe_fast_supported(C, lda, ldb, a, b) && C >= 16 && C <= 256 && tc::encode_fn() != nullptr;
}

// 4-D tensor map over an NHWC activation with a {C, bw, bh, 1} box (no swizzle: the tile lands as [bh][bw][C])
static bool encode_tile_map(CUtensorMap* m, const void* base, int64_t ld, int dtype, int N, int H, int W, int C, int bw, int bh) {
  tc::EncodeTiledFn enc = tc::encode_fn();
  if (!enc) return false;
  const cuuint64_t esz = dtype == NERVECL_F32 ? 4 : 2;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * esz, (cuuint64_t)W * ld * esz, (cuuint64_t)H * W * ld * esz};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, dtype == NERVECL_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
             const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int dwconv_slide(const void* x, int64_t ldx, const float* w, void* y, int64_t ldy, int dtype, int N, int H, int W, int C,
                 int flip, int accumulate, cudaStream_t s, const void* add, int64_t ldadd, const void* mask, int64_t ldmask) {
  const int TWp = kThreadsFe / (C >> 3);
  const int tiles_x = (W + TWp - 1) / TWp, tiles_y = (H + DW_TH - 1) / DW_TH;
  const int64_t ntiles = (int64_t)N * tiles_x * tiles_y;
  const bool masked = add != nullptr;                                          // (bf16 only: checked by the caller)
  const int blocks = (int)imax(1, imin(ntiles, kSMs * (masked ? 1 : 2)));
  const size_t esz = dtype == NERVECL_F32 ? 4 : 2;
  const size_t smem = kDwBufs * (size_t)(DW_TH + 2) * (TWp + 2) * C * esz;          // the tile buffers
  CUtensorMap tmx;
  if (!encode_tile_map(&tmx, x, ldx, dtype, N, H, W, C, TWp + 2, DW_TH + 2)) return NERVECL_EUNSUPPORTED;
  cudaError_t e;
#define NV_DW_LAUNCH(E, A, ...)                                                                                      \
  e = cudaFuncSetAttribute(dwconv_tile_kernel<E, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
  if (e != cudaSuccess) return (int)e;                                                                              \
  dwconv_tile_kernel<E, A><<<blocks, kThreadsFe, smem, s>>>(tmx, w, (E*)y, ldy, N, H, W, C, flip, tiles_x, tiles_y, ##__VA_ARGS__)
  if (masked) {
    NV_DW_LAUNCH(bf16, 2, (const bf16*)add, ldadd, (const bf16*)mask, ldmask);
  } else if (dtype == NERVECL_F32) {
    if (accumulate) { NV_DW_LAUNCH(float, 1); } else { NV_DW_LAUNCH(float, 0); }
  } else {
    if (accumulate) { NV_DW_LAUNCH(bf16, 1); } else { NV_DW_LAUNCH(bf16, 0); }
  }
#undef NV_DW_LAUNCH
  return launch_status();
}

int dwconv_wgrad_slide(const void* x, int64_t ldx, const void* dy, int64_t lddy, int dtype, float* dw, int N, int H, int W,
                       int C, cudaStream_t s) {
  const int TWp = kThreadsFe / (C >> 3);
  const int tiles_x = (W + TWp - 1) / TWp, tiles_y = (H + WG_TH - 1) / WG_TH;
  const int64_t ntiles = (int64_t)N * tiles_x * tiles_y;
  const int blocks = (int)imax(1, imin(ntiles, kSMs * 2));
  const size_t esz = dtype == NERVECL_F32 ? 4 : 2;
  const size_t smem = imax((int64_t)(2 * ((size_t)(WG_TH + 2) * (TWp + 2) + (size_t)WG_TH * TWp) * C * esz), (int64_t)C * 9 * 4);
  CUtensorMap tmx, tmd;
  if (!encode_tile_map(&tmx, x, ldx, dtype, N, H, W, C, TWp + 2, WG_TH + 2) ||
      !encode_tile_map(&tmd, dy, lddy, dtype, N, H, W, C, TWp, WG_TH))
    return NERVECL_EUNSUPPORTED;
  cudaError_t e;
#define NV_DW_LAUNCH(E)                                                                                              \
  e = cudaFuncSetAttribute(dwconv_wgrad_tile_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
  if (e != cudaSuccess) return (int)e;                                                                              \
  dwconv_wgrad_tile_kernel<E><<<blocks, kThreadsFe, smem, s>>>(tmx, tmd, dw, N, H, W, C, tiles_x, tiles_y)
  if (dtype == NERVECL_F32) { NV_DW_LAUNCH(float); } else { NV_DW_LAUNCH(bf16); }
#undef NV_DW_LAUNCH
  return launch_status();
}

bool bn_sums_fast_supported(int C, int64_t lda, int64_t ldb) {
  const int cg = C >> 2;
  return !(C & 3) && cg <= kThreadsFe && (cg & (cg - 1)) == 0 && !(lda & 3) && !(ldb & 3);
}

int bn_sums_fast(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* stat, const float* gamma,
                 const float* beta, int dtype, int C, int64_t npix, int groups, double* sums, cudaStream_t s) {
  const int lanes = kThreadsFe / (C >> 2);
  // one resident wave shared evenly by the groups; never more slabs than 64-pixel rounds
  const int bpg = (int)imax(1, imin((kSMs * (dy ? 3 : 4)) / groups, cdiv(npix, (int64_t)lanes * 64)));
  const int blocks = bpg * groups;
  const size_t smem = (size_t)C * 2 * sizeof(double);
  if (dy) {
    NV_DISPATCH_DTYPE(dtype, E, (bn_sums_fast_kernel<E, true, 4><<<blocks, kThreadsFe, smem, s>>>(
                                    (const E*)x, ldx, (const E*)dy, lddy, stat, gamma, beta, C, npix, bpg, sums)));
  } else {
    NV_DISPATCH_DTYPE(dtype, E, (bn_sums_fast_kernel<E, false, 8><<<blocks, kThreadsFe, smem, s>>>(
                                    (const E*)x, ldx, nullptr, 0, nullptr, nullptr, nullptr, C, npix, bpg, sums)));
  }
  return launch_status();
}

int bn_relu_fwd_fast(const void* x, int64_t ldx, const float* stat, const float* gamma, const float* beta, const void* res,
                     int64_t ldres, void* y, int64_t ldy, int dtype, int C, int64_t npix, int groups, cudaStream_t s) {
  const int chunks = (int)imax(1, imin(cdiv(npix * (C >> 3), kThreadsFe * 4), (kSMs * 3) / groups));
  dim3 grid(chunks, groups);
  NV_DISPATCH_DTYPE(dtype, E, (bn_relu_fwd_fast_kernel<E><<<grid, kThreadsFe, 0, s>>>((const E*)x, ldx, stat, gamma, beta,
                                                                                      (const E*)res, ldres, (E*)y, ldy, C,
                                                                                      npix)));
  return launch_status();
}

int bn_bwd_apply_fast(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* stat, const float* gamma,
                      const float* beta, const double* bsums, void* dx, int64_t lddx, float* dgamma, float* dbeta, int dtype,
                      int C, int64_t npix, int groups, int training, cudaStream_t s) {
  const int chunks = (int)imax(1, imin(cdiv(npix * (C >> 3), kThreadsFe * 4), (kSMs * 3) / groups));
  dim3 grid(chunks, groups);
  NV_DISPATCH_DTYPE(dtype, E, (bn_bwd_apply_fast_kernel<E><<<grid, kThreadsFe, 0, s>>>(
                                  (const E*)x, ldx, (const E*)dy, lddy, stat, gamma, beta, bsums, (E*)dx, lddx, dgamma, dbeta,
                                  C, npix, groups, training)));
  return launch_status();
}

}  // namespace nv
