// Shared-memory tiled 81-displacement correlation for bf16 / C = 64 (the configuration the network runs):
//   out[p, i*9+j] = (1/C) sum_c x1[p,c] * x2[p + (i-4, j-4), c]
//
// A block walks a 64-pixel-wide strip top to bottom keeping the 9 x2 rows that the current output row needs
// in a shared-memory ring (each x2 row is fetched from HBM/L2 once per strip, +12 % horizontal halo), so the
// 81x re-read of the naive kernel goes to shared memory instead of L1/L2.  A thread owns 4 consecutive pixels
// and one vertical displacement i: 36 fp32 accumulators; per 8-channel chunk it reads 4 x1 + 12 x2 vectors
// (LDS.128, XOR-swizzled so the 8 lanes of a quarter-warp hit distinct banks) for 288 FMAs.  The next rows are
// prefetched into registers while the current row is computed.
//
// The gradient kernel is the same gather run twice:
//   dx1[p,c] = (1/C) sum_d g[p,d]  * x2[p+d,c]                       (G row = g[y],             X = x2)
//   dx2[q,c] = (1/C) sum_d g~[q,d] * x1[q+d,c],  g~[q,d] = g[q+d,-d]   (G row gathered from 9 g rows, X = x1)
#include "common.cuh"

using namespace nv;

namespace {

constexpr int C = 64, RAD = 4, ND = 9, NDISP = 81;
constexpr int TW = 64;               // output pixels per strip
constexpr int PW = TW + 2 * RAD;     // staged pixels per x2 row
constexpr int ROWCH = PW * 8;        // 16-byte chunks per ring row
constexpr int RING = 9;

__device__ __forceinline__ int swz(int px, int chunk) { return px * 8 + (chunk ^ ((px >> 2) & 7)); }

__device__ __forceinline__ void cvt8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

// 16-byte chunk `ch` (0..7) of pixel (n, y, x) of a pitched NHWC bf16 tensor, zero outside the image
__device__ __forceinline__ uint4 ld_px_chunk(const bf16* __restrict__ base, int64_t ld, int n, int y, int x, int ch, int H,
                                             int W) {
  if (y < 0 || y >= H || x < 0 || x >= W) return make_uint4(0, 0, 0, 0);
  return __ldg(reinterpret_cast<const uint4*>(base + (((int64_t)n * H + y) * W + x) * ld) + ch);
}

// ---------------------------------------------------------------------------------------
// forward.  144 threads: xg = t % 16 (pixels 4xg..4xg+3 of the strip), i = t / 16.
// ---------------------------------------------------------------------------------------
constexpr int FT = 144;
constexpr int F_X2 = (ROWCH + FT - 1) / FT;      // chunks of one x2 row per thread (4)
constexpr int F_X1 = (TW * 8 + FT - 1) / FT;     // chunks of one x1 row per thread (4)

__global__ void __launch_bounds__(FT, 2)
corr_fwd_tiled_kernel(const bf16* __restrict__ x1, int64_t ld1, const bf16* __restrict__ x2, int64_t ld2,
                      bf16* __restrict__ out, int64_t ldo, int N, int H, int W, int cout_pad, int TH, int segs,
                      int strips) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* ring = reinterpret_cast<uint4*>(smem_raw);                  // [RING][PW*8]
  uint4* x1row = ring + RING * ROWCH;                                // [TW*8]
  bf16* stage = reinterpret_cast<bf16*>(x1row + TW * 8);             // [TW][cout_pad]
  const int t = threadIdx.x, xg = t & 15, i = t >> 4;
  int item = blockIdx.x;
  const int seg = item % segs; item /= segs;
  const int strip = item % strips;
  const int n = item / strips;
  const int x0 = strip * TW, y0 = seg * TH, y1 = min(H, y0 + TH);

  // prologue: x2 rows y0-4 .. y0+4 and x1 row y0
  for (int r = 0; r < RING; ++r) {
    const int yy = y0 - RAD + r;
    const int slot = ((yy % RING) + RING) % RING;
    for (int e = t; e < ROWCH; e += FT) {
      const int px = e >> 3, ch = e & 7;
      ring[slot * ROWCH + swz(px, ch)] = ld_px_chunk(x2, ld2, n, yy, x0 - RAD + px, ch, H, W);
    }
  }
  for (int e = t; e < TW * 8; e += FT) x1row[swz(e >> 3, e & 7)] = ld_px_chunk(x1, ld1, n, y0, x0 + (e >> 3), e & 7, H, W);
  for (int e = t; e < TW * (cout_pad - NDISP); e += FT)              // pad channels stay zero for every row
    stage[(e / (cout_pad - NDISP)) * cout_pad + NDISP + e % (cout_pad - NDISP)] = __float2bfloat16_rn(0.f);
  __syncthreads();

  const float inv_c = 1.f / (float)C;
  for (int y = y0; y < y1; ++y) {
    // prefetch the rows of the next iteration into registers
    uint4 p2[F_X2], p1[F_X1];
    const bool more = y + 1 < y1;
    if (more) {
#pragma unroll
      for (int k = 0; k < F_X2; ++k) {
        const int e = t + k * FT;
        p2[k] = e < ROWCH ? ld_px_chunk(x2, ld2, n, y + 1 + RAD, x0 - RAD + (e >> 3), e & 7, H, W) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < F_X1; ++k) {
        const int e = t + k * FT;
        p1[k] = e < TW * 8 ? ld_px_chunk(x1, ld1, n, y + 1, x0 + (e >> 3), e & 7, H, W) : make_uint4(0, 0, 0, 0);
      }
    }
    // ---- compute: 4 pixels x 9 horizontal displacements of vertical displacement i ----
    float acc[4][ND];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < ND; ++j) acc[a][j] = 0.f;
    const int yy = y + i - RAD;
    const uint4* row2 = ring + (((yy % RING) + RING) % RING) * ROWCH;
#pragma unroll 2
    for (int ch = 0; ch < 8; ++ch) {
      float av[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a) cvt8(x1row[swz(4 * xg + a, ch)], av[a]);
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        float bv[8];
        cvt8(row2[swz(4 * xg + m, ch)], bv);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int j = m - a;
          if (j >= 0 && j < ND) {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[a][j] = fmaf(av[a][k], bv[k], acc[a][j]);
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < ND; ++j) stage[(4 * xg + a) * cout_pad + i * ND + j] = __float2bfloat16_rn(acc[a][j] * inv_c);
    __syncthreads();
    // ---- store the staged output row (coalesced), then install the prefetched rows ----
    {
      const int cpp = cout_pad >> 3;                         // 16-byte chunks per pixel
      const uint4* st4 = reinterpret_cast<const uint4*>(stage);
      for (int e = t; e < TW * cpp; e += FT) {
        const int px = e / cpp, ch = e - px * cpp;
        if (x0 + px < W)
          *(reinterpret_cast<uint4*>(out + (((int64_t)n * H + y) * W + x0 + px) * ldo) + ch) = st4[e];
      }
    }
    if (more) {
      const int slot = (((y + 1 + RAD) % RING) + RING) % RING;   // == slot of row y-4, no longer needed
#pragma unroll
      for (int k = 0; k < F_X2; ++k) {
        const int e = t + k * FT;
        if (e < ROWCH) ring[slot * ROWCH + swz(e >> 3, e & 7)] = p2[k];
      }
#pragma unroll
      for (int k = 0; k < F_X1; ++k) {
        const int e = t + k * FT;
        if (e < TW * 8) x1row[swz(e >> 3, e & 7)] = p1[k];
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// gradient gather.  128 threads: cc = t % 8 (channels 8cc..8cc+7), xg = t / 8 (pixels 4xg..4xg+3).
//   dx[p, c] (+)= (1/C) sum_{i,j} G[p, i*9+j] * X[p + (i-4, j-4), c]
// MODE 0: G[p,d] = g[p,d].   MODE 1: G[p,(i,j)] = g[p + (i-4, j-4), (8-i)*9 + (8-j)]  (zero outside the image).
// ---------------------------------------------------------------------------------------
constexpr int GT = 128;
constexpr int GPITCH = 98;                         // bf16 per staged G pixel (81 used): 196 B keeps 4-px strides off one bank
constexpr int G_X = (ROWCH + GT - 1) / GT;         // 5

template <int MODE>
__global__ void __launch_bounds__(GT, 2)
corr_grad_tiled_kernel(const bf16* __restrict__ X, int64_t ldX, const bf16* __restrict__ g, int64_t ldg,
                       bf16* __restrict__ dx, int64_t lddx, int accumulate, int N, int H, int W, int TH, int segs,
                       int strips) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* ring = reinterpret_cast<uint4*>(smem_raw);                  // [RING][PW*8]
  bf16* G = reinterpret_cast<bf16*>(ring + RING * ROWCH);            // [TW][GPITCH]
  const int t = threadIdx.x, cc = t & 7, xg = t >> 3;
  int item = blockIdx.x;
  const int seg = item % segs; item /= segs;
  const int strip = item % strips;
  const int n = item / strips;
  const int x0 = strip * TW, y0 = seg * TH, y1 = min(H, y0 + TH);
  const bool vec_g = ldg >= 88 && !(ldg & 7) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);

  for (int r = 0; r < RING; ++r) {
    const int yy = y0 - RAD + r;
    const int slot = ((yy % RING) + RING) % RING;
    for (int e = t; e < ROWCH; e += GT) {
      const int px = e >> 3, ch = e & 7;
      ring[slot * ROWCH + swz(px, ch)] = ld_px_chunk(X, ldX, n, yy, x0 - RAD + px, ch, H, W);
    }
  }
  const float inv_c = 1.f / (float)C;
  for (int y = y0; y < y1; ++y) {
    // ---- stage the G row of this output row ----
    if (MODE == 0) {
      if (vec_g) {                              // 11 x 16-byte chunks cover channels 0..87 of a pixel
        for (int e = t; e < TW * 11; e += GT) {
          const int px = e / 11, ch = e - px * 11;
          const int x = x0 + px;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (x < W) v = __ldg(reinterpret_cast<const uint4*>(g + (((int64_t)n * H + y) * W + x) * ldg) + ch);
          uint32_t* dst = reinterpret_cast<uint32_t*>(G + px * GPITCH + ch * 8);
          dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
      } else {
        for (int e = t; e < TW * NDISP; e += GT) {
          const int px = e / NDISP, d = e - px * NDISP;
          const int x = x0 + px;
          G[px * GPITCH + d] = x < W ? g[(((int64_t)n * H + y) * W + x) * ldg + d] : __float2bfloat16_rn(0.f);
        }
      }
    } else {
      // source-pixel major: pixel (y+i-4, sx) holds, in its 9 consecutive channels (8-i)*9 + u, the entries
      // G[px = sx - x0 - 4 + u][i*9 + 8 - u] of 9 neighbouring output pixels
      for (int e = t; e < ND * PW; e += GT) {
        const int i = e / PW, sp = e - i * PW;
        const int sy = y + i - RAD, sx = x0 - RAD + sp;
        const bool ok = sy >= 0 && sy < H && sx >= 0 && sx < W;
        const bf16* src = g + (((int64_t)n * H + sy) * W + sx) * ldg + (8 - i) * ND;
#pragma unroll
        for (int u = 0; u < ND; ++u) {
          const int px = sp - 8 + u;           // sx - x0 - 4 + u - ... (sp = sx - x0 + 4)
          if (px >= 0 && px < TW) G[px * GPITCH + i * ND + 8 - u] = ok ? src[u] : __float2bfloat16_rn(0.f);
        }
      }
    }
    // prefetch next X row into registers
    uint4 pf[G_X];
    const bool more = y + 1 < y1;
    if (more) {
#pragma unroll
      for (int k = 0; k < G_X; ++k) {
        const int e = t + k * GT;
        pf[k] = e < ROWCH ? ld_px_chunk(X, ldX, n, y + 1 + RAD, x0 - RAD + (e >> 3), e & 7, H, W) : make_uint4(0, 0, 0, 0);
      }
    }
    __syncthreads();
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[a][k] = 0.f;
#pragma unroll 1
    for (int i = 0; i < ND; ++i) {
      const int yy = y + i - RAD;
      const uint4* row = ring + (((yy % RING) + RING) % RING) * ROWCH;
      float gv[4][ND];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < ND; ++j) gv[a][j] = __bfloat162float(G[(4 * xg + a) * GPITCH + i * ND + j]);
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        float xv[8];
        cvt8(row[swz(4 * xg + m, cc)], xv);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int j = m - a;
          if (j >= 0 && j < ND) {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[a][k] = fmaf(gv[a][j], xv[k], acc[a][k]);
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int x = x0 + 4 * xg + a;
      if (x < W) {
        bf16* dp = dx + (((int64_t)n * H + y) * W + x) * lddx + 8 * cc;
        f8 o;
        if (accumulate) o = ld8(dp); else { for (int k = 0; k < 8; ++k) o.v[k] = 0.f; }
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] += acc[a][k] * inv_c;
        st8(dp, o);
      }
    }
    __syncthreads();
    if (more) {
      const int slot = (((y + 1 + RAD) % RING) + RING) % RING;
#pragma unroll
      for (int k = 0; k < G_X; ++k) {
        const int e = t + k * GT;
        if (e < ROWCH) ring[slot * ROWCH + swz(e >> 3, e & 7)] = pf[k];
      }
    }
    // the next iteration's __syncthreads (after staging G) orders these ring writes before the reads
  }
}

// rows per block so that the grid is a few waves of 2 blocks/SM
inline int pick_th(int N, int H, int strips) {
  const int64_t target = (int64_t)kSMs * 2 * 4;
  int segs = (int)imax(1, imin(H / 8 > 0 ? H / 8 : 1, cdiv(target, (int64_t)N * strips)));
  return (int)cdiv(H, segs);
}

}  // namespace

namespace nv {

bool corr_tiled_supported(int dtype, int Cc, int64_t ld1, int64_t ld2, const void* x1, const void* x2) {
  return dtype == NERVECL_BF16 && Cc == C && !(ld1 & 7) && !(ld2 & 7) && aligned(x1, 16) && aligned(x2, 16);
}

int corr_fwd_tiled(const void* x1, int64_t ld1, const void* x2, int64_t ld2, void* out, int64_t ldo, int N, int H, int W,
                   int cout_pad, cudaStream_t s) {
  const int strips = (int)cdiv(W, TW);
  const int TH = pick_th(N, H, strips);
  const int segs = (int)cdiv(H, TH);
  const size_t smem = (size_t)(RING * ROWCH + TW * 8) * 16 + (size_t)TW * cout_pad * 2;
  cudaError_t e = cudaFuncSetAttribute(corr_fwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  corr_fwd_tiled_kernel<<<(unsigned)((int64_t)N * strips * segs), FT, smem, s>>>(
      (const bf16*)x1, ld1, (const bf16*)x2, ld2, (bf16*)out, ldo, N, H, W, cout_pad, TH, segs, strips);
  return launch_status();
}

int corr_bwd_tiled(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* g, int64_t ldg, void* dx1,
                   int64_t lddx1, int acc1, void* dx2, int64_t lddx2, int acc2, int N, int H, int W, cudaStream_t s) {
  const int strips = (int)cdiv(W, TW);
  const int TH = pick_th(N, H, strips);
  const int segs = (int)cdiv(H, TH);
  const size_t smem = (size_t)(RING * ROWCH) * 16 + (size_t)TW * GPITCH * 2;
  const unsigned grid = (unsigned)((int64_t)N * strips * segs);
  cudaError_t e = cudaFuncSetAttribute(corr_grad_tiled_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(corr_grad_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  corr_grad_tiled_kernel<0><<<grid, GT, smem, s>>>((const bf16*)x2, ld2, (const bf16*)g, ldg, (bf16*)dx1, lddx1, acc1, N, H,
                                                  W, TH, segs, strips);
  corr_grad_tiled_kernel<1><<<grid, GT, smem, s>>>((const bf16*)x1, ld1, (const bf16*)g, ldg, (bf16*)dx2, lddx2, acc2, N, H,
                                                  W, TH, segs, strips);
  return launch_status();
}

}  // namespace nv
