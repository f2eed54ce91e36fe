// Tensor-core 81-displacement correlation (bf16, C = 64): forward and both gradients as banded Gram products.
//
//   out[p, i*9+j] = (1/C) sum_c x1[p,c] * x2[p + (i-4, j-4), c]
//
// For a block of 16 pixels of an image row and one vertical displacement i, the 9 horizontal displacements
// are the band k - r in [0, 8] of the 16 x 24 Gram matrix  S[r, k] = <x1[px0 + r], x2_i[px0 - 4 + k]>, i.e.
// three m16n8k16 bf16 MMAs per 16-channel k-step (mma.sync: the band is picked out of the accumulator
// FRAGMENTS per thread, which tcgen05's lane-uniform TMEM loads cannot do).  37 % of the MMA work is used, at
// ~20x fewer issued instructions than the CUDA-core version (288 FMAs + 144 conversions per 8 channels).
//
// The gradients are the transposed product with the band as the A operand:
//   dx[r, c] = (1/C) sum_i sum_k Gb_i[r, k] * X_i[px0 - 4 + k, c],   Gb_i[r, k] = G[px0 + r, i*9 + (k - r)] (0 <= k-r <= 8)
// with (G, X) = (g, x2) for dx1 and (g~, x1) for dx2, g~[q, d] = g[q + d, -d]  (same gather as motion_tiled.cu).
// Gb fragments are assembled in registers from the staged G row; X fragments come from ldmatrix.trans.
//
// Block = 64-pixel strip x TH rows, 4 warps (16 pixels each); the 9 x2 (or X) rows live in a shared-memory ring
// of 80-pixel rows (72 used + zero pad so that the unused k in [24, 32) of the second k-step read zeros), 128-byte
// pixels with the 16-byte chunk XOR-swizzled by (pixel & 7): conflict-free for ldmatrix.
#include "common.cuh"

using namespace nv;

namespace {

constexpr int C = 64, RAD = 4, ND = 9, NDISP = 81;
constexpr int TW = 64;
constexpr int PWU = TW + 2 * RAD;    // pixels of a ring row that are loaded (72)
constexpr int PW = 80;               // pixels of a ring row including the zero pad
constexpr int ROWCH = PW * 8;        // 16-byte chunks per ring row (gradient kernel)
constexpr int ROWCH_F = PWU * 8;     // forward kernel: no pad needed (k < 24)
constexpr int RING = 9;
constexpr int NT = 256;                // forward: 8 warps = 4 pixel groups x 2 halves of the vertical displacements
// Staged gradient row G[px][i][j] (gradient kernels): 10 half-words per vertical displacement i (9 + 1 unused)
// and an ODD pixel pitch, so that the two band entries (j0, j0+1) an A-fragment register needs are ALWAYS one
// aligned 32-bit word: (px*GP + i*GS + k - px) has the parity of k = 2*tig + {0,8,16,24}.  Entries outside the
// band (j0 = -1 or j0+1 = 9 at the band edge) are masked with a per-slot constant instead of zero padding.
constexpr int GS = 10;               // half-words per displacement row
constexpr int GP = 91;               // half-words per pixel (9*GS + 1)
constexpr int GFRONT = 2;            // half-words before pixel 0: the j0 = -1 word of (px 0, i 0) stays inside the buffer
constexpr int GBUF = ((GFRONT + 64 * GP) * 2 + 15) / 16 * 8;     // half-words per G buffer (16-byte multiple, TW = 64)
__device__ __forceinline__ int gidx(int px, int i, int j) { return GFRONT + px * GP + i * GS + j; }

__device__ __forceinline__ int swz(int px, int chunk) { return px * 8 + (chunk ^ (px & 7)); }

__device__ __forceinline__ uint4 ld_px_chunk(const bf16* __restrict__ base, int64_t ld, int n, int y, int x, int ch, int H,
                                             int W) {
  if (y < 0 || y >= H || x < 0 || x >= W) return make_uint4(0, 0, 0, 0);
  return __ldg(reinterpret_cast<const uint4*>(base + (((int64_t)n * H + y) * W + x) * ld) + ch);
}

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// common prologue / row installation ---------------------------------------------------------
template <int RCH>
__device__ __forceinline__ void zero_ring(uint4* ring) {
  for (int e = threadIdx.x; e < RING * RCH; e += NT) ring[e] = make_uint4(0, 0, 0, 0);
}
template <int RCH>
__device__ __forceinline__ void load_ring_row(uint4* ring, const bf16* __restrict__ X, int64_t ldX, int n, int yy, int x0,
                                              int H, int W) {
  const int slot = ((yy % RING) + RING) % RING;
  for (int e = threadIdx.x; e < PWU * 8; e += NT) {
    const int px = e >> 3, ch = e & 7;
    ring[slot * RCH + swz(px, ch)] = ld_px_chunk(X, ldX, n, yy, x0 - RAD + px, ch, H, W);
  }
}
constexpr int PF_X = (PWU * 8 + NT - 1) / NT;    // 5 chunks of a ring row per thread
constexpr int PF_1 = TW * 8 / NT;                // 4 chunks of an x1 row per thread

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 2)
corr_fwd_mma_kernel(const bf16* __restrict__ x1, int64_t ld1, const bf16* __restrict__ x2, int64_t ld2,
                    bf16* __restrict__ out, int64_t ldo, int N, int H, int W, int cout_pad, int TH, int segs, int strips) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* ring = reinterpret_cast<uint4*>(smem_raw);                  // [RING][PW*8]
  uint4* x1row = ring + RING * ROWCH_F;                                // [TW*8]
  bf16* stage = reinterpret_cast<bf16*>(x1row + TW * 8);             // [TW][cout_pad]
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int gid = lane >> 2, tig = lane & 3;
  int item = blockIdx.x;
  const int seg = item % segs; item /= segs;
  const int strip = item % strips;
  const int n = item / strips;
  const int x0 = strip * TW, y0 = seg * TH, y1 = min(H, y0 + TH);
  const int px0 = (warp & 3) * 16;                                    // this warp's pixels within the strip
  const int i_lo = (warp >> 2) ? 5 : 0, i_hi = (warp >> 2) ? ND : 5;  // ... and its vertical displacements

  zero_ring<ROWCH_F>(ring);
  for (int e = t; e < TW * cout_pad; e += NT) stage[e] = __float2bfloat16_rn(0.f);   // pad channels stay zero
  __syncthreads();
  for (int r = 0; r < RING; ++r) load_ring_row<ROWCH_F>(ring, x2, ld2, n, y0 - RAD + r, x0, H, W);
  for (int e = t; e < TW * 8; e += NT) x1row[swz(e >> 3, e & 7)] = ld_px_chunk(x1, ld1, n, y0, x0 + (e >> 3), e & 7, H, W);
  __syncthreads();

  const float inv_c = 1.f / (float)C;
  // ldmatrix row addresses of this lane: A (x1): matrices (rows 0-7 | 8-15) x (chunk 2ks | 2ks+1)
  const int a_px = px0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int a_ch = lane >> 4;
  // B (x2_i): x4 = n-tile rows (8 px') x chunks (2ks', 2ks'+1, 2ks'+2, 2ks'+3) of one n-tile
  const int b_row = lane & 7, b_ch = lane >> 3;
  for (int y = y0; y < y1; ++y) {
    uint4 p2[PF_X], p1[PF_1];
    const bool more = y + 1 < y1;
    if (more) {
#pragma unroll
      for (int k = 0; k < PF_X; ++k) {
        const int e = t + k * NT;
        p2[k] = e < PWU * 8 ? ld_px_chunk(x2, ld2, n, y + 1 + RAD, x0 - RAD + (e >> 3), e & 7, H, W) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < PF_1; ++k) {
        const int e = t + k * NT;
        p1[k] = ld_px_chunk(x1, ld1, n, y + 1, x0 + (e >> 3), e & 7, H, W);
      }
    }
    uint32_t af[4][4];                                               // A fragments of the 4 k-steps
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm4(af[ks], x1row + swz(a_px, 2 * ks + a_ch));
#pragma unroll 1
    for (int i = i_lo; i < i_hi; ++i) {
      const int yy = y + i - RAD;
      const uint4* row = ring + (((yy % RING) + RING) % RING) * ROWCH_F;
      float acc[3][4];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[nt][q] = 0.f;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const int bpx = px0 + nt * 8 + b_row;                        // ring pixel = strip pixel + RAD - RAD
#pragma unroll
        for (int h = 0; h < 2; ++h) {                                // k-steps 2h, 2h+1
          uint32_t bf[4];
          ldsm4(bf, row + swz(bpx, 4 * h + b_ch));
          mma16816(acc[nt], af[2 * h], bf[0], bf[1]);
          mma16816(acc[nt], af[2 * h + 1], bf[2], bf[3]);
        }
      }
      // band: element (r, k) of S with k = 8*nt + col is displacement j = k - r
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = gid + (q >> 1) * 8, k = nt * 8 + 2 * tig + (q & 1);
          const int j = k - r;
          if (j >= 0 && j < ND) stage[(px0 + r) * cout_pad + i * ND + j] = __float2bfloat16_rn(acc[nt][q] * inv_c);
        }
    }
    __syncthreads();
    {
      const int cpp = cout_pad >> 3;
      const uint4* st4 = reinterpret_cast<const uint4*>(stage);
      for (int e = t; e < TW * cpp; e += NT) {
        const int px = e / cpp, ch = e - px * cpp;
        if (x0 + px < W) *(reinterpret_cast<uint4*>(out + (((int64_t)n * H + y) * W + x0 + px) * ldo) + ch) = st4[e];
      }
    }
    if (more) {
      const int slot = (((y + 1 + RAD) % RING) + RING) % RING;
#pragma unroll
      for (int k = 0; k < PF_X; ++k) {
        const int e = t + k * NT;
        if (e < PWU * 8) ring[slot * ROWCH_F + swz(e >> 3, e & 7)] = p2[k];
      }
#pragma unroll
      for (int k = 0; k < PF_1; ++k) {
        const int e = t + k * NT;
        x1row[swz(e >> 3, e & 7)] = p1[k];
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// gradient gather:  dx[p, c] (+)= (1/C) sum_{i,j} G[p, i*9+j] * X[p + (i-4, j-4), c]
// MODE 0: G[p,d] = g[p,d].   MODE 1: G[p,(i,j)] = g[p + (i-4, j-4), (8-i)*9 + (8-j)]  (zero outside the image).
// ---------------------------------------------------------------------------------------
// Threads 0..255 (8 warps = 4 pixel groups x 2 halves of the 64 output channels) run the MMAs of output row y;
// threads 256..383 stage everything row y+1 needs while
// they do: the next X row into registers (installed in the ring slot row y-4 frees) and the next G row straight
// into the other half of a double-buffered G (all of a loader thread's global loads are issued before its first
// store, so one row costs ~one memory latency, hidden behind the MMAs).
constexpr int GNT = 384, GMMA = 256;
constexpr int GQ = (ND * PWU + 127) / 128;        // (i, source pixel) items per loader thread in MODE 1 (6)
constexpr int GV = (TW * 11 + 127) / 128;         // 16-byte chunks per loader thread in MODE 0 (6)
constexpr int GXP = (PWU * 8 + 127) / 128;        // X-row chunks per loader thread (5)

// 16 consecutive half-words starting at half-word `h0` (compile-time) of the word array w -> pick half-word u
template <int NW>
__device__ __forceinline__ uint16_t pick_half(const uint32_t (&w)[NW], int h) {
  return (uint16_t)((h & 1) ? (w[h >> 1] >> 16) : (w[h >> 1] & 0xFFFFu));
}

template <int MODE>
__device__ __forceinline__ void stage_G(bf16* __restrict__ G, const bf16* __restrict__ g, int64_t ldg, int n, int y, int x0,
                                        int H, int W, int lt, bool vec_g) {
  if (MODE == 0) {
    if (vec_g) {
      // thread = (pixel, half of the displacement rows): rows 0-4 live in bytes [0, 90) of the pixel = chunks 0-5,
      // rows 5-8 in bytes [90, 162) = chunks 5-10; six vector loads, compile-time half-word picks, 2-byte stores
      const int px = lt & 63, ih = lt >> 6;                 // (ih is warp-uniform)
      const bool ok = x0 + px < W;
      const uint4* src = reinterpret_cast<const uint4*>(g + (((int64_t)n * H + y) * W + x0 + px) * ldg) + ih * 5;
      uint32_t w[24];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const uint4 v = ok ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
      }
      bf16* dst = G + gidx(px, 0, 0);
      if (ih == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int j = 0; j < ND; ++j) dst[i * GS + j] = __ushort_as_bfloat16(pick_half(w, i * ND + j));
      } else {
#pragma unroll
        for (int i = 5; i < ND; ++i)
#pragma unroll
          for (int j = 0; j < ND; ++j) dst[i * GS + j] = __ushort_as_bfloat16(pick_half(w, i * ND + j - 40));
      }
    } else {
      for (int e = lt; e < TW * NDISP; e += 128) {
        const int px = e / NDISP, d = e - px * NDISP;
        G[gidx(px, d / ND, d % ND)] =
            x0 + px < W ? g[(((int64_t)n * H + y) * W + x0 + px) * ldg + d] : __float2bfloat16_rn(0.f);
      }
    }
  } else {
    // source-pixel major: pixel (y+i-4, sx) holds, in its 9 consecutive channels (8-i)*9 + u, the entries
    // G[px = sp - 8 + u][i][8 - u] of 9 neighbouring output pixels (sp = sx - x0 + 4)
    if (vec_g) {
      // one source pixel per loader thread, all nine vertical displacements: the 18 bytes of displacement row i
      // start (8-i)*18 bytes into the pixel, i.e. inside two aligned 16-byte chunks -- two vector loads and
      // compile-time half-word picks instead of nine 2-byte loads
      const int sp = lt, sx = x0 - RAD + sp;
#pragma unroll
      for (int ib = 0; ib < ND; ib += 5) {
        uint4 q0[5], q1[5];
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          const int i = ib + ii;
          q0[ii] = q1[ii] = make_uint4(0, 0, 0, 0);
          if (i < ND) {
            const int sy = y + i - RAD;
            if (sp < PWU && sy >= 0 && sy < H && sx >= 0 && sx < W) {
              const uint4* src = reinterpret_cast<const uint4*>(g + (((int64_t)n * H + sy) * W + sx) * ldg) + (((8 - i) * 18) >> 4);
              q0[ii] = __ldg(src);
              q1[ii] = __ldg(src + 1);
            }
          }
        }
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          const int i = ib + ii;
          if (i < ND && sp < PWU) {
            const uint32_t w[8] = {q0[ii].x, q0[ii].y, q0[ii].z, q0[ii].w, q1[ii].x, q1[ii].y, q1[ii].z, q1[ii].w};
            const int sh = (((8 - i) * 18) & 15) >> 1;          // first half-word inside the two chunks
#pragma unroll
            for (int u = 0; u < ND; ++u) {
              const int px = sp - 8 + u;
              if (px >= 0 && px < TW) G[gidx(px, i, 8 - u)] = __ushort_as_bfloat16(pick_half(w, sh + u));
            }
          }
        }
      }
      return;
    }
    uint16_t v[GQ][ND];
#pragma unroll
    for (int k = 0; k < GQ; ++k) {
      const int e = lt + k * 128;
      const int i = e / PWU, sp = e - i * PWU;
      const int sy = y + i - RAD, sx = x0 - RAD + sp;
      const bool ok = e < ND * PWU && sy >= 0 && sy < H && sx >= 0 && sx < W;
      const bf16* src = g + (((int64_t)n * H + sy) * W + sx) * ldg + (8 - i) * ND;
#pragma unroll
      for (int u = 0; u < ND; ++u) v[k][u] = ok ? __bfloat16_as_ushort(src[u]) : (uint16_t)0;
    }
#pragma unroll
    for (int k = 0; k < GQ; ++k) {
      const int e = lt + k * 128;
      const int i = e / PWU, sp = e - i * PWU;
      if (e < ND * PWU) {
#pragma unroll
        for (int u = 0; u < ND; ++u) {
          const int px = sp - 8 + u;
          if (px >= 0 && px < TW) G[gidx(px, i, 8 - u)] = __ushort_as_bfloat16(v[k][u]);
        }
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(GNT, 2)
corr_grad_mma_kernel(const bf16* __restrict__ X, int64_t ldX, const bf16* __restrict__ g, int64_t ldg,
                     bf16* __restrict__ dx, int64_t lddx, int accumulate, int N, int H, int W, int TH, int segs,
                     int strips) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* ring = reinterpret_cast<uint4*>(smem_raw);                  // [RING][PWU*8]
  bf16* Gbuf = reinterpret_cast<bf16*>(ring + RING * ROWCH_F);       // [2][GBUF]: staged gradient rows (see GS / GP)
  constexpr int OP = 32 + 4;                                         // fp32 output staging pitch (aliases the G in use)
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const bool loader = t >= GMMA;
  const int lt = t - GMMA;                                           // loader thread index (0..127)
  const int chalf = (warp >> 2) & 1;                                 // MMA warps: which 32 output channels
  const int gid = lane >> 2, tig = lane & 3;
  int item = blockIdx.x;
  const int seg = item % segs; item /= segs;
  const int strip = item % strips;
  const int n = item / strips;
  const int x0 = strip * TW, y0 = seg * TH, y1 = min(H, y0 + TH);
  const int px0 = (warp & 3) * 16;
  const bool vec_g = ldg >= 88 && !(ldg & 7) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);

  // prologue (all threads): ring rows y0-4 .. y0+4 and G of row y0
  for (int r = 0; r < RING; ++r) {
    const int yy = y0 - RAD + r;
    const int slot = ((yy % RING) + RING) % RING;
    for (int e = t; e < PWU * 8; e += GNT) ring[slot * ROWCH_F + swz(e >> 3, e & 7)] = ld_px_chunk(X, ldX, n, yy, x0 - RAD + (e >> 3), e & 7, H, W);
  }
  if (loader) stage_G<MODE>(Gbuf + (y0 & 1) * GBUF, g, ldg, n, y0, x0, H, W, lt, vec_g);
  __syncthreads();

  // A-fragment register (ks, q) of displacement row i is ONE aligned word of the staged row: byte offset inside a
  // G buffer (+ i*GS*2) and the mask that clears band-edge neighbours -- both fixed per thread
  uint32_t aoff[8], amask[8];
#pragma unroll
  for (int sl = 0; sl < 8; ++sl) {
    const int ks = sl >> 2, q = sl & 3;
    const int r = gid + (q & 1) * 8;
    const int k = ks * 16 + 2 * tig + (q >> 1) * 8;
    const int j0 = k - r;
    const bool lo = j0 >= 0 && j0 < ND, hi = j0 + 1 >= 0 && j0 + 1 < ND;
    amask[sl] = (lo ? 0xFFFFu : 0u) | (hi ? 0xFFFF0000u : 0u);
    aoff[sl] = (lo || hi) ? (uint32_t)(gidx(px0 + r, 0, 0) + j0) * 2u : 0u;
  }

  const float inv_c = 1.f / (float)C;
  const int b_k = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int b_ch = lane >> 4;
  for (int y = y0; y < y1; ++y) {
    const bool more = y + 1 < y1;
    bf16* G = Gbuf + (y & 1) * GBUF;
    uint4 pf[GXP];
    if (loader) {
      if (more) {
#pragma unroll
        for (int k = 0; k < GXP; ++k) {
          const int e = lt + k * 128;
          pf[k] = e < PWU * 8 ? ld_px_chunk(X, ldX, n, y + 1 + RAD, x0 - RAD + (e >> 3), e & 7, H, W) : make_uint4(0, 0, 0, 0);
        }
        stage_G<MODE>(Gbuf + ((y + 1) & 1) * GBUF, g, ldg, n, y + 1, x0, H, W, lt, vec_g);
      }
    } else {
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[nt][q] = 0.f;
      const uint8_t* Gb = reinterpret_cast<const uint8_t*>(G);
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        const int yy = y + i - RAD;
        const uint4* row = ring + (((yy % RING) + RING) % RING) * ROWCH_F;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          // A fragment = band of G: element (r, k) = G[px0 + r][i][k - r] for 0 <= k - r <= 8
          uint32_t af[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            af[q] = *reinterpret_cast<const uint32_t*>(Gb + aoff[ks * 4 + q] + i * GS * 2) & amask[ks * 4 + q];
          // ring pixel of this lane's k row; k >= 24 lies outside every band (A is zero there), so clamp into the
          // row instead of padding it: the operand only has to be finite
          const int kpx = min(px0 + ks * 16 + b_k, PWU - 1);
#pragma unroll
          for (int q = 0; q < 2; ++q) {                               // this warp's n-tiles 2q, 2q+1 (of its channel half)
            uint32_t bf[4];
            ldsm4_t(bf, row + swz(kpx, 2 * (2 * chalf + q) + b_ch));
            mma16816(acc[2 * q], af, bf[0], bf[1]);
            mma16816(acc[2 * q + 1], af, bf[2], bf[3]);
          }
        }
      }
      // stage the 16 x 32 results of the four warps of one channel half (over this row's G), then one coalesced
      // (read-modify-)write of that half by all 256 MMA threads; only the MMA warps take part (named barrier 1)
      float* ost = reinterpret_cast<float*>(G);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (chalf == half) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int r = gid + (q >> 1) * 8, c = nt * 8 + 2 * tig + (q & 1);
              ost[(px0 + r) * OP + c] = acc[nt][q] * inv_c;
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        {
          const int px = t >> 2, ch = t & 3;                            // TW * 4 = 256 items: one per MMA thread
          const int x = x0 + px;
          if (x < W) {
            bf16* dp = dx + (((int64_t)n * H + y) * W + x) * lddx + 32 * half + 8 * ch;
            f8 o;
            if (accumulate) o = ld8(dp); else { for (int k = 0; k < 8; ++k) o.v[k] = 0.f; }
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] += ost[px * OP + 8 * ch + k];
            st8(dp, o);
          }
        }
      }
    }
    __syncthreads();                       // MMAs of row y are done with ring slot (y-4); G of row y+1 is staged
    if (loader && more) {
      const int slot = (((y + 1 + RAD) % RING) + RING) % RING;
#pragma unroll
      for (int k = 0; k < GXP; ++k) {
        const int e = lt + k * 128;
        if (e < PWU * 8) ring[slot * ROWCH_F + swz(e >> 3, e & 7)] = pf[k];
      }
    }
    __syncthreads();
  }
}

inline int pick_th(int N, int H, int strips) {
  const int64_t target = (int64_t)kSMs * 2 * 4;
  int segs = (int)imax(1, imin(H / 8 > 0 ? H / 8 : 1, cdiv(target, (int64_t)N * strips)));
  return (int)cdiv(H, segs);
}

}  // namespace

namespace nv {

int corr_fwd_mma(const void* x1, int64_t ld1, const void* x2, int64_t ld2, void* out, int64_t ldo, int N, int H, int W,
                 int cout_pad, cudaStream_t s) {
  const int strips = (int)cdiv(W, TW);
  const int TH = pick_th(N, H, strips);
  const int segs = (int)cdiv(H, TH);
  const size_t smem = (size_t)(RING * ROWCH_F + TW * 8) * 16 + (size_t)TW * cout_pad * 2;
  cudaError_t e = cudaFuncSetAttribute(corr_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  corr_fwd_mma_kernel<<<(unsigned)((int64_t)N * strips * segs), NT, smem, s>>>(
      (const bf16*)x1, ld1, (const bf16*)x2, ld2, (bf16*)out, ldo, N, H, W, cout_pad, TH, segs, strips);
  return launch_status();
}

int corr_bwd_mma(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* g, int64_t ldg, void* dx1,
                 int64_t lddx1, int acc1, void* dx2, int64_t lddx2, int acc2, int N, int H, int W, cudaStream_t s) {
  const int strips = (int)cdiv(W, TW);
  const int TH = pick_th(N, H, strips);
  const int segs = (int)cdiv(H, TH);
  const size_t smem = (size_t)(RING * ROWCH_F) * 16 + 2 * (size_t)GBUF * 2;
  const unsigned grid = (unsigned)((int64_t)N * strips * segs);
  cudaError_t e = cudaFuncSetAttribute(corr_grad_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(corr_grad_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  corr_grad_mma_kernel<0><<<grid, GNT, smem, s>>>((const bf16*)x2, ld2, (const bf16*)g, ldg, (bf16*)dx1, lddx1, acc1, N, H, W,
                                                TH, segs, strips);
  corr_grad_mma_kernel<1><<<grid, GNT, smem, s>>>((const bf16*)x1, ld1, (const bf16*)g, ldg, (bf16*)dx2, lddx2, acc2, N, H, W,
                                                TH, segs, strips);
  return launch_status();
}

}  // namespace nv
