// 81-displacement correlation on tcgen05 (bf16, C = 64): 2-D pixel tiles instead of the row-banded mma.sync Gram
// products of motion_mma.cu.
//
//   out[p, i*9+j] = (1/C) sum_c x1[p,c] * x2[p + (i-4, j-4), c]
//
// For a tile of 8 x 16 pixels p the 81 displacements reach the 16 x 24 pixel region q around it, so ONE dense Gram
// product  S[p, q] = <x1[p], x2[q]>  (M = 128 tile pixels, N = 384 region pixels, K = 64 channels: 8 tcgen05.mma with
// N = 192) holds every value the tile needs: 21 % of the product is used, against 6 % for a row band of the same M, and
// the tensor work is 768 cycles per 128 pixels.  Both operands are plain K-major TMA boxes ({64 ch, 16 px, 8 rows} and
// {64 ch, 24 px, 16 rows}, zero fill outside the image = the correlation's zero padding).
//
// The accumulator lands in tensor memory as S[lane = tile pixel][column = region pixel]; pixel (py, px) needs columns
// (py + i) * 24 + px + j -- a lane-dependent diagonal that tcgen05.ld (same columns for all lanes) cannot address.  Each
// epilogue warp (32 lanes = tile rows 2w, 2w+1) therefore bounces one region row (24 columns) at a time through a
// [column][lane] shared-memory slab -- stores and the diagonal reads are both bank-conflict free -- and keeps its 81
// results in registers; lanes 0-15 and 16-31 sit one tile row apart, so at step i they read two consecutive slabs and the
// register index i*9+j stays compile-time.  A thread then writes its pixel's 96 channels (81 + zero pad) as six 32-byte
// stores.  Shared-memory bandwidth bounds the epilogue at ~1300 cycles per tile, the same order as the 64 KB of TMA
// loads per tile, i.e. the kernel sits near the HBM roofline of its algorithmic bytes instead of 3-4x above it.
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

using namespace nv;
using namespace nv::tc;

namespace {

constexpr int C = 64, ND = 9, NDISP = 81;
constexpr int TY = 8, TX = 16;                       // tile pixels (TY * TX = 128 = UMMA M)
constexpr int RY = TY + 8, RX = TX + 8;              // region pixels (16 x 24 = 384 = 2 x UMMA N)
constexpr int NHALF = RY * RX / 2;                   // 192
constexpr uint32_t ROWB = C * 2;                     // bytes per pixel
constexpr uint32_t A_BYTES = TY * TX * ROWB;         // 16 KB
constexpr uint32_t B_BYTES = RY * RX * ROWB;         // 48 KB
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int kStages = 2;
constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr uint32_t SLAB = RX * 32 * 4;               // one region row of one warp: [24 columns][32 lanes] fp32 = 3 KB

struct FwdArgs {
  bf16* out;
  int64_t ldo;
  int N, H, W, tiles_x, tiles_y, cout_pad;
};

__global__ void __launch_bounds__(kThreads, 1)
corr_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* slabs = smem + kStages * STAGE_BYTES;                       // [kEpiWarps][2][SLAB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(slabs + kEpiWarps * 2 * SLAB);
  uint64_t* full = bars;                 // [kStages] TMA -> MMA
  uint64_t* empty = full + kStages;      // [kStages] MMA -> TMA
  uint64_t* acc_full = empty + kStages;  // MMA -> epilogue
  uint64_t* acc_empty = acc_full + 1;    // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (int64_t)a.N * a.tiles_y * a.tiles_x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tx = (int)(t % a.tiles_x);
        const int64_t r = t / a.tiles_x;
        const int ty = (int)(r % a.tiles_y), n = (int)(r / a.tiles_y);
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], STAGE_BYTES);
        uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
        tma_load_4d(sa, &tmap_a, &full[stage], 0, tx * TX, ty * TY, n);
        tma_load_4d(sa + A_BYTES, &tmap_b, &full[stage], 0, tx * TX - 4, ty * TY - 4, n);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16((uint32_t)NHALF);
    int stage = 0;
    uint32_t phase = 0, accp = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait(acc_empty, accp ^ 1);                 // the previous tile's Gram matrix has been read out
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint64_t da = make_kmajor_desc(sa, ROWB);
        const uint64_t db0 = make_kmajor_desc(sa + A_BYTES, ROWB);
        const uint64_t db1 = make_kmajor_desc(sa + A_BYTES + NHALF * ROWB, ROWB);
#pragma unroll
        for (int k = 0; k < C / 16; ++k) {            // 32 bytes (16 channels) per k-step inside the 128-byte swizzle span
          umma_bf16(tmem_base, da + 2u * k, db0 + 2u * k, idesc, k > 0);
          umma_bf16(tmem_base + NHALF, da + 2u * k, db1 + 2u * k, idesc, k > 0);
        }
        umma_commit(&empty[stage]);
        umma_commit(acc_full);
      }
      __syncwarp();
      accp ^= 1;
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ================= epilogue: warps 2..5 <-> TMEM lane quarters (warp & 3) <-> tile rows 2q, 2q+1 =================
    const int q = warp & 3;
    const int px = lane & 15, pyl = lane >> 4;        // pixel of this lane: tile row 2q + pyl, column px
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* slab = reinterpret_cast<float*>(slabs + (size_t)(warp - 2) * 2 * SLAB);
    const float inv_c = 1.f / (float)C;
    uint32_t accp = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int tx = (int)(t % a.tiles_x);
      const int64_t r = t / a.tiles_x;
      const int ty = (int)(r % a.tiles_y), n = (int)(r / a.tiles_y);
      mbar_wait(acc_full, accp);
      tc_fence_after();
      float res[NDISP];
      // Region row 2q + s (24 accumulator columns, the same for every lane) is displacement row i = s of the lanes in tile
      // row 2q and i = s - 1 of the lanes in tile row 2q + 1.  A lane needs its columns px .. px + 8: the 24 registers are
      // shifted down by px with four select stages (8, 4, 2, 1) -- a lane-dependent register index without the
      // shared-memory bounce of the first version, whose 1300 cycles per tile shared the shared-memory pipe with the
      // tile's 64 KB of TMA writes and the MMAs' operand reads.
      const bool s8 = px & 8, s4 = px & 4, s2 = px & 2, s1 = px & 1, row0 = pyl == 0;
#pragma unroll
      for (int s = 0; s <= ND; ++s) {
        const uint32_t col = (uint32_t)((2 * q + s) * RX);
        uint32_t lo[16], v[24];
        tmem_ld16(lane_addr + col, lo);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23])
                     : "r"(lane_addr + col + 16u)
                     : "memory");
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = lo[c];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = s8 ? v[c + 8] : v[c];
#pragma unroll
        for (int c = 0; c < 12; ++c) v[c] = s4 ? v[c + 4] : v[c];
#pragma unroll
        for (int c = 0; c < 10; ++c) v[c] = s2 ? v[c + 2] : v[c];
#pragma unroll
        for (int c = 0; c < 9; ++c) v[c] = s1 ? v[c + 1] : v[c];
#pragma unroll
        for (int j = 0; j < ND; ++j) {
          const float r = __uint_as_float(v[j]) * inv_c;
          if (s < ND) res[s * ND + j] = row0 ? r : res[s * ND + j];
          if (s > 0) res[(s - 1) * ND + j] = row0 ? res[(s - 1) * ND + j] : r;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);           // tensor memory is free for the next tile's MMAs
      accp ^= 1;
      const int y = ty * TY + 2 * q + pyl, x = tx * TX + px;
      if (y < a.H && x < a.W) {
        bf16* op = a.out + (((int64_t)n * a.H + y) * a.W + x) * a.ldo;
#pragma unroll
        for (int c0 = 0; c0 < 96; c0 += 16) {
          if (c0 >= a.cout_pad) break;
          f16v o;
#pragma unroll
          for (int k = 0; k < 16; ++k) o.v[k] = (c0 + k < NDISP) ? res[c0 + k < NDISP ? c0 + k : 0] : 0.f;
          st16(op + c0, o);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Gradients.  dx[p, c] = (1/C) sum_{i,j} G[p, i*9+j] * X[p + (i-4, j-4), c]  with (G, X) = (g, x2) for dx1 and
// (g~, x1) for dx2, g~[q, i*9+j] = g[q + (i-4, j-4), (8-i)*9 + (8-j)] (corr_transpose_kernel below).
//
// Per 8 x 16 pixel tile this is ONE GEMM  D[128 px, 64 ch] = A[128, 384] * B[384, 64]:  B = the 16 x 24 pixel region of
// X exactly as its TMA box lands in shared memory (region pixel = K row, 64 channels contiguous: the MN-major operand
// layout), A = the tile's gradients scattered onto the region (A[p, q] = G[p, d(q - p)], zero where q is out of reach).
// Row i of a pixel's 9 x 9 gradient block is NINE CONSECUTIVE K entries of its A row (k = k0 + 24 i + j, j = 0..8), and
// 24 i is a multiple of 8, so every run starts at element px % 8 of a 16-byte "unit" (8 bf16): the nine values are
// shifted into place in registers and occupy two units; no other row of the pixel touches those units.

struct GradArgs {
  const bf16* g;
  int64_t ldg;
  bf16* dx;
  int64_t lddx;
  int accumulate;
  int N, H, W, tiles_x, tiles_y;
};

__device__ __forceinline__ uint64_t make_mn_desc(uint32_t saddr) {     // MN-major, 128-byte K rows, SBO = 8 rows
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((B_BYTES >> 4) & 0x3FFF) << 16;                      // LBO: next 64-channel block (there is none: N = 64)
  d |= (uint64_t)((8u * ROWB) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// A lives in TENSOR MEMORY (tcgen05.mma with a TMEM A operand: lane = row, two bf16 per 32-bit column): the builders write
// their pixel's row of A with tcgen05.st -- no shared-memory round trip, no swizzle, no proxy fence -- A is double
// buffered (2 x 192 columns next to the two 64-column accumulators = all 512 columns), so the build of tile t + 1
// overlaps the MMAs of tile t, and the MMAs read only B from shared memory (2 KB instead of 6 KB per k-step), which
// leaves room for four B stages.
//
// tcgen05.st writes the SAME columns for all 32 lanes of a warp, but a pixel's runs start at unit 3 py + px / 8, which
// takes four values c = 3 (lane / 16) + (lane % 16) / 8 in {0, 1, 3, 4} (+ 6 q) inside a warp.  Every lane therefore
// assembles a 15-unit (60-column) image of its part of the row in registers -- each run OR-ed in at each of the four
// candidate positions under a lane mask -- and the warp stores the image at a warp-uniform address.  Warp group H owns
// the units [15 H, 15 H + 15) relative to 6 q (no run straddles a multiple of 3 because c = 2 does not occur):
// displacement rows 0..4 (row 4 only for c <= 1) for H = 0, rows 4 (c >= 3)..8 for H = 1.
//
// Measured (clock64, cfg-2 shape): the build is bound by the integer pipe's THROUGHPUT (~450 SEL / LOP3 / SHF per
// pixel and tile, two builder warps per scheduler: ~1900 cycles per tile); four groups per quarter (16 warps, 12
// instead of 10 shifted rows) were 28 % slower.  MMAs (24 x 32 cycles) and the 48 KB of TMA per tile hide behind it.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTsStages = 4;
constexpr int kTsBuild = 8;                          // builder / drain warps (two per TMEM lane quarter)
constexpr int kTsThreads = 32 * (2 + kTsBuild);
constexpr uint32_t TS_D = 0, TS_A = 128, TS_ACOLS = RY * RX / 2;      // TMEM columns: accumulators, A buffers (192 each)

template <int H>
__device__ __forceinline__ void corr_grad_ts_builder(const GradArgs& a, uint64_t* a_full, uint64_t* a_free, uint64_t* acc_full,
                                                     uint64_t* acc_empty, uint32_t tmem_base, int q, int lane, int64_t ntiles) {
  constexpr int U0 = 15 * H, U1 = U0 + 15;
  constexpr int NW = (U1 - U0) * 4;                  // image words (TMEM columns): 60
  constexpr int I0 = H ? 4 : 0, I1 = H ? ND : 5;     // displacement rows this group touches
  constexpr int Q0 = H ? 4 : 0, NQ = H ? 7 : 6;      // 16-byte chunks of the pixel's gradient row it reads
  const int p = q * 32 + lane;
  const int py = p >> 4, px = p & 15;
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t img_col = TS_A + (uint32_t)(6 * q + U0) * 4u;         // first column of this group's image (buffer 0)
  const float inv_c = 1.f / (float)C;
  const bool b0 = px & 1, b1 = px & 2, b2 = px & 4;
  const int cl = 3 * (lane >> 4) + ((lane & 15) >> 3);
  const uint32_t mk[4] = {cl == 0 ? ~0u : 0u, cl == 1 ? ~0u : 0u, cl == 3 ? ~0u : 0u, cl == 4 ? ~0u : 0u};
  // The units outside [6q, 6q + 30) of this quarter's lanes are never written: zero them once in both buffers (group 0
  // the ones below, group 1 the ones above; everything inside the range is rewritten for every tile by its owner).
  {
    const uint32_t zero4[4] = {0u, 0u, 0u, 0u};
    const int u_lo = H ? 6 * q + 30 : 0, u_hi = H ? RY * RX / 8 : 6 * q;
    for (int u = u_lo; u < u_hi; ++u) {
      tmem_st4(lane_base + TS_A + (uint32_t)u * 4u, zero4);
      tmem_st4(lane_base + TS_A + TS_ACOLS + (uint32_t)u * 4u, zero4);
    }
    tmem_st_wait();
  }
  uint4 gq[NQ];
  auto fetch = [&](int64_t t) {
    const int tx = (int)(t % a.tiles_x);
    const int64_t r = t / a.tiles_x;
    const int ty = (int)(r % a.tiles_y), n = (int)(r / a.tiles_y);
    const int y = ty * TY + py, x = tx * TX + px;
    const bool ok = t < ntiles && y < a.H && x < a.W;
    const uint4* src = reinterpret_cast<const uint4*>(a.g + (((int64_t)n * a.H + y) * a.W + x) * a.ldg) + Q0;
#pragma unroll
    for (int c = 0; c < NQ; ++c) gq[c] = ok ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
  };
  uint4 accq[4];
  bf16* dp = nullptr;
  auto locate = [&](int tx, int ty, int n) {
    const int y = ty * TY + py, x = tx * TX + px;
    dp = (y < a.H && x < a.W) ? a.dx + (((int64_t)n * a.H + y) * a.W + x) * a.lddx + 32 * H : nullptr;
    if (a.accumulate && dp) {
#pragma unroll
      for (int c = 0; c < 4; ++c) accq[c] = *(reinterpret_cast<const uint4*>(dp) + c);
    }
  };
  auto drain = [&](uint32_t jt) {
    const uint32_t ab = jt & 1u;
    mbar_wait(&acc_full[ab], (jt >> 1) & 1u);
    tc_fence_after();
    uint32_t v[2][16];
    tmem_ld16(lane_base + TS_D + ab * 64u + 32u * H, v[0]);
    tmem_ld16(lane_base + TS_D + ab * 64u + 32u * H + 16u, v[1]);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[ab]);
    if (dp) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        f16v o;
#pragma unroll
        for (int k = 0; k < 16; ++k) o.v[k] = __uint_as_float(v[c][k]) * inv_c;
        if (a.accumulate) {
          const uint32_t w8[8] = {accq[2 * c].x, accq[2 * c].y, accq[2 * c].z, accq[2 * c].w,
                                  accq[2 * c + 1].x, accq[2 * c + 1].y, accq[2 * c + 1].z, accq[2 * c + 1].w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            o.v[2 * k] += __uint_as_float(w8[k] << 16);
            o.v[2 * k + 1] += __uint_as_float(w8[k] & 0xFFFF0000u);
          }
        }
        st16(dp + 16 * c, o);
      }
    }
  };
  uint32_t it = 0;
  int ptx = 0, pty = 0, pn = 0;
  fetch(blockIdx.x);
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int tx = (int)(t % a.tiles_x);
    const int64_t r = t / a.tiles_x;
    const int ty = (int)(r % a.tiles_y), n = (int)(r / a.tiles_y);
    uint32_t w[NQ * 4];
#pragma unroll
    for (int c = 0; c < NQ; ++c) { w[4 * c] = gq[c].x; w[4 * c + 1] = gq[c].y; w[4 * c + 2] = gq[c].z; w[4 * c + 3] = gq[c].w; }
    fetch(t + gridDim.x);
    if (it > 0) locate(ptx, pty, pn);
    uint32_t img[NW];
#pragma unroll
    for (int m = 0; m < NW; ++m) img[m] = 0u;
#pragma unroll
    for (int i = I0; i < I1; ++i) {
      // the nine halfwords 9i .. 9i+8 of the gradient row, right-aligned into five words
      const int s = i * ND;
      const int lw = (s >> 1) - Q0 * 4;
      uint32_t v5[5];
      if ((s & 1) == 0) {
#pragma unroll
        for (int m = 0; m < 4; ++m) v5[m] = w[lw + m];
        v5[4] = w[lw + 4] & 0xFFFFu;
      } else {
#pragma unroll
        for (int m = 0; m < 4; ++m) v5[m] = __funnelshift_r(w[lw + m], w[lw + m + 1], 16);
        v5[4] = w[lw + 4] >> 16;
      }
      // shift left by px % 8 halfwords into the 16 slots of two units
      uint32_t x6[6], y7[7], z[8];
#pragma unroll
      for (int m = 0; m < 6; ++m) {
        const uint32_t lo = m > 0 ? v5[m - 1] : 0u, hi = m < 5 ? v5[m] : 0u;
        x6[m] = b0 ? __funnelshift_l(lo, hi, 16) : hi;
      }
#pragma unroll
      for (int m = 0; m < 7; ++m) y7[m] = b1 ? (m > 0 ? x6[m - 1] : 0u) : (m < 6 ? x6[m] : 0u);
#pragma unroll
      for (int m = 0; m < 8; ++m) z[m] = b2 ? (m > 1 ? y7[m - 2] : 0u) : (m < 7 ? y7[m] : 0u);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c = ci < 2 ? ci : ci + 1;             // 0, 1, 3, 4
        const int unit = 3 * i + c - U0;                // relative to this group's image
        if (unit < 0 || unit + 1 >= U1 - U0) continue;
#pragma unroll
        for (int m = 0; m < 8; ++m) img[unit * 4 + m] |= z[m] & mk[ci];
      }
    }
    const uint32_t ab = it & 1u;
    mbar_wait(&a_free[ab], ((it >> 1) & 1u) ^ 1u);    // the MMAs of tile it - 2 have read this A buffer
    tc_fence_after();
    const uint32_t dst = lane_base + img_col + ab * TS_ACOLS;
    tmem_st32(dst, img);
    tmem_st16(dst + 32u, img + 32);
    tmem_st8(dst + 48u, img + 48);
    tmem_st4(dst + 56u, img + 56);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&a_full[ab]);
    if (it > 0) drain(it - 1);
    ptx = tx; pty = ty; pn = n;
  }
  if (it > 0) {
    locate(ptx, pty, pn);
    drain(it - 1);
  }
}

__global__ void __launch_bounds__(kTsThreads, 1)
corr_grad_ts_kernel(const __grid_constant__ CUtensorMap tmap_x, const GradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                                  // [kTsStages][384 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kTsStages * B_BYTES);
  uint64_t* b_full = bars;                       // [kTsStages] TMA -> MMA
  uint64_t* b_empty = b_full + kTsStages;        // [kTsStages] MMA -> TMA
  uint64_t* a_full = b_empty + kTsStages;        // [2] builders -> MMA
  uint64_t* a_free = a_full + 2;                 // [2] MMA -> builders
  uint64_t* acc_full = a_free + 2;               // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;            // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (int64_t)a.N * a.tiles_y * a.tiles_x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < kTsStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], kTsBuild);
      mbar_init(&a_free[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kTsBuild);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tx = (int)(t % a.tiles_x);
        const int64_t r = t / a.tiles_x;
        const int ty = (int)(r % a.tiles_y), n = (int)(r / a.tiles_y);
        mbar_wait(&b_empty[stage], phase ^ 1);
        mbar_expect_tx(&b_full[stage], B_BYTES);
        tma_load_4d(sB + (size_t)stage * B_BYTES, &tmap_x, &b_full[stage], 0, tx * TX - 4, ty * TY - 4, n);
        if (++stage == kTsStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // A from tensor memory (K-major), B MN-major (bit 16), D fp32, M = 128, N = 64
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const uint32_t ab = it & 1u;
      mbar_wait(&acc_empty[ab], ((it >> 1) & 1u) ^ 1u);
      mbar_wait(&b_full[stage], phase);
      mbar_wait(&a_full[ab], (it >> 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sb = smem_u32(sB + (size_t)stage * B_BYTES);
        const uint32_t ta = tmem_base + TS_A + ab * TS_ACOLS;
#pragma unroll 4
        for (int ks = 0; ks < RY * RX / 16; ++ks)
          umma_bf16_ts(tmem_base + TS_D + ab * 64u, ta + 8u * (uint32_t)ks, make_mn_desc(sb + (uint32_t)ks * 16u * ROWB), idesc, ks > 0);
        umma_commit(&b_empty[stage]);
        umma_commit(&a_free[ab]);
        umma_commit(&acc_full[ab]);
      }
      __syncwarp();
      if (++stage == kTsStages) { stage = 0; phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    if (warp < 6) corr_grad_ts_builder<0>(a, a_full, a_free, acc_full, acc_empty, tmem_base, q, lane, ntiles);
    else          corr_grad_ts_builder<1>(a, a_full, a_free, acc_full, acc_empty, tmem_base, q, lane, ntiles);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// g~[q, i*9+j] = g[q + (i-4, j-4), (8-i)*9 + (8-j)]  (zero outside the image; channels 81..95 zero): the gradient of the
// correlation seen from the SECOND operand's pixels.  One CTA per 8 x 16 tile: the 16 x 24 source region is staged in
// shared memory at a 49-word pixel pitch (odd: the 2-byte diagonal gathers of a warp hit 32 different banks).
constexpr int GT_PITCH = 49;                         // 32-bit words per staged pixel (96 bf16 = 48 words + 1)
constexpr int GT_THREADS = 256;

__global__ void __launch_bounds__(GT_THREADS)
corr_transpose_kernel(const bf16* __restrict__ g, int64_t ldg, bf16* __restrict__ gt, int64_t ldt, int N, int H, int W, int tiles_x,
                      int tiles_y) {
  extern __shared__ __align__(16) uint32_t st[];     // [RY*RX][GT_PITCH]
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int n = t / tiles_y;
  const int y0 = ty * TY - 4, x0 = tx * TX - 4;
  // 384 pixels x 12 chunks = 4608 chunks = 18 per thread, requested six at a time before anything is stored (a plain
  // load-store loop pays one global-memory latency per chunk)
  constexpr int PER = RY * RX * 12 / GT_THREADS;       // 18
  constexpr int GB = 9;                                // chunks requested per round trip (two round trips per tile)
  static_assert(PER * GT_THREADS == RY * RX * 12 && PER % GB == 0, "chunk split");
#pragma unroll 1
  for (int b = 0; b < PER; b += GB) {
    uint4 u[GB];
#pragma unroll
    for (int k = 0; k < GB; ++k) {
      const int e = threadIdx.x + (b + k) * GT_THREADS;
      const int rp = e / 12, c = e - rp * 12;
      const int ry = rp / RX, rx = rp - ry * RX;
      const int y = y0 + ry, x = x0 + rx;
      u[k] = make_uint4(0, 0, 0, 0);
      if (y >= 0 && y < H && x >= 0 && x < W) u[k] = __ldg(reinterpret_cast<const uint4*>(g + (((int64_t)n * H + y) * W + x) * ldg) + c);
    }
#pragma unroll
    for (int k = 0; k < GB; ++k) {
      const int e = threadIdx.x + (b + k) * GT_THREADS;
      const int rp = e / 12, c = e - rp * 12;
      uint32_t* dst = st + rp * GT_PITCH + c * 4;
      dst[0] = u[k].x; dst[1] = u[k].y; dst[2] = u[k].z; dst[3] = u[k].w;
    }
  }
  __syncthreads();
  const int p = threadIdx.x & 127, half = threadIdx.x >> 7;        // pixel of the tile, which 48 output channels
  const int py = p >> 4, px = p & 15;
  const int y = ty * TY + py, x = tx * TX + px;
  const uint16_t* sh = reinterpret_cast<const uint16_t*>(st);
  uint32_t o[24];
  // (the two halves are whole warps: with `half` resolved per branch every displacement index below is a compile-time
  //  constant -- the run-time form divided d by 9 for each of the 48 gathers)
  const uint16_t* pix = sh + (py * RX + px) * (GT_PITCH * 2);
  auto gather = [&](auto HALF) {
    constexpr int hbase = decltype(HALF)::value * 48;
#pragma unroll
    for (int k = 0; k < 24; ++k) {
      uint32_t word = 0;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int d = hbase + 2 * k + hf;                            // output channel i*9+j (d >= 81: zero padding)
        uint16_t v = 0;
        if (d < NDISP) {
          const int i = d / ND, j = d - i * ND;
          v = pix[(i * RX + j) * (GT_PITCH * 2) + (NDISP - 1 - d)];
        }
        word |= (uint32_t)v << (16 * hf);
      }
      o[k] = word;
    }
  };
  if (half == 0) gather(std::integral_constant<int, 0>{}); else gather(std::integral_constant<int, 1>{});
  if (y < H && x < W) {
    uint4* dst = reinterpret_cast<uint4*>(gt + (((int64_t)n * H + y) * W + x) * ldt + half * 48);
#pragma unroll
    for (int k = 0; k < 6; ++k) dst[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
  }
}

}  // namespace

namespace nv {

bool corr_fwd_tc_supported(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* out, int64_t ldo, int H, int W,
                           int cout_pad) {
  if ((ld1 & 7) || (ld2 & 7) || !aligned(x1, 16) || !aligned(x2, 16)) return false;
  if (cout_pad != 96 || (ldo & 15) || !aligned(out, 32)) return false;       // 32-byte stores of whole 16-channel groups
  if (H < 8 || W < 16) return false;
  return encode_fn() != nullptr;
}

int corr_fwd_tc(const void* x1, int64_t ld1, const void* x2, int64_t ld2, void* out, int64_t ldo, int N, int H, int W,
                int cout_pad, cudaStream_t s) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return NERVECL_EUNSUPPORTED;
  CUtensorMap ta, tb;
  auto encode = [&](CUtensorMap* m, const void* base, int64_t ld, int bx, int by) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!encode(&ta, x1, ld1, TX, TY) || !encode(&tb, x2, ld2, RX, RY)) return NERVECL_EUNSUPPORTED;
  FwdArgs a;
  a.out = (bf16*)out; a.ldo = ldo; a.N = N; a.H = H; a.W = W; a.cout_pad = cout_pad;
  a.tiles_x = (W + TX - 1) / TX;
  a.tiles_y = (H + TY - 1) / TY;
  const int64_t ntiles = (int64_t)N * a.tiles_x * a.tiles_y;
  const size_t smem = 1024 + (size_t)kStages * STAGE_BYTES + (size_t)kEpiWarps * 2 * SLAB + 16 * sizeof(uint64_t);
  cudaError_t e = cudaFuncSetAttribute(corr_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const unsigned grid = (unsigned)imin(ntiles, sm_count());
  corr_fwd_tc_kernel<<<grid, kThreads, smem, s>>>(ta, tb, a);
  return launch_status();
}

}  // namespace nv

namespace nv {

bool corr_bwd_tc_supported(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* g, int64_t ldg, const void* dx1,
                           int64_t lddx1, const void* dx2, int64_t lddx2, const void* ws, int64_t ws_bytes, int N, int H, int W) {
  if ((ld1 & 7) || (ld2 & 7) || !aligned(x1, 16) || !aligned(x2, 16)) return false;
  if (ldg < 88 || (ldg & 7) || !aligned(g, 16)) return false;                  // eleven 16-byte loads per pixel
  if ((lddx1 & 15) || (lddx2 & 15) || !aligned(dx1, 32) || !aligned(dx2, 32)) return false;
  if (!ws || !aligned(ws, 16) || ws_bytes < (int64_t)N * H * W * 96 * 2) return false;
  if (H < 8 || W < 16) return false;
  return encode_fn() != nullptr;
}

static int corr_grad_tc(const void* X, int64_t ldX, const void* g, int64_t ldg, void* dx, int64_t lddx, int accumulate, int N,
                        int H, int W, cudaStream_t s) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return NERVECL_EUNSUPPORTED;
  CUtensorMap tx;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ldX * 2, (cuuint64_t)W * ldX * 2, (cuuint64_t)H * W * ldX * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)RX, (cuuint32_t)RY, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(X), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return NERVECL_EUNSUPPORTED;
  GradArgs a;
  a.g = (const bf16*)g; a.ldg = ldg; a.dx = (bf16*)dx; a.lddx = lddx; a.accumulate = accumulate;
  a.N = N; a.H = H; a.W = W;
  a.tiles_x = (W + TX - 1) / TX;
  a.tiles_y = (H + TY - 1) / TY;
  const int64_t ntiles = (int64_t)N * a.tiles_x * a.tiles_y;
  const size_t smem = 1024 + (size_t)kTsStages * B_BYTES + 24 * sizeof(uint64_t);
  cudaError_t e = cudaFuncSetAttribute(corr_grad_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  corr_grad_ts_kernel<<<(unsigned)imin(ntiles, sm_count()), kTsThreads, smem, s>>>(tx, a);
  return launch_status();
}

int corr_bwd_tc(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* g, int64_t ldg, void* dx1, int64_t lddx1,
                int acc1, void* dx2, int64_t lddx2, int acc2, void* ws, int N, int H, int W, cudaStream_t s) {
  int rc = corr_grad_tc(x2, ld2, g, ldg, dx1, lddx1, acc1, N, H, W, s);
  if (rc) return rc;
  const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
  const size_t smem = (size_t)RY * RX * GT_PITCH * 4;
  cudaError_t e = cudaFuncSetAttribute(corr_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  corr_transpose_kernel<<<(unsigned)((int64_t)N * tiles_x * tiles_y), GT_THREADS, smem, s>>>((const bf16*)g, ldg, (bf16*)ws, 96, N, H, W,
                                                                                       tiles_x, tiles_y);
  rc = launch_status();
  if (rc) return rc;
  return corr_grad_tc(x1, ld1, ws, 96, dx2, lddx2, acc2, N, H, W, s);
}

}  // namespace nv
