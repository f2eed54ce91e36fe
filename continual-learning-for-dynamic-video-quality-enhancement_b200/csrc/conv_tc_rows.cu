// Row-streaming 3x3 convolution for sm_100a: tcgen05.mma + TMEM accumulator ring + TMA, bf16 in / fp32 accumulate.
//
// Why not the per-tap kernel (conv_tc.cu): there every (tap, chunk) k-block re-fetches its shifted 128-pixel A
// tile through TMA -- 9x the L2 traffic of the input -- and with N = Cout = 32 each SS-mode MMA reads 5 KB of
// shared memory per 16 tensor cycles, so the kernel sat at ~10 % tensor-pipe activity (profiles/r01a_summary.md).
//
// Here a persistent CTA owns an equal share of all (image, 128-pixel column strip, row) output rows -- a contiguous
// range of the linearised row space, cut into work items (image n, strip, first row, row count) wherever it
// crosses a strip boundary -- and streams each item's rows+2 input rows top to bottom:
//   * TMA brings each halo'd input row {64 ch, 130 px} per 64-channel chunk into shared memory ONCE
//     (128B swizzle, out-of-image pixels / rows / channels are zero-filled = conv padding); the ring is
//     chunk-granular, so wide inputs still get several stages;
//   * the horizontal taps kx = 0,1,2 are three K-major smem descriptors into that same row, offset by one
//     pixel (128 B): the 128B swizzle is a function of the absolute smem address, so an operand may start at
//     any pixel of a 1024-aligned buffer;
//   * the vertical taps are merged into N: input row r contributes W[ky=2] to output row r-1, W[ky=1] to r and
//     W[ky=0] to r+1, whose accumulators occupy ADJACENT column slots of a TMEM ring, so one tcgen05.mma with
//     N = 3*NOUT (A = row r, B = [W(ky=2) | W(ky=1) | W(ky=0)] rows) updates all three: 4 KB of A per 3*NOUT
//     columns instead of per NOUT columns.  (3*NOUT > 256 falls back to one MMA per ky.)
//   * all weights stay resident in shared memory; output row o is final after input row o+1 and is drained by
//     8 epilogue warps (fused bias / ReLU / scale / residual / accumulate / ReLU-mask / store), which also
//     re-zero the accumulator slot (every MMA accumulates) and prefetch the ReLU-mask operand rows ahead;
//   * an optional SECOND input tensor x2 is a virtual channel concat ([x | x2], no copy); with x2_center its
//     channels only see the centre tap -- a fused 1x1 branch (the dense block's LFF data gradient).
// Output channels beyond what fits (smem for weights, 128 TMEM columns per row) are split over blockIdx.y.
#include "conv_internal.cuh"
#include "conv_tc_epilogue.cuh"
#include "tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <type_traits>

using namespace nv;
using namespace nv::tc;

namespace {

// profiling-only work-skipping switches exist only in -DNERVECL_TUNING builds: the shipped library cannot time a
// kernel that does no work
#ifdef NERVECL_TUNING
#define ROWS_DBG(a, bit) (((a).dbg & (bit)) != 0)
#else
#define ROWS_DBG(a, bit) false
#endif

constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);     // warp 0 = TMA, warp 1 = MMA issuer, warps 2..9 = epilogue
constexpr int KC = 64;                             // channels per chunk = one 128-byte swizzle span
constexpr uint32_t ROWB = KC * 2;                  // bytes per pixel row of a chunk
constexpr int PXB = BM + 2;                        // staged pixels per input row (one halo pixel each side)
constexpr uint32_t CHUNK_BYTES = (PXB * ROWB + 1023u) & ~1023u;
constexpr int kMaxSlots = 32;
constexpr int kMaxStages = 12;
constexpr size_t kSmemBudget = 226 * 1024;

struct RowArgs : EpiArgs {
  int N, H, W;
  int NOUT;         // output channels per CTA (multiple of 16, <= 128)
  int nchunks;      // 64-channel chunks of x
  int ksteps_last;  // 16-channel k-steps of the last x chunk
  int nchunks2;     // 64-channel chunks of x2 (0: no second input)
  int ksteps2_last;
  int x2_center;    // x2 channels only have a centre tap
  int x_center;     // 1x1 convolution: x only has a centre tap (its ky = 0 / 2 weight blocks are zero-filled)
  int strips, R, segs;
  int cps;          // chunks per ring stage (a stage is filled / released as a unit)
  int stages;       // ring depth in stages
  int slots;        // TMEM accumulator ring depth
  int merged;       // vertical taps merged into N
  int fast;         // lean epilogues -> bf16: 1: relu?(acc + bias?); 2: alpha * acc gated by the ReLU mask; 3: alpha * acc + res
  int dbg;          // NERVECL_ROWS_DBG bits (profiling only): 1 no MMAs, 2 no row loads, 4 no epilogue work
};

// weight tile index of (kx, chunk): x chunks first, then x2 chunks (one tile per chunk when centre-only)
__device__ __forceinline__ int wtile_index(const RowArgs& a, int kx, int c) {
  if (c < a.nchunks) return a.x_center ? c : kx * a.nchunks + c;
  const int c2 = c - a.nchunks;
  return (a.x_center ? 1 : 3) * a.nchunks + (a.x2_center ? c2 : kx * a.nchunks2 + c2);
}

// One 16-channel chunk of a lean epilogue (bf16 out): kind 1: relu?(acc + bias); 2: alpha * acc where mask > 0;
// 3: alpha * acc + residual.  e0/e1 hold the chunk's 16 mask / residual values.
__device__ __forceinline__ void lean_emit(int kind, const uint32_t (&v)[16], const uint4& e0, const uint4& e1, bf16* op,
                                          const float* bias, bool relu, float alpha, bool v256, bool add_res = false) {
  float f[16];
  if (kind == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      f[4 * i] = __uint_as_float(v[4 * i]) + b.x;
      f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b.y;
      f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b.z;
      f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b.w;
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (add_res) {
      const uint32_t w[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[2 * j] += __uint_as_float(w[j] << 16);
        f[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
      }
    }
  } else {
    const uint32_t w[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xFFFF0000u);
      if (kind == 2) {
        f[2 * j] = lo > 0.f ? alpha * __uint_as_float(v[2 * j]) : 0.f;
        f[2 * j + 1] = hi > 0.f ? alpha * __uint_as_float(v[2 * j + 1]) : 0.f;
      } else {
        f[2 * j] = fmaf(alpha, __uint_as_float(v[2 * j]), lo);
        f[2 * j + 1] = fmaf(alpha, __uint_as_float(v[2 * j + 1]), hi);
      }
    }
  }
  store16(op, f, v256);
}

// The CTA's share of the linearised output rows ((n * strips + strip) * H + y), split evenly (to within one row)
// over the CTAs of a channel group; next() yields the items of that share in order.  Every role (TMA producer,
// MMA issuer, epilogue) walks the same sequence.
struct ItemIter {
  int pos, end;                       // (N * strips * H < 2^31: checked by conv_rows_supported)
  __device__ __forceinline__ ItemIter(const RowArgs& a) {
    const int64_t total = (int64_t)a.N * a.strips * a.H;
    pos = (int)(total * blockIdx.x / gridDim.x);
    end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  }
  __device__ __forceinline__ bool next(const RowArgs& a, int& n, int& strip, int& y0, int& rows) {
    if (pos >= end) return false;
    const int unit = pos / a.H;
    y0 = pos - unit * a.H;
    rows = min(end - pos, a.H - y0);
    n = unit / a.strips;
    strip = unit - n * a.strips;
    pos += rows;
    return true;
  }
};

template <typename OutT, bool PF>
__global__ void __launch_bounds__(kThreads, 1)
conv_rows_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_x2,
                 const __grid_constant__ CUtensorMap tmap_w, const RowArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t blk_bytes = (uint32_t)a.NOUT * ROWB;          // one ky block of one (kx, chunk) weight tile
  const uint32_t wtile_bytes = 3u * blk_bytes;
  const int nct = a.nchunks + a.nchunks2;
  const int ntiles = (a.x_center ? 1 : 3) * a.nchunks + (a.x2_center ? 1 : 3) * a.nchunks2;
  const uint32_t w_bytes = (uint32_t)ntiles * wtile_bytes;
  uint8_t* w_smem = smem;
  uint8_t* ring = smem + w_bytes;
  const uint32_t stage_bytes = (uint32_t)a.cps * CHUNK_BYTES;
  const int ngrp = nct / a.cps;                     // stages per input row
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)a.stages * stage_bytes);
  uint64_t* ch_full = bars;                         // [stages]  TMA -> MMA
  uint64_t* ch_empty = ch_full + a.stages;          // [stages]  MMA -> TMA
  uint64_t* acc_full = ch_empty + a.stages;         // [slots]   MMA -> epilogue
  uint64_t* acc_empty = acc_full + a.slots;         // [slots]   epilogue -> MMA
  uint64_t* w_bar = acc_empty + a.slots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.y;                       // output-channel group
  const int S = a.slots;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    if (a.nchunks2) prefetch_tmap(&tmap_x2);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&ch_full[s], 1);
      mbar_init(&ch_empty[s], 1);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kEpiWarps);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_expect_tx(w_bar, w_bytes);
      for (int c = 0; c < nct; ++c)
        for (int kx = 0; kx < 3; ++kx) {
          if ((c >= a.nchunks ? a.x2_center : a.x_center) && kx != 1) continue;
          for (int b = 0; b < 3; ++b) {   // block b holds ky = 2 - b; a 1x1 filter is tap 0, its other blocks are
            const int tap = a.x_center ? (b == 1 ? 0 : 1) : (2 - b) * 3 + kx;    // out of range = zero fill
            tma_load_3d(w_smem + (size_t)wtile_index(a, kx, c) * wtile_bytes + (size_t)b * blk_bytes, &tmap_w, w_bar,
                        c * KC, grp * a.NOUT, tap);
          }
        }
      int stage = 0;
      uint32_t phase = 0;
      ItemIter it(a);
      int n, strip, y0, rows;
      while (it.next(a, n, strip, y0, rows)) {
        const int x0 = strip * BM;
        for (int ri = 0; ri < rows + 2; ++ri) {
          for (int gi = 0; gi < ngrp; ++gi) {
            mbar_wait(&ch_empty[stage], phase ^ 1);
            if (ROWS_DBG(a, 2)) {
              mbar_arrive(&ch_full[stage]);
            } else {
              mbar_expect_tx(&ch_full[stage], (uint32_t)a.cps * (PXB * ROWB));
              for (int cc = 0; cc < a.cps; ++cc) {
                const int c = gi * a.cps + cc;
                uint8_t* dst = ring + (size_t)stage * stage_bytes + (size_t)cc * CHUNK_BYTES;
                if (c < a.nchunks) tma_load_4d(dst, &tmap_x, &ch_full[stage], c * KC, x0 - 1, y0 - 1 + ri, n);
                else tma_load_4d(dst, &tmap_x2, &ch_full[stage], (c - a.nchunks) * KC, x0 - 1, y0 - 1 + ri, n);
              }
            }
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // All 32 lanes run the (warp-uniform) control flow; one elected lane issues the MMAs and commits, so
    // ptxas keeps descriptors in uniform registers instead of emitting a per-MMA lane-election loop.  A
    // dependent scalar instruction costs ~5 cycles and an N=96 MMA only 48, so the per-MMA work is two adds.
    // Accumulators are zeroed by the epilogue warps: every MMA accumulates.
    {
      const uint32_t w_lo = (smem_u32(w_smem) & 0x3FFFFu) >> 4;
      const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);     // | LBO field (unused) = 1
      const uint32_t blk_lo = blk_bytes >> 4, wtile_lo = wtile_bytes >> 4;
      const uint64_t desc_hi = make_kmajor_desc(0, ROWB) & 0xFFFFFFFF00000000ull;
      const uint32_t id1 = make_idesc_bf16((uint32_t)a.NOUT), id2 = make_idesc_bf16(2u * a.NOUT),
                     id3 = make_idesc_bf16(3u * a.NOUT);
      int stage = 0;
      uint32_t phase = 0;
      int jbase = 0;                      // output rows issued so far by this CTA (running index -> TMEM slot)
      bool ready = false;                 // early, non-blocking probe of the next chunk's barrier succeeded
      bool acc_ready = false;             // same for the accumulator slot the next input row starts
      // "output row final" commit of the previous input row, issued behind the first MMAs of the next one: the
      // issuing thread is blocked on the MMA queue there anyway, so the commit's ~125 cycles leave the row-to-row
      // critical path (the accumulator ring has >= 4 slots: the slot is not needed again before that)
      int pend = -1;
      // same for the "stage consumed" commit (behind the second MMA group) -- only with a deep ring: with the two
      // whole-row stages of the K >= 160 layers the refill latency is exposed and deferring costs 25-40 %
      int pend_stage = -1;
      const bool defer_stage = a.stages >= 4;
      mbar_wait(w_bar, 0);
      ItemIter it(a);
      int n_, strip_, y0_, rows;
      while (it.next(a, n_, strip_, y0_, rows)) {
        int s2 = jbase % S;               // slot of output row oi = ri (the row this input row starts)
        uint32_t r2 = (uint32_t)(jbase / S);   // how many times that slot has been used before
        acc_ready = false;                     // (the probe at the end of the previous item was for another slot)
        for (int ri = 0; ri < rows + 2; ++ri) {
          // input row ri feeds output rows oi = ri - ky (0 <= oi < rows); weight block b = 2 - ky <-> oi = ri - 2 + b
          const int blo = max(0, 2 - ri), bhi = min(2, rows + 1 - ri);
          const int s1 = s2 ? s2 - 1 : S - 1, s0 = s1 ? s1 - 1 : S - 1;
          // group adjacent blocks whose accumulator slots are adjacent columns into one MMA (merged mode)
          uint32_t gb0, gb1 = 0, gb2 = 0, gc0, gc1 = 0, gc2 = 0, gi0, gi1 = 0, gi2 = 0;
          int ng;
          if (a.merged && blo == 0 && bhi == 2 && s2 >= 2) {     // interior row, no ring wrap: one N = 3*NOUT MMA
            gb0 = 0; gc0 = (uint32_t)(s0 * a.NOUT); gi0 = id3; ng = 1;
          } else {
            const bool m01 = a.merged && blo == 0 && bhi >= 1 && s1 == s0 + 1;
            const bool m12 = a.merged && blo <= 1 && bhi == 2 && s2 == s1 + 1;
            auto sbf = [&](int b) { return b == 0 ? s0 : b == 1 ? s1 : s2; };
            int b = blo;
            int nb = 1;
            if (b == 0 && m01) nb = m12 ? 3 : 2; else if (b == 1 && m12) nb = 2;
            gb0 = (uint32_t)b * blk_lo; gc0 = (uint32_t)(sbf(b) * a.NOUT); gi0 = nb == 3 ? id3 : nb == 2 ? id2 : id1;
            ng = 1;
            b += nb;
            if (b <= bhi) {
              nb = (b == 1 && m12) ? 2 : 1;
              gb1 = (uint32_t)b * blk_lo; gc1 = (uint32_t)(sbf(b) * a.NOUT); gi1 = nb == 2 ? id2 : id1;
              ng = 2;
              b += nb;
              if (b <= bhi) { gb2 = (uint32_t)b * blk_lo; gc2 = (uint32_t)(sbf(b) * a.NOUT); gi2 = id1; ng = 3; }
            }
          }
          if (bhi == 2 && !acc_ready) mbar_wait(&acc_empty[s2], r2 & 1u);   // slot of the new output row is drained + zeroed
          acc_ready = false;
          for (int gi = 0; gi < ngrp; ++gi) {
            if (!ready) mbar_wait(&ch_full[stage], phase);
            tc_fence_after();
            int nstage = stage + 1;
            uint32_t nphase = phase;
            if (nstage == a.stages) { nstage = 0; nphase ^= 1; }
            const int commit_pend = gi == 0 ? pend : -1;
            if (elect_one()) {
              bool first = commit_pend >= 0;
              int second_left = pend_stage >= 0 ? 2 : 0;          // k-loops until the deferred stage commit goes out
              if (ROWS_DBG(a, 1) && first) { umma_commit(&acc_full[commit_pend]); first = false; }
              if (!ROWS_DBG(a, 1)) {
                for (int cc = 0; cc < a.cps; ++cc) {
                  const int c = gi * a.cps + cc;
                  const uint32_t a_lo = ring_lo + (uint32_t)stage * (stage_bytes >> 4) + (uint32_t)cc * (CHUNK_BYTES >> 4);
                  const bool second = c >= a.nchunks;
                  const int ks = (c == a.nchunks - 1) ? a.ksteps_last : (c == nct - 1 && second) ? a.ksteps2_last : KC / 16;
                  const bool ctr = second ? a.x2_center != 0 : a.x_center != 0;   // centre tap only
                  // weight tile of (kx, c): tile0 + kx * tstride
                  const uint32_t tile0 = (uint32_t)(second ? (a.x_center ? 1 : 3) * a.nchunks + (c - a.nchunks) : c);
                  const uint32_t tstride = (uint32_t)(ctr ? 0 : second ? a.nchunks2 : a.nchunks);
                  for (int g = 0; g < ng; ++g) {
                    const uint32_t d = tmem_base + (g == 0 ? gc0 : g == 1 ? gc1 : gc2);
                    const uint32_t gb = w_lo + (g == 0 ? gb0 : g == 1 ? gb1 : gb2) + tile0 * wtile_lo;
                    const uint32_t id = g == 0 ? gi0 : g == 1 ? gi1 : gi2;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                      if (ctr && kx != 1) continue;
                      const uint32_t al = a_lo + (uint32_t)kx * (ROWB >> 4);
                      const uint32_t bl = (1u << 16) | (gb + (uint32_t)kx * tstride * wtile_lo);
                      if (ks == KC / 16) {
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k)
                          umma_bf16_acc(d, desc_hi | (uint64_t)(al + 2u * k), desc_hi | (uint64_t)(bl + 2u * k), id);
                      } else {
#pragma unroll
                        for (int k = 0; k < KC / 16 - 1; ++k)
                          if (k < ks) umma_bf16_acc(d, desc_hi | (uint64_t)(al + 2u * k), desc_hi | (uint64_t)(bl + 2u * k), id);
                      }
                      if (first) { umma_commit(&acc_full[commit_pend]); first = false; }
                      if (second_left && --second_left == 0) umma_commit(&ch_empty[pend_stage]);
                    }
                  }
                }
              }
              if (first) umma_commit(&acc_full[commit_pend]);                 // (no MMA was issued in this stage)
              if (second_left) umma_commit(&ch_empty[pend_stage]);            // (fewer than two MMA groups in this stage)
              if (!defer_stage) umma_commit(&ch_empty[stage]);                // stage consumed when these MMAs retire
            }
            pend_stage = defer_stage ? stage : -1;
            if (gi == 0) pend = -1;
            // probe the next stage's barrier (and, after the row's last stage, the accumulator slot of the next
            // row) now: the latencies overlap each other and the MMAs just issued
            ready = mbar_try_wait(&ch_full[nstage], nphase);
            if (gi == ngrp - 1) {
              const int ns2 = s2 + 1 == S ? 0 : s2 + 1;
              const uint32_t nr2 = s2 + 1 == S ? r2 + 1 : r2;
              acc_ready = mbar_try_wait(&acc_empty[ns2], nr2 & 1u);
            }
            stage = nstage;
            phase = nphase;
          }
          if (ri >= 2) pend = s0;                                             // output row ri-2 (slot s0) is final
          if (++s2 == S) { s2 = 0; ++r2; }
        }
        jbase += rows;
      }
      if (elect_one()) {
        if (pend_stage >= 0) umma_commit(&ch_empty[pend_stage]);
        if (pend >= 0) umma_commit(&acc_full[pend]);
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue (warps 2..9) =================
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;                      // which half of the 16-column chunks
    const int row = q * 32 + lane;                         // pixel within the strip
    const int c_lo = grp * a.NOUT;
    const int nch = (min(a.Cout, c_lo + a.NOUT) - c_lo + 15) >> 4;
    const int nch_all = a.NOUT >> 4;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    // zero the whole accumulator ring once, then publish every slot as empty
    for (int col = part * 16; col < S * a.NOUT; col += 32) tmem_st16_zero(lane_addr + (uint32_t)col);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int sl = 0; sl < S; ++sl) mbar_arrive(&acc_empty[sl]);
    // ReLU-mask prefetch: with NOUT <= 32 a thread owns ONE 16-channel chunk per row; its mask values for the
    // next 4 rows of the item are kept in flight so the epilogue is not a chain of exposed HBM latencies.
    const int cm = c_lo + part * 16;                       // absolute first channel of that chunk
    // (PF kernels are only launched with a mask, no mask_sub and NOUT <= 32)
    const bool pf = PF && part < nch && cm >= a.mask_c0 && cm + 16 <= a.Cout && !ROWS_DBG(a, 4);
    uint4 mqa[4], mqb[4];            // [row slot]: channels cm..cm+7 and cm+8..cm+15
#pragma unroll
    for (int d = 0; d < 4; ++d) mqa[d] = mqb[d] = make_uint4(0, 0, 0, 0);
    int slot = 0;
    uint32_t par = 0;
    if (!PF && a.fast && nch_all > 4 && sizeof(OutT) == 2) {
      // ---- lean epilogue, 80..128 output channels per CTA: a thread owns chunks part, part+2, part+4, part+6; all
      //      mask / residual operands of the row are requested before the accumulator wait, the chunks are drained
      //      two at a time (the generic path made these launches epilogue-bound: 0.95 ms vs 0.26 ms without it)
      const int kind = a.fast;
      const bool relu = a.relu != 0;
      const float alpha = a.alpha;
      const bool vo = a.v256_out != 0, vi = a.v256_in != 0;
      ItemIter it(a);
      int n, strip, y0, rows;
      while (it.next(a, n, strip, y0, rows)) {
        const int x = strip * BM + row;
        const bool valid = x < a.W && !ROWS_DBG(a, 4);
        const int64_t p0 = ((int64_t)n * a.H + y0) * a.W + x;
        const int cf = c_lo + part * 16;                         // first channel of this thread's first chunk
        bf16* op = reinterpret_cast<bf16*>(a.out) + p0 * a.ldo + cf;
        const int64_t ostride = (int64_t)a.W * a.ldo;
        const bool has_e = kind >= 2 || a.res != nullptr;       // kind 1 may carry a residual (added after the ReLU)
        const bf16* ep = kind != 2 ? a.res + p0 * a.ldres + cf : a.mask + p0 * a.ldmask + cf;
        const int64_t estride = (int64_t)a.W * (kind != 2 ? a.ldres : a.ldmask);
        for (int oi = 0; oi < rows; ++oi) {
          uint4 e[4][2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            e[i][0] = e[i][1] = make_uint4(0, 0, 0, 0);
            if (has_e && valid && part + 2 * i < nch_all) {
              load32B(ep + i * 32, e[i][0], e[i][1], vi);
            }
          }
          ep += estride;
          mbar_wait(&acc_full[slot], par);
          tc_fence_after();
          const uint32_t tcol = lane_addr + (uint32_t)(slot * a.NOUT + part * 16);
#pragma unroll
          for (int i = 0; i < 4; i += 2) {
            const bool ha = part + 2 * i < nch_all, hb = part + 2 * i + 2 < nch_all;
            uint32_t va[16], vb[16];
            if (ha) tmem_ld16(tcol + (uint32_t)(i * 32), va);
            if (hb) tmem_ld16(tcol + (uint32_t)(i * 32 + 32), vb);
            tmem_ld_wait();
            if (ha) tmem_st16_zero(tcol + (uint32_t)(i * 32));                  // re-arm the slot for its next output row
            if (hb) tmem_st16_zero(tcol + (uint32_t)(i * 32 + 32));
            if (valid) {
              if (ha) lean_emit(kind, va, e[i][0], e[i][1], op + i * 32, a.bias ? a.bias + cf + i * 32 : nullptr, relu, alpha, vo, kind == 1 && has_e);
              if (hb) lean_emit(kind, vb, e[i + 1][0], e[i + 1][1], op + i * 32 + 32,
                                a.bias ? a.bias + cf + i * 32 + 32 : nullptr, relu, alpha, vo, kind == 1 && has_e);
            }
          }
          op += ostride;
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[slot]);
          if (++slot == S) { slot = 0; par ^= 1u; }
        }
      }
    } else if (a.fast && sizeof(OutT) == 2) {
      // ---- lean epilogues (the generic one below is ~300 dependent instructions per warp and row, which made
      //      the epilogue the bottleneck of every small-K launch): whole 16-channel chunks, bf16 output, bias
      //      kept in registers, row pointers advanced instead of recomputed.  A thread owns chunk `part` and,
      //      for NOUT = 64, chunk part + 2.
      const bool has0 = part < nch_all, two = !PF && part + 2 < nch_all;     // (PF kernels: NOUT <= 32)
      const int ch0 = c_lo + part * 16, ch1 = ch0 + 32;
      // (PF kernels always have a mask: fast == 1 never runs there, and its 32 bias registers make room for the
      //  16 running column sums of the fused bias gradient)
      float bz0[16], bz1[16], cs[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        bz0[j] = (!PF && a.fast == 1 && a.bias && has0) ? __ldg(a.bias + ch0 + j) : 0.f;
        bz1[j] = (!PF && a.fast == 1 && a.bias && two) ? __ldg(a.bias + ch1 + j) : 0.f;
        cs[j] = 0.f;
      }
      const bool want_cs = PF && a.colsum != nullptr;
      const bool relu = a.relu != 0;
      const float alpha = a.alpha;
      const bool vo = a.v256_out != 0, vi = a.v256_in != 0;
      ItemIter it(a);
      int n, strip, y0, rows;
      while (it.next(a, n, strip, y0, rows)) {
        const int x = strip * BM + row;
        const bool valid = x < a.W;
        const int64_t p0 = ((int64_t)n * a.H + y0) * a.W + x;
        bf16* op = reinterpret_cast<bf16*>(a.out) + p0 * a.ldo + ch0;
        const int64_t ostride = (int64_t)a.W * a.ldo;
        const bf16* mp = a.mask + p0 * a.ldmask + ch0;
        const int64_t mstride = (int64_t)a.W * a.ldmask;
        // second epilogue operand when it is not prefetched: the mask (fast 2, NOUT = 64) or the residual (fast 3)
        const bool res1 = !PF && a.fast == 1 && a.res != nullptr;   // bias / ReLU epilogue with a residual added after
        const bf16* ep = (a.fast == 3 || res1) ? a.res + p0 * a.ldres + ch0 : mp;
        const int64_t estride = (a.fast == 3 || res1) ? (int64_t)a.W * a.ldres : mstride;
        if (PF && valid && has0) {
#pragma unroll
          for (int d = 0; d < 4; ++d)
            if (d < rows) {
              load32B(mp + d * mstride, mqa[d], mqb[d], vi);
            }
        }
        for (int oi4 = 0; oi4 < rows; oi4 += 4) {
#pragma unroll
          for (int D = 0; D < 4; ++D) {
            const int oi = oi4 + D;
            if (oi >= rows) break;
            uint4 m0 = mqa[D], m1 = mqb[D];
            if (PF && valid && has0 && oi + 4 < rows) {
              load32B(mp + (int64_t)(oi + 4) * mstride, mqa[D], mqb[D], vi);
            }
            uint4 e0 = m0, e1 = m1, e2 = m0, e3 = m1;
            if (!PF && (a.fast >= 2 || res1) && valid && has0) {       // issued before the wait: latency overlaps it
              load32B(ep, e0, e1, vi);
              if (two) load32B(ep + 32, e2, e3, vi);
            }
            ep += estride;
            mbar_wait(&acc_full[slot], par);
            tc_fence_after();
            const uint32_t tcol = lane_addr + (uint32_t)(slot * a.NOUT + part * 16);
            uint32_t v0[16], v1[16];
            if (has0) tmem_ld16(tcol, v0);
            if (two) tmem_ld16(tcol + 32u, v1);
            tmem_ld_wait();
            if (has0) tmem_st16_zero(tcol);                          // re-arm the slot for its next output row
            if (two) tmem_st16_zero(tcol + 32u);
            if (valid && has0 && !ROWS_DBG(a, 4)) {
              float f[16];
              if (!PF && a.fast == 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  f[j] = __uint_as_float(v0[j]) + bz0[j];
                  if (relu) f[j] = fmaxf(f[j], 0.f);
                }
                if (res1) {
                  const uint32_t rw[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    f[2 * j] += __uint_as_float(rw[j] << 16);
                    f[2 * j + 1] += __uint_as_float(rw[j] & 0xFFFF0000u);
                  }
                }
                store16(op, f, vo);
                if (two) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    f[j] = __uint_as_float(v1[j]) + bz1[j];
                    if (relu) f[j] = fmaxf(f[j], 0.f);
                  }
                  if (res1) {
                    const uint32_t sw[8] = {e2.x, e2.y, e2.z, e2.w, e3.x, e3.y, e3.z, e3.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      f[2 * j] += __uint_as_float(sw[j] << 16);
                      f[2 * j + 1] += __uint_as_float(sw[j] & 0xFFFF0000u);
                    }
                  }
                  store16(op + 32, f, vo);
                }
              } else if (a.fast == 2) {                              // alpha * acc where mask > 0
                const uint32_t mw[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float ml = __uint_as_float(mw[j] << 16), mh = __uint_as_float(mw[j] & 0xFFFF0000u);
                  f[2 * j] = ml > 0.f ? alpha * __uint_as_float(v0[2 * j]) : 0.f;
                  f[2 * j + 1] = mh > 0.f ? alpha * __uint_as_float(v0[2 * j + 1]) : 0.f;
                }
                store16(op, f, vo);
                if (want_cs) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) cs[j] += f[j];
                }
                if (two) {
                  const uint32_t nw[8] = {e2.x, e2.y, e2.z, e2.w, e3.x, e3.y, e3.z, e3.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float ml = __uint_as_float(nw[j] << 16), mh = __uint_as_float(nw[j] & 0xFFFF0000u);
                    f[2 * j] = ml > 0.f ? alpha * __uint_as_float(v1[2 * j]) : 0.f;
                    f[2 * j + 1] = mh > 0.f ? alpha * __uint_as_float(v1[2 * j + 1]) : 0.f;
                  }
                  store16(op + 32, f, vo);
                }
              } else {                                               // fast == 3: alpha * acc + residual
                const uint32_t rw[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  f[2 * j] = fmaf(alpha, __uint_as_float(v0[2 * j]), __uint_as_float(rw[j] << 16));
                  f[2 * j + 1] = fmaf(alpha, __uint_as_float(v0[2 * j + 1]), __uint_as_float(rw[j] & 0xFFFF0000u));
                }
                store16(op, f, vo);
                if (two) {
                  const uint32_t sw[8] = {e2.x, e2.y, e2.z, e2.w, e3.x, e3.y, e3.z, e3.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    f[2 * j] = fmaf(alpha, __uint_as_float(v1[2 * j]), __uint_as_float(sw[j] << 16));
                    f[2 * j + 1] = fmaf(alpha, __uint_as_float(v1[2 * j + 1]), __uint_as_float(sw[j] & 0xFFFF0000u));
                  }
                  store16(op + 32, f, vo);
                }
              }
            }
            op += ostride;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[slot]);
            if (++slot == S) { slot = 0; par ^= 1u; }
          }
        }
      }
      if (want_cs && has0) {
        // fused bias gradient: the warp's 32 pixels, then one atomic per channel and warp
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = warp_sum(cs[j]);
          if (lane == 0) atomicAdd(a.colsum + ch0 + j, t);
        }
      }
    } else
    {
     ItemIter it(a);
     int n, strip, y0, rows;
     while (it.next(a, n, strip, y0, rows)) {
      const int x = strip * BM + row;
      const bool valid = x < a.W;
      const int64_t p0 = ((int64_t)n * a.H + y0) * a.W + x;
      const bool pfv = pf && valid;
      if (pfv) {
#pragma unroll
        for (int d = 0; d < 4; ++d)
          if (d < rows) {
            const uint4* mp = reinterpret_cast<const uint4*>(a.mask + (p0 + (int64_t)d * a.W) * a.ldmask + cm);
            mqa[d] = mp[0];
            mqb[d] = mp[1];
          }
      }
      for (int oi4 = 0; oi4 < rows; oi4 += 4) {
#pragma unroll
       for (int D = 0; D < 4; ++D) {             // static D after unrolling: the prefetch slots stay in registers
        const int oi = oi4 + D;
        if (oi >= rows) break;
        const int64_t p = p0 + (int64_t)oi * a.W;
        const uint4 cur0 = mqa[D], cur1 = mqb[D];
        if (pfv && oi + 4 < rows) {
          const uint4* mp = reinterpret_cast<const uint4*>(a.mask + (p + 4 * (int64_t)a.W) * a.ldmask + cm);
          mqa[D] = mp[0];
          mqb[D] = mp[1];
        }
        mbar_wait(&acc_full[slot], par);
        tc_fence_after();
        const uint32_t tcol = lane_addr + (uint32_t)(slot * a.NOUT);
        // column address such that (taddr + absolute channel) is the accumulator column of that channel
        const uint32_t taddr = tcol - (uint32_t)c_lo;
        if (PF) {                                                  // one chunk per thread (c = part), mask prefetched
          if (part < nch_all) {
            EpiChunk<OutT> e0;
            const bool live0 = part < nch && !ROWS_DBG(a, 4);
            if (live0) e0.issue(a, taddr, c_lo + part * 16, valid, p, pfv, cur0, cur1);
            tmem_ld_wait();
            tmem_st16_zero(tcol + (uint32_t)(part * 16));          // re-arm the slot for its next output row
            if (live0) e0.finish(a, c_lo + part * 16, valid, p);
          }
        } else {
          for (int c = part; c < nch_all; c += 4) {
            EpiChunk<OutT> e0, e1;
            const bool two = c + 2 < nch_all;
            const bool live0 = c < nch && !ROWS_DBG(a, 4), live1 = two && c + 2 < nch && !ROWS_DBG(a, 4);
            if (live0) e0.issue(a, taddr, c_lo + c * 16, valid, p);
            if (live1) e1.issue(a, taddr, c_lo + (c + 2) * 16, valid, p);
            tmem_ld_wait();
            tmem_st16_zero(tcol + (uint32_t)(c * 16));            // re-arm the slot for its next output row
            if (two) tmem_st16_zero(tcol + (uint32_t)((c + 2) * 16));
            if (live0) e0.finish(a, c_lo + c * 16, valid, p);
            if (live1) e1.finish(a, c_lo + (c + 2) * 16, valid, p);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
        if (++slot == S) { slot = 0; par ^= 1u; }
       }
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

struct RowPlan {
  int NOUT, nsplit, nchunks, ksteps_last, nchunks2, ksteps2_last, cps, stages, slots, merged, strips, R, segs;
  size_t smem;
};

// Choose the per-CTA channel group, ring depths and the rows-per-item that minimises the longest CTA.
bool plan_rows(const nervecl_conv_params& a, int sms, RowPlan& p) {
  p.nchunks = (a.Cin + KC - 1) / KC;
  p.ksteps_last = (a.Cin - (p.nchunks - 1) * KC + 15) / 16;
  p.nchunks2 = a.x2 ? (a.Cin2 + KC - 1) / KC : 0;
  p.ksteps2_last = a.x2 ? (a.Cin2 - (p.nchunks2 - 1) * KC + 15) / 16 : 0;
  const int nct = p.nchunks + p.nchunks2;
  const int ntiles = (a.K == 1 ? 1 : 3) * p.nchunks + ((a.x2 && a.x2_center) ? 1 : 3) * p.nchunks2;
  const int cout16 = (a.Cout + 15) / 16 * 16;
  p.nsplit = 0;
  for (int ns = 1; ns <= 16 && !p.nsplit; ++ns) {
    const int nout = ((cout16 + ns - 1) / ns + 15) / 16 * 16;
    if (nout > 128) continue;
    const size_t w_bytes = (size_t)3 * ntiles * nout * ROWB;
    const size_t fixed = 1024 + w_bytes + (2 * kMaxStages + 2 * kMaxSlots + 2) * sizeof(uint64_t) + 64;
    if (fixed >= kSmemBudget) continue;
    const int chunks_fit = (int)((kSmemBudget - fixed) / CHUNK_BYTES);
    // a ring stage holds `cps` chunks (a divisor of the chunks per row): whole rows when two of them fit
    // (one barrier round trip and one commit per row), else the largest group that still leaves >= 3 stages
    static const int cps_cap = nv::tune_env("NERVECL_ROWS_CPS") ? atoi(nv::tune_env("NERVECL_ROWS_CPS")) : 1 << 20;   // (tuning knob)
    for (int cps = nct; cps >= 1; --cps) {
      if (nct % cps || cps > cps_cap) continue;
      const int st = chunks_fit / cps;
      if (st >= (cps == nct ? 2 : 3)) {
        p.nsplit = ns;
        p.NOUT = nout;
        p.cps = cps;
        p.stages = (int)imin(kMaxStages, st);
        p.smem = fixed + (size_t)p.stages * cps * CHUNK_BYTES;
        break;
      }
    }
  }
  if (!p.nsplit) return false;
  p.slots = (int)imin(kMaxSlots, 512 / p.NOUT);
  p.merged = 3 * p.NOUT <= 256;
  p.strips = (a.W + BM - 1) / BM;
  const int ctas = (int)imax(1, sms / p.nsplit);
  int64_t best = -1;
  p.R = a.H;
  for (int R = (int)imin(a.H, 6); R <= a.H && R <= 96; ++R) {
    const int segs = (a.H + R - 1) / R;
    const int64_t items = (int64_t)a.N * p.strips * segs;
    const int64_t cost = ((items + ctas - 1) / ctas) * (R + 3);
    if (best < 0 || cost <= best) { best = cost; p.R = R; }
  }
  p.segs = (a.H + p.R - 1) / p.R;
  if (p.smem < 120 * 1024) p.smem = 120 * 1024;     // one CTA per SM: each allocates all 512 TMEM columns
  return true;
}

}  // namespace

namespace nv {

bool conv_rows_supported(const nervecl_conv_params& a) {
  nervecl_conv_params b = a;
  b.x2 = nullptr;
  if (!conv_tc_fwd_supported(b)) return false;       // dtype / alignment / epilogue constraints are the same
  if (a.K != 3 && a.K != 1) return false;
  if (a.K == 1 && a.x2) return false;
  if (a.Cin < 16 || a.Cin % 16 || a.Cin > 512) return false;
  if (a.W < 64 || a.H < 3) return false;
  if ((int64_t)a.N * ((a.W + BM - 1) / BM) * a.H >= (int64_t)1 << 30) return false;
  if (a.x2) {
    if (a.Cin2 < 16 || a.Cin2 % 16 || a.Cin2 > 256 || a.ldx2 % 8 || !aligned(a.x2, 16)) return false;
    if (a.w_ld < (a.Cin + KC - 1) / KC * KC + a.Cin2) return false;   // x2 weights start at the next 64-column boundary
  }
  RowPlan p;
  return plan_rows(a, sm_count(), p);
}

int conv_rows_fwd(const nervecl_conv_params& a, cudaStream_t s) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return NERVECL_EUNSUPPORTED;
  const int sms = sm_count();
  RowPlan p;
  if (!plan_rows(a, sms, p)) return NERVECL_EUNSUPPORTED;

  CUtensorMap tx, tx2, tw;
  auto encode_act = [&](CUtensorMap* m, const void* base, int C, int64_t ld) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)a.W * ld * 2, (cuuint64_t)a.H * a.W * ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)PXB, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    // a channel count that is not a multiple of 64 ends inside a 128-byte line whose other half belongs to a
    // neighbouring slice of the same buffer: 128-byte L2 promotion would fetch it from DRAM (ncu: 96 -> 32 read
    // 128 channels' worth), 64-byte promotion does not
    static const bool promo64 = !nv::tune_env("NERVECL_ROWS_PROMO128");
    const CUtensorMapL2promotion promo = (C % KC && promo64) ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!encode_act(&tx, a.x, a.Cin, a.ldx)) return NERVECL_EUNSUPPORTED;
  if (a.x2) {
    if (!encode_act(&tx2, a.x2, a.Cin2, a.ldx2)) return NERVECL_EUNSUPPORTED;
  } else {
    tx2 = tx;
  }
  {
    // the x2 weight columns start at column nchunks*64 of the packed rows; TMA boxes address them as chunk
    // (nchunks + c2), so one map over the whole row length serves both inputs
    cuuint64_t dims[3] = {(cuuint64_t)a.w_ld, (cuuint64_t)a.w_rows, (cuuint64_t)(a.K * a.K)};
    cuuint64_t strides[2] = {(cuuint64_t)a.w_ld * 2, (cuuint64_t)a.w_rows * a.w_ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)p.NOUT, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a.w), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NERVECL_EUNSUPPORTED;
  }

  RowArgs t;
  t.Cout = a.Cout; t.relu = a.relu; t.accumulate = a.accumulate; t.res_channels = a.res ? a.res_channels : 0;
  t.mask_c0 = a.mask_c0; t.alpha = a.alpha; t.bias = a.bias;
  t.res = (const bf16*)a.res; t.ldres = a.ldres;
  t.mask = (const bf16*)a.mask; t.ldmask = a.ldmask;
  t.msub = (const bf16*)a.mask_sub; t.ldmsub = a.ldmask_sub;
  t.out = a.out; t.ldo = a.ldo;
  t.colsum = a.colsum;
  // 256-bit epilogue accesses: 16-channel chunks start on 32-byte boundaries of 32-byte aligned pixels
  t.v256_out = a.out_dtype == NERVECL_BF16 && a.ldo % 16 == 0 && aligned(a.out, 32);
  t.v256_in = 0;
  t.v256_gen = epi_v256(a);
  t.N = a.N; t.H = a.H; t.W = a.W;
  t.NOUT = p.NOUT; t.nchunks = p.nchunks; t.ksteps_last = p.ksteps_last;
  t.nchunks2 = p.nchunks2; t.ksteps2_last = p.ksteps2_last; t.x2_center = a.x2 ? a.x2_center : 0;
  t.x_center = a.K == 1;
  t.strips = p.strips; t.R = p.R; t.segs = p.segs; t.cps = p.cps; t.stages = p.stages; t.slots = p.slots; t.merged = p.merged;
  { const char* d = nv::tune_env("NERVECL_ROWS_DBG"); t.dbg = d ? atoi(d) : 0; }
  const bool whole = a.Cout % 16 == 0 && a.Cout % p.NOUT == 0 && a.out_dtype == NERVECL_BF16 && !a.mask_sub && a.relu != 2 &&
                     p.NOUT <= 128 && !ROWS_DBG(t, 64);
  t.fast = 0;
  if (whole && !a.accumulate) {
    if ((!a.res || a.res_channels >= a.Cout) && !a.mask && a.alpha == 1.0f && (a.bias || a.relu || !a.res)) t.fast = 1;
    if (!a.res && a.mask && !a.bias && !a.relu && a.mask_c0 == 0) t.fast = 2;
    if (a.res && a.res_channels >= a.Cout && !a.mask && !a.bias && !a.relu) t.fast = 3;
  } else if (whole && !a.res && !a.mask && !a.bias && !a.relu) {
    // out += alpha * acc is the residual form with the output itself as the residual (each element is read and
    // written by the same thread)
    t.fast = 3;
    t.res = (const bf16*)a.out; t.ldres = a.ldo; t.res_channels = a.Cout;
  }

  // one CTA per SM (per channel group); each takes an equal share of the N * strips * H output rows
  if (t.fast == 2) t.v256_in = a.ldmask % 16 == 0 && aligned(a.mask, 32);
  if (t.fast == 3 || (t.fast == 1 && t.res)) t.v256_in = t.ldres % 16 == 0 && aligned(t.res, 32);
  if (nv::tune_env("NERVECL_NO_V256")) t.v256_out = t.v256_in = 0;
  if (a.colsum && !(t.fast == 2 && p.NOUT <= 32)) return NERVECL_EUNSUPPORTED;
  const int64_t total_rows = (int64_t)a.N * p.strips * a.H;
  dim3 grid((unsigned)imin(cdiv(total_rows, 4), imax(1, sms / p.nsplit)), (unsigned)p.nsplit);
  cudaError_t e = cudaSuccess;
  const bool pf = a.mask && !a.mask_sub && p.NOUT <= 32;
#define NV_LAUNCH_ROWS(OT, PFV)                                                                                       \
  do {                                                                                                                \
    e = cudaFuncSetAttribute(conv_rows_kernel<OT, PFV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);     \
    if (e != cudaSuccess) return (int)e;                                                                              \
    conv_rows_kernel<OT, PFV><<<grid, kThreads, p.smem, s>>>(tx, tx2, tw, t);                                          \
  } while (0)
  if (a.out_dtype == NERVECL_F32) {
    if (pf) NV_LAUNCH_ROWS(float, true); else NV_LAUNCH_ROWS(float, false);
  } else {
    if (pf) NV_LAUNCH_ROWS(bf16, true); else NV_LAUNCH_ROWS(bf16, false);
  }
#undef NV_LAUNCH_ROWS
  int rc = launch_status();
  if (rc == (int)cudaErrorLaunchOutOfResources) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, conv_rows_kernel<bf16, false>) == cudaSuccess)
      fprintf(stderr, "nervecl conv_rows: launch out of resources: regs %d maxThreads %d static smem %zu maxDyn %d; requested "
                      "dyn smem %zu, block %d, grid %u x %u\n",
              fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, p.smem, kThreads, grid.x,
              grid.y);
  }
  return rc;
}

}  // namespace nv
