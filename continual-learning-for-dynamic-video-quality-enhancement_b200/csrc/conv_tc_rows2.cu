// Row-streaming 3x3 convolution on a CTA PAIR (tcgen05.mma.cta_group::2, M = 256): the second generation of
// conv_tc_rows.cu for the launches that dominate the training step (dense-block layers, their fused slice gradients,
// flow / attention convs) -- bf16 in, bf16 out, lean epilogues.
//
// What the 1-CTA kernel was bound by (profiles/r01f_summary.md): an N = 96 MMA is fed from shared
// memory at 7 KB / 128 B per clock = 55 cycles against 48 cycles of tensor work, the issuing thread spends another
// ~1000 cycles per 128-pixel row on barrier round trips and commits, and layers with >= 64 output channels over >= 192
// input channels could not keep their weights resident (the launch was split into channel groups that each re-read
// the input).  Here two CTAs of one TPC share every MMA:
//   * each CTA streams ITS OWN pixels (any 128-pixel column strip of any image: the two CTAs of a pair only have to
//     walk row ranges of equal length) and holds HALF of the weight rows, so an MMA reads 4 + 1.5 KB per SM and the
//     per-row issue overhead is paid once per 256 pixels;
//   * every input row is ONE merged MMA group with N = 3*NOUT over three adjacent accumulator slots (output rows r-1,
//     r, r+1): the accumulator ring has two MIRROR slots behind its end, so a group that would wrap writes past the
//     end instead and the epilogue adds slot s and its mirror for the two output rows per lap that were split; item
//     borders are absorbed by two GAP slots between consecutive items (the first / last input rows of an item add
//     their out-of-item taps there; the epilogue drains them without storing).  No per-ky fallback MMAs, so the B
//     operand split between the CTAs is static: rank 0 holds tile rows [0, 1.5 NOUT), rank 1 the rest;
//   * the TMA producers of both CTAs signal the leader's "stage full" barrier (cta_group::2 loads), the leader's
//     commits are multicast to both CTAs' "stage empty" / "accumulator full" barriers, the epilogue warps of both
//     CTAs arrive on the leader's "accumulator empty" barrier.
#include "conv_internal.cuh"
#include "conv_tc_epilogue.cuh"
#include "tc_common.cuh"
#include <cstdio>

using namespace nv;
using namespace nv::tc;

namespace {

constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);     // warp 0 = TMA, warp 1 = MMA issuer (leader CTA), warps 2..9 = epilogue
constexpr int KC = 64;
constexpr uint32_t ROWB = KC * 2;
constexpr int PXB = BM + 2;
constexpr uint32_t CHUNK_BYTES = (PXB * ROWB + 1023u) & ~1023u;
constexpr int kMaxStages = 12, kMaxSlots = 30;
constexpr size_t kSmemBudget = 226 * 1024;

struct Pair2Args {
  // epilogue: kind 1: relu?(acc + bias?) (+ res);  2: alpha * acc where mask > 0 (+ column sums);  3: alpha * acc + res
  int kind, relu;
  float alpha;
  const float* bias;
  const bf16* eop;      // mask (kind 2) or residual (kind 1 / 3), nullptr if none
  int64_t ldeop;
  bf16* out;
  int64_t ldo;
  float* colsum;
  uint16_t* sbits;      // packed ReLU signs, [Cout / 16][pixel] (nervecl_conv_params.sign_bits)
  int smode;            // 0 none, 1 write (kind 1), 2 read as the mask (kind 2)
  int64_t npix;         // N * H * W (plane stride of the sign words)
  int v256_out, v256_in;
  // geometry
  int N, H, W, Cout;
  int NOUT;             // output channels per CTA pair (multiple of 16, <= 80)
  int nchunks, ksteps_last, nchunks2, ksteps2_last, x2_center;
  int strips, cps, stages, slots;
  int R;                // output rows per CTA (nominal length of a CTA's row range)
};

// Lock-step work split.  The linearised output rows ((n * strips + strip) * H + y) are cut into ranges of R rows; pair p
// owns ranges 2p (rank 0) and 2p + 1 (rank 1).  Both CTAs walk their range in PIECES that end wherever EITHER range
// crosses a column (image, strip) boundary, so piece k has the same row count in both CTAs (one MMA serves both) and
// lies inside one column in each.  A rank whose range lies beyond the end of the work runs the leader's piece as a
// dead item: loads and MMAs happen, nothing is stored.
struct PieceIter {
  int64_t T, base0, base1;
  int H, R, strips, o;
  __device__ __forceinline__ PieceIter(const Pair2Args& a, int pair) {
    T = (int64_t)a.N * a.strips * a.H;
    H = a.H; R = a.R; strips = a.strips; o = 0;
    base0 = (int64_t)(2 * pair) * a.R;
    base1 = base0 + a.R;
  }
  __device__ __forceinline__ bool next(int rank, int& n, int& strip, int& y0, int& rows, bool& live) {
    const int64_t p0 = base0 + o, p1 = base1 + o;
    if (o >= R || p0 >= T) return false;
    const int b0 = H - (int)(p0 % H);
    const int b1 = p1 < T ? H - (int)(p1 % H) : R;
    rows = min(min(R - o, b0), b1);
    int64_t p = rank ? p1 : p0;
    live = p < T;
    if (!live) p = p0;
    const int64_t unit = p / H;
    y0 = (int)(p - unit * H);
    n = (int)(unit / strips);
    strip = (int)(unit - (int64_t)n * strips);
    o += rows;
    return true;
  }
};

__device__ __forceinline__ void unpack8(const uint4& a, const uint4& b, float (&f)[16]) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}

// NC = 16-channel chunks per epilogue thread (1: NOUT <= 32, 2: NOUT <= 64, 3: NOUT <= 80): a compile-time bound keeps
// the single-chunk epilogue of the dense-block launches small enough for a three-row operand prefetch queue.
template <int NC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_rows2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_x2,
                  const __grid_constant__ CUtensorMap tmap_w, const Pair2Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int grp = blockIdx.y;
  const uint32_t half_rows = 3u * a.NOUT / 2u;                  // weight rows of a tile held by this CTA
  const uint32_t tile_bytes = half_rows * ROWB;
  const int nct = a.nchunks + a.nchunks2;
  const int ntiles = 3 * a.nchunks + (a.x2_center ? 1 : 3) * a.nchunks2;
  const uint32_t w_bytes = (uint32_t)ntiles * tile_bytes;
  uint8_t* w_smem = smem;
  uint8_t* ring = smem + w_bytes;
  const uint32_t stage_bytes = (uint32_t)a.cps * CHUNK_BYTES;
  const int ngrp = nct / a.cps;
  const int S = a.slots;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)a.stages * stage_bytes);
  uint64_t* ch_full = bars;                         // [stages]  both TMA producers -> leader's MMA issuer
  uint64_t* ch_empty = ch_full + a.stages;          // [stages]  leader's commits (multicast) -> each CTA's producer
  uint64_t* acc_full = ch_empty + a.stages;         // [slots]   leader's commits (multicast) -> each CTA's epilogue
  uint64_t* acc_empty = acc_full + S;               // [slots]   epilogue warps of BOTH CTAs -> leader's MMA issuer
  uint64_t* w_bar = acc_empty + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    if (a.nchunks2) prefetch_tmap(&tmap_x2);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&ch_full[s], 1);                    // the leader producer's arrive.expect_tx (bytes of both CTAs)
      mbar_init(&ch_empty[s], 1);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 2 * kEpiWarps);      // epilogue warps of both CTAs
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();                               // barriers of both CTAs initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- weights: this CTA's half of every tile, then a rendezvous so that the leader may read both halves ----
  if (warp == 0 && lane == 0) {
    // tile rows are [ky=2 | ky=1 | ky=0] blocks of NOUT rows; rank 0 holds rows [0, 1.5 NOUT), rank 1 rows
    // [1.5 NOUT, 3 NOUT) -- three half-blocks of NOUT/2 rows each
    const uint32_t hb_bytes = (uint32_t)(a.NOUT / 2) * ROWB;
    mbar_expect_tx(w_bar, w_bytes);
    for (int c = 0; c < nct; ++c)
      for (int kx = 0; kx < 3; ++kx) {
        if (c >= a.nchunks && a.x2_center && kx != 1) continue;
        // tile index: x chunks kx-major, then x2 chunks (one tile per chunk when centre-only)
        const int t = c < a.nchunks ? kx * a.nchunks + c
                                    : 3 * a.nchunks + (a.x2_center ? (c - a.nchunks) : kx * a.nchunks2 + (c - a.nchunks));
        for (int h = 0; h < 3; ++h) {
          const int hbi = (int)rank * 3 + h;                 // half-block index 0..5 within the tile
          const int b = hbi >> 1, hh = hbi & 1;              // block (ky = 2 - b), which half of its rows
          tma_load_3d(w_smem + (size_t)t * tile_bytes + (size_t)h * hb_bytes, &tmap_w, w_bar, c * KC,
                      grp * a.NOUT + hh * (a.NOUT / 2), (2 - b) * 3 + kx);
        }
      }
  }
  if (warp == 1) {
    mbar_wait(w_bar, 0);
    tc_fence_before();
  }
  cluster_sync_all();

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      PieceIter it(a, pair);
      int n, strip, y0, rows;
      bool live;
      const uint32_t bytes = (uint32_t)a.cps * (PXB * ROWB);
      while (it.next((int)rank, n, strip, y0, rows, live)) {
        const int x0 = strip * BM;
        for (int ri = 0; ri < rows + 2; ++ri) {
          for (int gi = 0; gi < ngrp; ++gi) {
            mbar_wait(&ch_empty[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&ch_full[stage], 2u * bytes);
            for (int cc = 0; cc < a.cps; ++cc) {
              const int c = gi * a.cps + cc;
              uint8_t* dst = ring + (size_t)stage * stage_bytes + (size_t)cc * CHUNK_BYTES;
              if (c < a.nchunks) tma_load_4d_pair(dst, &tmap_x, &ch_full[stage], c * KC, x0 - 1, y0 - 1 + ri, n);
              else tma_load_4d_pair(dst, &tmap_x2, &ch_full[stage], (c - a.nchunks) * KC, x0 - 1, y0 - 1 + ri, n);
            }
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================= MMA issuer (leader CTA) =================
    tc_fence_after();
    const uint32_t w_lo = (smem_u32(w_smem) & 0x3FFFFu) >> 4;
    const uint32_t ring_lo = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t tile_lo = tile_bytes >> 4;
    const uint64_t desc_hi = make_kmajor_desc(0, ROWB) & 0xFFFFFFFF00000000ull;
    const uint32_t id3 = make_idesc_bf16_m256(3u * a.NOUT);
    int stage = 0;
    uint32_t phase = 0;
    int p = 0;                                       // accumulator slot of virtual output row v = u (u = input rows so far)
    int s2 = 2 % S;                                  // slot of v = u + 2 (first written by input row u) ...
    uint32_t k2 = 2 / S;                             // ... and how many times it has been used before
    // As in the 1-CTA kernel, the commits leave the row-to-row critical path: "virtual output row final" goes out behind
    // the first MMA group of the NEXT input row, "stage consumed" behind the second group of the next stage (deep rings
    // only), and the barriers the next iteration needs are probed right after this one's MMAs were issued.
    int pend = -1, pend_stage = -1;
    bool ready = false, acc_ready = false;
    const bool defer_stage = a.stages >= 4;
    PieceIter it(a, pair);
    int n_, strip_, y0_, rows;
    bool live_;
    // (-DR2_PROF: clock64 breakdown of the issuer's and the epilogue's row loop, printed by pair 5; profiles/r02h_experiments.md)
#ifdef R2_PROF
    long long t_acc = 0, t_full = 0, t_mma = 0, t_rest = 0, tp0 = clock64(), tp1;
    int nrows_p = 0;
#define R2T(x) tp1 = clock64(); x += tp1 - tp0; tp0 = tp1;
#else
#define R2T(x)
#endif
    while (it.next(0, n_, strip_, y0_, rows, live_)) {
      for (int ri = 0; ri < rows + 2; ++ri) {
        R2T(t_rest)
        if (!acc_ready) mbar_wait(&acc_empty[s2], k2 & 1u);  // both CTAs drained + re-zeroed that slot
        acc_ready = false;
        R2T(t_acc)
#ifdef R2_PROF
        ++nrows_p;
#endif
        const uint32_t d = tmem_base + (uint32_t)(p * a.NOUT);
        for (int gi = 0; gi < ngrp; ++gi) {
          if (!ready) mbar_wait(&ch_full[stage], phase);
          tc_fence_after();
          R2T(t_full)
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == a.stages) { nstage = 0; nphase ^= 1; }
          const int commit_pend = gi == 0 ? pend : -1;
          if (elect_one()) {
            bool first = commit_pend >= 0;
            int second_left = pend_stage >= 0 ? 2 : 0;
            for (int cc = 0; cc < a.cps; ++cc) {
              const int c = gi * a.cps + cc;
              const uint32_t a_lo = ring_lo + (uint32_t)stage * (stage_bytes >> 4) + (uint32_t)cc * (CHUNK_BYTES >> 4);
              const bool second = c >= a.nchunks;
              const int ks = (c == a.nchunks - 1) ? a.ksteps_last : (c == nct - 1 && second) ? a.ksteps2_last : KC / 16;
              const bool ctr = second && a.x2_center;
              const uint32_t tile0 = (uint32_t)(second ? 3 * a.nchunks + (c - a.nchunks) : c);
              const uint32_t tstride = (uint32_t)(ctr ? 0 : second ? a.nchunks2 : a.nchunks);
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                if (ctr && kx != 1) continue;
                const uint32_t al = a_lo + (uint32_t)kx * (ROWB >> 4);
                const uint32_t bl = (1u << 16) | (w_lo + (tile0 + (uint32_t)kx * tstride) * tile_lo);
                if (ks == KC / 16) {
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k)
                    umma_bf16_acc_pair(d, desc_hi | (uint64_t)(al + 2u * k), desc_hi | (uint64_t)(bl + 2u * k), id3);
                } else {
#pragma unroll
                  for (int k = 0; k < KC / 16 - 1; ++k)
                    if (k < ks) umma_bf16_acc_pair(d, desc_hi | (uint64_t)(al + 2u * k), desc_hi | (uint64_t)(bl + 2u * k), id3);
                }
                if (first) { umma_commit_pair(&acc_full[commit_pend]); first = false; }
                if (second_left && --second_left == 0) umma_commit_pair(&ch_empty[pend_stage]);
              }
            }
            if (first) umma_commit_pair(&acc_full[commit_pend]);
            if (second_left) umma_commit_pair(&ch_empty[pend_stage]);
            if (!defer_stage) umma_commit_pair(&ch_empty[stage]);      // stage consumed when these MMAs retire
          }
          R2T(t_mma)
          pend_stage = defer_stage ? stage : -1;
          if (gi == 0) pend = -1;
          ready = mbar_try_wait(&ch_full[nstage], nphase);
          if (gi == ngrp - 1) {
            const int ns2 = s2 + 1 == S ? 0 : s2 + 1;
            const uint32_t nk2 = s2 + 1 == S ? k2 + 1 : k2;
            acc_ready = mbar_try_wait(&acc_empty[ns2], nk2 & 1u);
          }
          stage = nstage;
          phase = nphase;
        }
        pend = p;                                    // virtual output row v = u is final once these MMAs retire
        if (++p == S) p = 0;
        if (++s2 == S) { s2 = 0; ++k2; }
      }
    }
    if (elect_one()) {
      if (pend_stage >= 0) umma_commit_pair(&ch_empty[pend_stage]);
      if (pend >= 0) umma_commit_pair(&acc_full[pend]);
    }
    __syncwarp();
#ifdef R2_PROF
    if (pair == 5 && lane == 0 && blockIdx.y == 0)
      printf("MMA issuer: rows %d  per row: wait acc_empty %lld  wait ch_full %lld  issue %lld  rest %lld\n", nrows_p,
             t_acc / nrows_p, t_full / nrows_p, t_mma / nrows_p, t_rest / nrows_p);
#endif
  } else if (warp >= 2) {
    // ================= epilogue (warps 2..9 of both CTAs) =================
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int c_lo = grp * a.NOUT;
    const int nch_all = a.NOUT >> 4;                  // <= 5
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t mirror = (uint32_t)(S * a.NOUT);   // column offset of a slot's mirror
    for (int col = part * 16; col < (S + 2) * a.NOUT; col += 32) tmem_st16_zero(lane_addr + (uint32_t)col);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    const uint32_t empty0 = mapa_u32(acc_empty, 0);   // the leader's acc_empty[0] in the cluster address space
    if (lane == 0)
      for (int sl = 0; sl < S; ++sl) {
        if (rank == 0) mbar_arrive(&acc_empty[sl]); else mbar_arrive_cluster(empty0 + 8u * sl);
      }

    bool has[NC];
    int ch[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      has[i] = part + 2 * i < nch_all;
      ch[i] = c_lo + (part + 2 * i) * 16;
    }
    const int kind = a.kind;
    const bool relu = a.relu != 0, has_e = a.eop != nullptr;
    const float alpha = a.alpha;
    const bool vo = a.v256_out != 0, vi = a.v256_in != 0;
    const bool want_cs = a.colsum != nullptr;        // (host: only with kind 2 and NOUT <= 32, i.e. one chunk per thread)
    const int smode = NC <= 2 ? a.smode : 0;       // (host: sign bits only with <= 64 output channels per CTA pair)
    const int sgrp = c_lo / 16 + part;               // 16-channel group of the thread's first chunk (chunk i: sgrp + 2 i)
    // one register array serves both: the bias of the thread's first chunk (kind 1) or the running column sums (kind 2)
    float cs[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) cs[j] = (kind == 1 && a.bias && has[0]) ? __ldg(a.bias + ch[0] + j) : 0.f;

#ifdef R2_PROF
    long long e_wait = 0, e_ld = 0, e_comp = 0, e_rel = 0, ep0 = clock64(), ep1;
    int erows = 0;
#define R2E(x) ep1 = clock64(); x += ep1 - ep0; ep0 = ep1;
#else
#define R2E(x)
#endif
    int slot = 0;
    uint32_t par = 0;
    auto release = [&]() {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&acc_empty[slot]); else mbar_arrive_cluster(empty0 + 8u * slot);
      }
      if (++slot == S) { slot = 0; par ^= 1u; }
    };

    PieceIter it(a, pair);
    int n, strip, y0, rows;
    bool live;
    while (it.next((int)rank, n, strip, y0, rows, live)) {
      const int x = strip * BM + row;
      const bool valid = live && x < a.W;
      const int64_t p0 = ((int64_t)n * a.H + y0) * a.W + x;
      bf16* op = a.out + p0 * a.ldo;
      const int64_t ostride = (int64_t)a.W * a.ldo;
      const bf16* ep = has_e ? a.eop + p0 * a.ldeop : nullptr;
      const int64_t estride = (int64_t)a.W * a.ldeop;
      // first operand row in flight before anything is waited on
      uint4 pre[NC][2];
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        pre[i][0] = pre[i][1] = make_uint4(0, 0, 0, 0);
        if (has_e && valid && has[i]) load32B(ep + ch[i], pre[i][0], pre[i][1], vi);
      }
      // sign words of rows oi .. oi + 3 (mode 2: 2 bytes per row instead of the 32-byte mask operand)
      const uint16_t* sp = a.sbits + (int64_t)sgrp * a.npix + p0;       // (a warp's 32 pixels: 64 contiguous bytes)
      const int64_t sstride = a.W, splane = 2 * a.npix;                 // next row / next chunk of this thread
      constexpr int SBQ = 8;                          // sign words in flight per chunk (2 bytes each: a deep queue is free)
      uint32_t sbq[NC][SBQ];
#pragma unroll
      for (int i = 0; i < NC; ++i) {
#pragma unroll
        for (int d = 0; d < SBQ; ++d) {
          sbq[i][d] = 0;
          if (smode == 2 && valid && has[i] && d < rows) sbq[i][d] = __ldg(sp + i * splane + (int64_t)d * sstride);
        }
      }
      uint4 q1[2], q2[2];                           // NC == 1: rows oi + 1 and oi + 2 of the operand are in flight as well
      q1[0] = q1[1] = q2[0] = q2[1] = make_uint4(0, 0, 0, 0);
      if (NC == 1 && has_e && valid && has[0]) {
        if (rows > 1) load32B(ep + estride + ch[0], q1[0], q1[1], vi);
        if (rows > 2) load32B(ep + 2 * estride + ch[0], q2[0], q2[1], vi);
      }
      // the two gap slots in front of the item: drained and re-zeroed, nothing stored
      for (int gslot = 0; gslot < 2; ++gslot) {
        mbar_wait(&acc_full[slot], par);
        tc_fence_after();
        const uint32_t tcol = lane_addr + (uint32_t)(slot * a.NOUT);
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (has[i]) {
            tmem_st16_zero(tcol + (uint32_t)((part + 2 * i) * 16));
            if (slot < 2) tmem_st16_zero(tcol + mirror + (uint32_t)((part + 2 * i) * 16));
          }
        release();
      }
      for (int oi = 0; oi < rows; ++oi) {
        uint4 cur[NC][2];
#pragma unroll
        for (int i = 0; i < NC; ++i) { cur[i][0] = pre[i][0]; cur[i][1] = pre[i][1]; }
        uint32_t sbc[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          sbc[i] = sbq[i][0];
          if (smode == 2) {
#pragma unroll
            for (int d = 0; d + 1 < SBQ; ++d) sbq[i][d] = sbq[i][d + 1];
            if (valid && has[i] && oi + SBQ < rows) sbq[i][SBQ - 1] = __ldg(sp + i * splane + (int64_t)(oi + SBQ) * sstride);
          }
        }
        // the operand rows behind the current one: NC == 1 keeps rows oi + 1 .. oi + 3 in flight, wider threads row oi + 1
        if (NC == 1) {
          pre[0][0] = q1[0]; pre[0][1] = q1[1];
          q1[0] = q2[0]; q1[1] = q2[1];
          if (has_e && valid && has[0] && oi + 3 < rows) load32B(ep + (int64_t)(oi + 3) * estride + ch[0], q2[0], q2[1], vi);
        } else if (has_e && valid && oi + 1 < rows) {
          const bf16* en = ep + (int64_t)(oi + 1) * estride;
#pragma unroll
          for (int i = 0; i < NC; ++i)
            if (has[i]) load32B(en + ch[i], pre[i][0], pre[i][1], vi);
        }
        R2E(e_comp)
        mbar_wait(&acc_full[slot], par);
        tc_fence_after();
        R2E(e_wait)
#ifdef R2_PROF
        ++erows;
#endif
        const uint32_t tcol = lane_addr + (uint32_t)(slot * a.NOUT);
        uint32_t v[NC][16];
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (has[i]) tmem_ld16(tcol + (uint32_t)((part + 2 * i) * 16), v[i]);
        tmem_ld_wait();
        if (slot < 2) {                               // this output row's first taps were written to the mirror slot
#pragma unroll 1
          for (int i = 0; i < NC; ++i)                 // (two rows per lap of the ring: one chunk at a time, few registers)
            if (has[i]) {
              uint32_t m[16];
              tmem_ld16(tcol + mirror + (uint32_t)((part + 2 * i) * 16), m);
              tmem_ld_wait();
              tmem_st16_zero(tcol + mirror + (uint32_t)((part + 2 * i) * 16));
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float t = __uint_as_float(m[j]);
                if (i == 0) v[0][j] = __float_as_uint(__uint_as_float(v[0][j]) + t);
                else if (NC > 1 && i == 1) v[NC > 1 ? 1 : 0][j] = __float_as_uint(__uint_as_float(v[NC > 1 ? 1 : 0][j]) + t);
                else if (NC > 2) v[NC > 2 ? 2 : 0][j] = __float_as_uint(__uint_as_float(v[NC > 2 ? 2 : 0][j]) + t);
              }
            }
        }
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (has[i]) tmem_st16_zero(tcol + (uint32_t)((part + 2 * i) * 16));
        R2E(e_ld)
        if (valid) {
#pragma unroll
          for (int i = 0; i < NC; ++i) {
            if (!has[i]) continue;
            float f[16], e[16];
            if (has_e) unpack8(cur[i][0], cur[i][1], e);
            if (kind == 1) {
              if (i == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[i][j]) + cs[j];
              } else {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 b = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + ch[i]) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
                  f[4 * j4] = __uint_as_float(v[i][4 * j4]) + b.x;
                  f[4 * j4 + 1] = __uint_as_float(v[i][4 * j4 + 1]) + b.y;
                  f[4 * j4 + 2] = __uint_as_float(v[i][4 * j4 + 2]) + b.z;
                  f[4 * j4 + 3] = __uint_as_float(v[i][4 * j4 + 3]) + b.w;
                }
              }
              if (relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
              }
              if (has_e) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] += e[j];
              }
              if (smode == 1) {                       // (relu: every value is >= 0)
                uint32_t r8[8];
                pack16(f, r8);
                a.sbits[(int64_t)(sgrp + 2 * i) * a.npix + p0 + (int64_t)oi * a.W] = (uint16_t)sign_word16(r8);
                store16(op + ch[i], r8, vo);
                continue;
              }
            } else if (kind == 2) {
              if (smode == 2) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = ((sbc[i] >> sign_bit_pos(j)) & 1u) ? alpha * __uint_as_float(v[i][j]) : 0.f;
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = e[j] > 0.f ? alpha * __uint_as_float(v[i][j]) : 0.f;
              }
              if (want_cs && i == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) cs[j] += f[j];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaf(alpha, __uint_as_float(v[i][j]), e[j]);
            }
            store16(op + ch[i], f, vo);
          }
        }
        op += ostride;
        R2E(e_comp)
        release();
        R2E(e_rel)
      }
    }
#ifdef R2_PROF
    if (pair == 5 && lane == 0 && blockIdx.y == 0 && rank == 0 && (warp == 2 || warp == 7))
      printf("epilogue warp %d: rows %d  per row: wait acc_full %lld  tmem ld/zero %lld  compute+store %lld  release %lld\n", warp,
             erows, e_wait / erows, e_ld / erows, e_comp / erows, e_rel / erows);
#endif
    if (want_cs && has[0]) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float t = warp_sum(cs[j]);
        if (lane == 0) atomicAdd(a.colsum + ch[0] + j, t);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                                // nobody exits while the peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

struct Pair2Plan {
  int NOUT, nsplit, nchunks, ksteps_last, nchunks2, ksteps2_last, cps, stages, slots, strips, pairs, R;
  size_t smem;
};

bool plan_rows2(const nervecl_conv_params& a, int sms, Pair2Plan& p) {
  p.nchunks = (a.Cin + KC - 1) / KC;
  p.ksteps_last = (a.Cin - (p.nchunks - 1) * KC + 15) / 16;
  p.nchunks2 = a.x2 ? (a.Cin2 + KC - 1) / KC : 0;
  p.ksteps2_last = a.x2 ? (a.Cin2 - (p.nchunks2 - 1) * KC + 15) / 16 : 0;
  const int nct = p.nchunks + p.nchunks2;
  const int ntiles = 3 * p.nchunks + ((a.x2 && a.x2_center) ? 1 : 3) * p.nchunks2;
  p.nsplit = 0;
  for (int ns = 1; ns <= 16 && !p.nsplit; ++ns) {
    if (a.Cout % ns) continue;
    const int nout = a.Cout / ns;
    if (nout % 16 || nout > 80) continue;
    const size_t w_bytes = (size_t)ntiles * (3 * nout / 2) * ROWB;
    const size_t fixed = 1024 + w_bytes + (2 * kMaxStages + 2 * kMaxSlots + 2) * sizeof(uint64_t) + 64;
    if (fixed >= kSmemBudget) continue;
    const int chunks_fit = (int)((kSmemBudget - fixed) / CHUNK_BYTES);
    for (int cps = nct; cps >= 1; --cps) {
      if (nct % cps) continue;
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
  p.slots = (int)imin(kMaxSlots, 512 / p.NOUT - 2);
  if (p.slots < 4) return false;
  p.strips = (a.W + BM - 1) / BM;
  const int64_t total = (int64_t)a.N * p.strips * a.H;
  p.pairs = (int)imax(1, (sms / 2) / p.nsplit);
  if (total < (int64_t)p.pairs * 2 * 8) return false;          // too little work to feed every pair: 1-CTA kernel
  p.R = (int)cdiv(total, 2 * (int64_t)p.pairs);
  if (p.smem < 120 * 1024) p.smem = 120 * 1024;                // one CTA per SM: each allocates all 512 TMEM columns
  return true;
}

int lean_kind(const nervecl_conv_params& a) {
  // the same three fused epilogues as the 1-CTA kernel's lean paths
  if (a.out_dtype != NERVECL_BF16 || a.mask_sub || a.relu == 2) return 0;
  if (a.sign_mode == 2)                                        // the mask comes from packed sign bits
    return (!a.accumulate && !a.res && !a.mask && !a.bias && !a.relu) ? 2 : 0;
  if (a.sign_mode == 1 && (a.relu != 1 || a.accumulate || a.mask)) return 0;
  if (!a.accumulate) {
    if ((!a.res || a.res_channels >= a.Cout) && !a.mask && a.alpha == 1.0f && (a.bias || a.relu || !a.res)) return 1;
    if (!a.res && a.mask && !a.bias && !a.relu && a.mask_c0 == 0) return 2;
    if (a.res && a.res_channels >= a.Cout && !a.mask && !a.bias && !a.relu) return 3;
    return 0;
  }
  if (!a.res && !a.mask && !a.bias && !a.relu) return 3;       // out += alpha * acc: the output is its own residual
  return 0;
}

}  // namespace

namespace nv {

bool conv_rows2_supported(const nervecl_conv_params& a) {
  nervecl_conv_params b = a;
  b.x2 = nullptr;
  if (!conv_tc_fwd_supported(b)) return false;
  if (a.K != 3 || a.dtype != NERVECL_BF16) return false;
  if (a.Cin < 16 || a.Cin % 16 || a.Cin > 512 || a.Cout % 16) return false;
  if (a.W < 64 || a.H < 3) return false;
  if ((int64_t)a.N * ((a.W + BM - 1) / BM) * a.H >= (int64_t)1 << 30) return false;
  if (a.x2) {
    if (a.Cin2 < 16 || a.Cin2 % 16 || a.Cin2 > 256 || a.ldx2 % 8 || !aligned(a.x2, 16)) return false;
    if (a.w_ld < (a.Cin + KC - 1) / KC * KC + a.Cin2) return false;
  }
  const int kind = lean_kind(a);
  if (!kind) return false;
  Pair2Plan p;
  if (!plan_rows2(a, sm_count(), p)) return false;
  if (a.colsum && !(kind == 2 && p.NOUT <= 32)) return false;
  if (a.sign_mode && (p.NOUT > 64 || !a.sign_bits || (a.Cout & 15) || kind != a.sign_mode)) return false;
  return true;
}

int conv_rows2_fwd(const nervecl_conv_params& a, cudaStream_t s) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return NERVECL_EUNSUPPORTED;
  Pair2Plan p;
  if (!plan_rows2(a, sm_count(), p)) return NERVECL_EUNSUPPORTED;
  const int kind = lean_kind(a);
  if (!kind) return NERVECL_EUNSUPPORTED;

  CUtensorMap tx, tx2, tw;
  auto encode_act = [&](CUtensorMap* m, const void* base, int C, int64_t ld) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)a.W * ld * 2, (cuuint64_t)a.H * a.W * ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)PXB, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapL2promotion promo = (C % KC) ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
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
    cuuint64_t dims[3] = {(cuuint64_t)a.w_ld, (cuuint64_t)a.w_rows, (cuuint64_t)(a.K * a.K)};
    cuuint64_t strides[2] = {(cuuint64_t)a.w_ld * 2, (cuuint64_t)a.w_rows * a.w_ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)(p.NOUT / 2), 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a.w), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NERVECL_EUNSUPPORTED;
  }

  Pair2Args t;
  t.kind = kind; t.relu = a.relu; t.alpha = a.alpha; t.bias = a.bias;
  t.eop = nullptr; t.ldeop = 0;
  if (kind == 2) { t.eop = (const bf16*)a.mask; t.ldeop = a.ldmask; }
  else if (a.accumulate) { t.eop = (const bf16*)a.out; t.ldeop = a.ldo; }
  else if (a.res) { t.eop = (const bf16*)a.res; t.ldeop = a.ldres; }
  t.out = (bf16*)a.out; t.ldo = a.ldo; t.colsum = a.colsum;
  t.sbits = (uint16_t*)a.sign_bits; t.smode = a.sign_mode; t.npix = (int64_t)a.N * a.H * a.W;
  t.v256_out = a.ldo % 16 == 0 && aligned(a.out, 32);
  t.v256_in = t.eop && t.ldeop % 16 == 0 && aligned(t.eop, 32);
  t.N = a.N; t.H = a.H; t.W = a.W; t.Cout = a.Cout;
  t.NOUT = p.NOUT; t.nchunks = p.nchunks; t.ksteps_last = p.ksteps_last; t.nchunks2 = p.nchunks2;
  t.ksteps2_last = p.ksteps2_last; t.x2_center = a.x2 ? a.x2_center : 0;
  t.strips = p.strips; t.cps = p.cps; t.stages = p.stages; t.slots = p.slots; t.R = p.R;

  const dim3 grid((unsigned)(2 * p.pairs), (unsigned)p.nsplit, 1);
  cudaError_t e = cudaSuccess;
#define NV_LAUNCH_PAIR(NCV)                                                                                          \
  do {                                                                                                               \
    e = cudaFuncSetAttribute(conv_rows2_kernel<NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);       \
    if (e != cudaSuccess) return (int)e;                                                                             \
    conv_rows2_kernel<NCV><<<grid, kThreads, p.smem, s>>>(tx, tx2, tw, t);                                            \
  } while (0)
  if (p.NOUT <= 32) NV_LAUNCH_PAIR(1); else if (p.NOUT <= 64) NV_LAUNCH_PAIR(2); else NV_LAUNCH_PAIR(3);
#undef NV_LAUNCH_PAIR
  return launch_status();
}

}  // namespace nv
