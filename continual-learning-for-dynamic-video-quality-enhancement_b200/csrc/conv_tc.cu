// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 in, fp32 accumulate in TMEM).
//
// GEMM view of conv2d (stride 1, "same" zero padding) on NHWC activations:
//     D[pixel, co] = sum over k-blocks (tap, 64- or 32-channel chunk) of  A[pixel + tap, ci] * B[tap][co][ci]
//   M tile = 128 output pixels (BH x BW patch, BH*BW = 128),  N = Cout rounded up to 16 (<= 256, one N tile),
//   K loop = K*K taps x ceil(Cin / KC) chunks.
// There is no im2col buffer: the A tile of one (tap, chunk) is ONE 4-D TMA box {KC, BW, BH, 1} of the
// activation tensor at coordinates {c0, x0 + kx - r, y0 + ky - r, n}; TMA's out-of-bounds zero fill
// provides the conv padding, the channel tail and the ragged right/bottom tiles.  The box lands in
// shared memory as 128 rows of KC bf16 with the 128B (KC=64) / 64B (KC=32) swizzle, which is exactly the
// canonical K-major UMMA operand layout, so tcgen05.mma consumes it through a shared-memory descriptor.
//
// Persistent, warp-specialised CTA (one per SM): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM
// owner), warps 2..5 = epilogue (TMEM -> registers -> fused bias/ReLU/scale/residual/accumulate/mask ->
// global).  Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include "conv_internal.cuh"
#include "conv_tc_epilogue.cuh"
#include "tc_common.cuh"

using namespace nv;
using namespace nv::tc;

namespace {

constexpr int kEpiWarps = 8;                       // 2 warps per TMEM lane quarter, splitting the columns
constexpr int kThreads = 32 * (2 + kEpiWarps);     // warp 0 = TMA, warp 1 = MMA, warps 2..9 = epilogue

struct TcArgs : EpiArgs {
  int N, H, W;
  int BN;          // Cout rounded up to 16
  int K;           // kernel size
  int nchunks;     // ceil(Cin / KC)
  int bw_shift;    // BW = 1 << bw_shift, BH = 128 >> bw_shift
  int tiles_x, tiles_y;
  int stages;
};

template <int KC, typename OutT>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const TcArgs a) {
  constexpr uint32_t ROW_BYTES = KC * 2;
  constexpr uint32_t A_BYTES = BM * ROW_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A tile | B tile)] 1024-aligned, then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t b_bytes = (uint32_t)a.BN * ROW_BYTES;
  const uint32_t stage_bytes = (A_BYTES + b_bytes + 1023u) & ~1023u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
  uint64_t* full_bar = bars;                       // [stages]  TMA -> MMA
  uint64_t* empty_bar = bars + a.stages;           // [stages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * a.stages;       // [2]       MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = a.N * a.tiles_y * a.tiles_x;
  const int kblocks = a.K * a.K * a.nchunks;
  const uint32_t tmem_cols = a.BN * 2 <= 32 ? 32 : a.BN * 2 <= 64 ? 64 : a.BN * 2 <= 128 ? 128 : a.BN * 2 <= 256 ? 256 : 512;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);     // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int R = a.K / 2;
  const int BW = 1 << a.bw_shift;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int tx = tile % a.tiles_x;
        int r = tile / a.tiles_x;
        int ty = r % a.tiles_y;
        int n = r / a.tiles_y;
        int x0 = tx << a.bw_shift, y0 = ty * (BM >> a.bw_shift);
        for (int kb = 0; kb < kblocks; ++kb) {
          int tap = kb / a.nchunks, chunk = kb - tap * a.nchunks;
          int ky = tap / a.K, kx = tap - ky * a.K;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], A_BYTES + b_bytes);
          tma_load_4d(sa, &tmap_x, &full_bar[stage], chunk * KC, x0 + kx - R, y0 + ky - R, n);
          tma_load_3d(sa + A_BYTES, &tmap_w, &full_bar[stage], chunk * KC, 0, tap);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // warp-uniform control flow, one elected lane issues (keeps descriptors in uniform registers; see
    // conv_tc_rows.cu)
    {
      const uint32_t idesc = make_idesc_bf16((uint32_t)a.BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      bool ready = false;     // result of an early, non-blocking probe of the next stage's barrier
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          if (!ready) mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t da = make_kmajor_desc(sa, ROW_BYTES);
          const uint64_t db = make_kmajor_desc(sa + A_BYTES, ROW_BYTES);
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == a.stages) { nstage = 0; nphase ^= 1; }
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              // advance 16 elements (32 B) along K inside the swizzle span: +2 in the (>>4) address field
              umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            }
            umma_commit(&empty_bar[stage]);                  // frees the smem stage when these MMAs retire
            if (kb == kblocks - 1) umma_commit(&tfull_bar[acc]);   // accumulator ready for the epilogue
          }
          __syncwarp();
          // probe the next stage now: its latency overlaps the MMAs just issued
          ready = mbar_try_wait(&full_bar[nstage], nphase);
          stage = nstage;
          phase = nphase;
        }
      }
    }
  } else {
    // ================= epilogue (warps 2..9) =================
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;                      // which half of the 16-column chunks
    const int row = q * 32 + lane;                         // tile row == pixel within the tile
    const int ty_in = row >> a.bw_shift, tx_in = row & (BW - 1);
    const int nch = a.BN >> 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int tx = tile % a.tiles_x;
      int r = tile / a.tiles_x;
      int ty = r % a.tiles_y;
      int n = r / a.tiles_y;
      const int y = ty * (BM >> a.bw_shift) + ty_in, x = (tx << a.bw_shift) + tx_in;
      const bool valid = (y < a.H) && (x < a.W);
      const int64_t p = ((int64_t)n * a.H + y) * a.W + x;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * a.BN);
      for (int c = part; c < nch; c += 4) {
        EpiChunk<OutT> e0, e1;
        const bool two = c + 2 < nch;
        e0.issue(a, taddr, c * 16, valid, p);
        if (two) e1.issue(a, taddr, (c + 2) * 16, valid, p);
        tmem_ld_wait();
        e0.finish(a, c * 16, valid, p);
        if (two) e1.finish(a, (c + 2) * 16, valid, p);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

inline int pick_kc(int cin) { return cin >= 64 ? 64 : 32; }   // channel tails are zero-filled by TMA

}  // namespace

namespace nv {

bool conv_tc_fwd_supported(const nervecl_conv_params& a) {
  if (a.dtype != NERVECL_BF16) return false;
  if (a.out_dtype != NERVECL_BF16 && a.out_dtype != NERVECL_F32) return false;
  if (a.K != 1 && a.K != 3) return false;
  if (a.Cin < 8 || a.Cin % 8 || a.Cout < 1 || a.Cout > 256) return false;
  if (a.ldx % 8 || a.w_ld % 8 || !aligned(a.x, 16) || !aligned(a.w, 16)) return false;
  if (a.Cout >= 16) {   // vector epilogue: 16 consecutive channels per thread
    const int v = a.out_dtype == NERVECL_BF16 ? 8 : 4;
    if (a.ldo % v || !aligned(a.out, 16)) return false;
  }
  if (a.bias && !aligned(a.bias, 16)) return false;
  if (a.res && (a.ldres % 8 || !aligned(a.res, 16))) return false;
  if (a.mask && (a.ldmask % 8 || !aligned(a.mask, 16))) return false;
  if (a.mask_sub && (a.ldmask_sub % 8 || !aligned(a.mask_sub, 16))) return false;
  if (a.accumulate && a.out_dtype != NERVECL_BF16) return false;
  if (a.w_rows < a.Cout) return false;
  if ((int64_t)a.N * a.H * a.W < 128) return false;
  return encode_fn() != nullptr;
}

int conv_tc_fwd(const nervecl_conv_params& a, cudaStream_t s) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return NERVECL_EUNSUPPORTED;
  const int KC = pick_kc(a.Cin);
  const int BN = (a.Cout + 15) / 16 * 16;
  const int bw_shift = tile_bw_shift(a.W);
  const int BW = 1 << bw_shift, BH = BM >> bw_shift;
  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

  CUtensorMap tx, tw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N};
    cuuint64_t strides[3] = {(cuuint64_t)a.ldx * 2, (cuuint64_t)a.W * a.ldx * 2, (cuuint64_t)a.H * a.W * a.ldx * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)BW, (cuuint32_t)BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    // (channel count not a multiple of the 64-channel box: promote to 64 bytes only, or the last box drags the
    //  buffer's pad channels in from DRAM -- the 224-channel dense-block buffer was read as 256)
    if (enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.x), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, (a.Cin % 64) ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NERVECL_EUNSUPPORTED;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.w_ld, (cuuint64_t)a.w_rows, (cuuint64_t)(a.K * a.K)};
    cuuint64_t strides[2] = {(cuuint64_t)a.w_ld * 2, (cuuint64_t)a.w_rows * a.w_ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a.w), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NERVECL_EUNSUPPORTED;
  }

  TcArgs t;
  t.N = a.N; t.H = a.H; t.W = a.W; t.Cout = a.Cout; t.BN = BN; t.K = a.K;
  t.nchunks = (a.Cin + KC - 1) / KC;
  t.bw_shift = bw_shift;
  t.tiles_x = (a.W + BW - 1) / BW;
  t.tiles_y = (a.H + BH - 1) / BH;
  t.relu = a.relu; t.accumulate = a.accumulate; t.res_channels = a.res ? a.res_channels : 0; t.mask_c0 = a.mask_c0;
  t.alpha = a.alpha;
  t.bias = a.bias;
  t.res = (const bf16*)a.res; t.ldres = a.ldres;
  t.mask = (const bf16*)a.mask; t.ldmask = a.ldmask;
  t.msub = (const bf16*)a.mask_sub; t.ldmsub = a.ldmask_sub;
  t.out = a.out; t.ldo = a.ldo;
  t.colsum = nullptr; t.v256_out = 0; t.v256_in = 0; t.v256_gen = epi_v256(a);

  const size_t stage_bytes = ((size_t)BM * KC * 2 + (size_t)BN * KC * 2 + 1023) & ~(size_t)1023;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return NERVECL_EUNSUPPORTED;
  t.stages = stages;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 4) * sizeof(uint64_t) + 16;

  const int64_t num_tiles = (int64_t)a.N * t.tiles_y * t.tiles_x;
  const int grid = (int)imin(num_tiles, sm_count());
  cudaError_t e = cudaSuccess;
#define NV_LAUNCH_TC(KCV, OT)                                                                                   \
  do {                                                                                                          \
    e = cudaFuncSetAttribute(conv_tc_kernel<KCV, OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    if (e != cudaSuccess) return (int)e;                                                                        \
    conv_tc_kernel<KCV, OT><<<grid, kThreads, smem, s>>>(tx, tw, t);                                            \
  } while (0)
  const bool f32out = a.out_dtype == NERVECL_F32;
  if (KC == 64 && !f32out) NV_LAUNCH_TC(64, bf16);
  else if (KC == 64) NV_LAUNCH_TC(64, float);
  else if (!f32out) NV_LAUNCH_TC(32, bf16);
  else NV_LAUNCH_TC(32, float);
#undef NV_LAUNCH_TC
  return launch_status();
}

}  // namespace nv
