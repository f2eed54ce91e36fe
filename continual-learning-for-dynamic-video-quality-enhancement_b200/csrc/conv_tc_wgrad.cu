// tcgen05 weight-gradient kernel:  dW[co][ci][tap] += scale * sum_p dY[p][co] * X[p + tap][ci]
//
// GEMM view: D[M][N] += A[M][K] * B[N][K]^T with K = pixels, M = (tap, 64-channel chunk of Cin) units taken
// two at a time (UMMA M = 128), N = Cout.  Both operands are "MN-major": a TMA box {64 ch, BW, BH, 1} of the
// NHWC tensor lands in shared memory as 128 pixel rows of channels, i.e. K runs along the rows and M (or
// N) is contiguous -- the canonical MN-major swizzled UMMA layout -- so the same zero-filling TMA boxes as
// the forward kernel feed the tensor core with no transpose.  The X box of tap (ky,kx) is the pixel tile
// shifted by (ky-r, kx-r); out-of-image rows are zero, which is exactly conv padding.
//
// The whole dW slice of a CTA stays resident in TMEM (<= 512 fp32 columns) while it streams over its share
// of the pixel tiles; only at the end are the accumulators drained with fp32 atomics into the OIHW
// gradient (param.grad layout).  Layers whose dW does not fit 512 columns are split over "M groups" of CTAs.
#include "conv_internal.cuh"
#include "tc_common.cuh"

using namespace nv;
using namespace nv::tc;

extern "C" int nervecl_chan_sum(const void* x, int64_t ldx, int dtype, int N, int64_t pix_per_image, int C,
                                float scale, float* out, nervecl_stream_t stream);

namespace {

constexpr int kThreads = 192;
constexpr uint32_t XBOX_BYTES = BM * 128;       // 128 pixels x 64 bf16

// MN-major operand: rows = K (pixels) of `row_bytes` (one swizzle span), LBO = stride between MN blocks
// (next 64- or 32-channel box), SBO = stride between groups of 8 K rows.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t row_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_bf16_mn(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// bias gradient for channel counts the vectorised chan_sum kernel does not take (C % 4 != 0, C <= 32)
__global__ void __launch_bounds__(256)
colsum_small_kernel(const bf16* __restrict__ dy, int64_t ldy, int64_t npix, int C, float scale, float* __restrict__ db) {
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (c < C) acc[c] += ldf(dy + p * ldy + c);
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    if (c < C) {
      float v = warp_sum(acc[c]);
      if ((threadIdx.x & 31) == 0) atomicAdd(db + c, scale * v);
    }
  }
}

struct WgArgs {
  int N, H, W, Cin, Cout, K;
  int NM;            // UMMA N (Cout rounded up to 16)
  int nb_blocks;     // dY boxes per tile (1, or 2 for Cout = 128)
  int dy_row_bytes;  // 64 (SW64, Cout <= 32) or 128
  int nci;           // 64-channel chunks of Cin
  int units;         // K*K*nci
  int pairs;         // ceil(units / 2)
  int G;             // pairs per M group (G * NM <= 512)
  int MG;            // number of M groups
  int bw_shift, tiles_x, tiles_y;
  int stages;
  int ndy;           // dY tile buffers in flight (2 .. 6)
  float scale;
  float* dw;
  float* db;         // bias gradient taken in-kernel from the dY tiles (drain warps of M group 0), or nullptr
};

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                     const WgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = 2 * XBOX_BYTES;
  const uint32_t dyblk_bytes = BM * (uint32_t)a.dy_row_bytes;
  const uint32_t dy_bytes = dyblk_bytes * (uint32_t)a.nb_blocks;
  uint8_t* dy_smem = smem + (size_t)a.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dy_smem + (size_t)a.ndy * dy_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + a.stages;
  uint64_t* dyfull_bar = bars + 2 * a.stages;      // [ndy]
  uint64_t* dyempty_bar = dyfull_bar + a.ndy;      // [ndy]
  uint64_t* done_bar = dyempty_bar + a.ndy;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mg = blockIdx.x % a.MG, split = blockIdx.x / a.MG, nsplit = gridDim.x / a.MG;
  const int pair0 = mg * a.G;
  const int npairs = min(a.G, a.pairs - pair0);
  const int num_tiles = a.N * a.tiles_y * a.tiles_x;
  const uint32_t need_cols = (uint32_t)(a.G * a.NM);
  const uint32_t tmem_cols = need_cols <= 32 ? 32 : need_cols <= 64 ? 64 : need_cols <= 128 ? 128 : need_cols <= 256 ? 256 : 512;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_dy);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < a.ndy; ++s) {
      mbar_init(&dyfull_bar[s], 1);
      mbar_init(&dyempty_bar[s], a.db ? 5 : 1);      // the MMA commit (+ the four column-sum warps)
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int R = a.K / 2;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = split; tile < num_tiles; tile += nsplit, ++it) {
        int tx = tile % a.tiles_x;
        int r = tile / a.tiles_x;
        int ty = r % a.tiles_y;
        int n = r / a.tiles_y;
        int x0 = tx << a.bw_shift, y0 = ty * (BM >> a.bw_shift);
        const int db = it % a.ndy;
        mbar_wait(&dyempty_bar[db], ((it / a.ndy) & 1) ^ 1);
        mbar_expect_tx(&dyfull_bar[db], dy_bytes);
        for (int b = 0; b < a.nb_blocks; ++b)
          tma_load_4d(dy_smem + (size_t)db * dy_bytes + (size_t)b * dyblk_bytes, &tmap_dy, &dyfull_bar[db], b * 64, x0,
                      y0, n);
        for (int pl = 0; pl < npairs; ++pl) {
          const int u0 = 2 * (pair0 + pl);
          const int nbox = (u0 + 1 < a.units) ? 2 : 1;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], nbox * XBOX_BYTES);
          for (int h = 0; h < nbox; ++h) {
            const int u = u0 + h;
            const int tap = u / a.nci, chunk = u - tap * a.nci;
            const int ky = tap / a.K, kx = tap - ky * a.K;
            tma_load_4d(sa + (size_t)h * XBOX_BYTES, &tmap_x, &full_bar[stage], chunk * 64, x0 + kx - R, y0 + ky - R, n);
          }
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_mn((uint32_t)a.NM);
      const uint32_t b_kstep = 16u * (uint32_t)a.dy_row_bytes;      // 16 pixel rows of the dY tile
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = split; tile < num_tiles; tile += nsplit, ++it) {
        const int db = it % a.ndy;
        mbar_wait(&dyfull_bar[db], (it / a.ndy) & 1);
        tc_fence_after();
        const uint32_t sb = smem_u32(dy_smem + (size_t)db * dy_bytes);
        for (int pl = 0; pl < npairs; ++pl) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t d_tmem = tmem_base + (uint32_t)(pl * a.NM);
#pragma unroll
          for (int k = 0; k < BM / 16; ++k) {
            const uint64_t da = make_mnmajor_desc(sa + (uint32_t)k * 2048u, 128, XBOX_BYTES);
            const uint64_t dbd = make_mnmajor_desc(sb + (uint32_t)k * b_kstep, (uint32_t)a.dy_row_bytes, dyblk_bytes);
            umma_bf16(d_tmem, da, dbd, idesc, (it | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&dyempty_bar[db]);
      }
      umma_commit(done_bar);
    }
  } else {
    // drain: TMEM -> fp32 atomics into OIHW dW
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int half = row >> 6, ci_local = row & 63;
    const bool any_tile = split < num_tiles;
    if (a.db) {
      // Bias gradient db[co] = scale * sum_p dY[p, co] while the main loop runs: these warps are idle until the drain,
      // and every dY tile passes through shared memory anyway (128 pixel rows of 64 channels, 128B swizzle) -- a
      // separate reduction pass would read dY from HBM a second time.  Lane <-> channel pair, warp <-> 32 pixel rows.
      float s0 = 0.f, s1 = 0.f;
      int it = 0;
      for (int tile = split; tile < num_tiles; tile += nsplit, ++it) {
        const int db = it % a.ndy;
        mbar_wait(&dyfull_bar[db], (it / a.ndy) & 1);
        if (mg == 0) {
          const uint8_t* t = dy_smem + (size_t)db * dy_bytes;
#pragma unroll 8
          for (int r = q * 32; r < q * 32 + 32; ++r) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(t + r * 128 + ((((uint32_t)lane >> 2) ^ ((uint32_t)r & 7u)) << 4) +
                                                                   (((uint32_t)lane & 3u) << 2));
            s0 += __uint_as_float(w << 16);
            s1 += __uint_as_float(w & 0xFFFF0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&dyempty_bar[db]);
      }
      if (mg == 0 && any_tile) {
        if (2 * lane < a.Cout) atomicAdd(a.db + 2 * lane, a.scale * s0);
        if (2 * lane + 1 < a.Cout) atomicAdd(a.db + 2 * lane + 1, a.scale * s1);
      }
    }
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int KK = a.K * a.K;
    for (int pl = 0; pl < npairs; ++pl) {
      const int u = 2 * (pair0 + pl) + half;
      const int tap = u / a.nci, chunk = u - tap * a.nci;
      const int ci = chunk * 64 + ci_local;
      const bool row_ok = any_tile && u < a.units && ci < a.Cin;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pl * a.NM);
      for (int c0 = 0; c0 < a.NM; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = c0 + j;
            if (co < a.Cout) atomicAdd(a.dw + ((int64_t)co * a.Cin + ci) * KK + tap, a.scale * __uint_as_float(v[j]));
          }
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace

namespace nv {

bool conv_tc_wgrad_supported(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype, int N, int H, int W,
                             int Cin, int Cout, int K) {
  if (dtype != NERVECL_BF16) return false;
  if (K != 1 && K != 3) return false;
  if (Cin < 1 || Cout < 1 || Cout > 128 || (Cout > 32 && Cout % 64)) return false;
  if (ldx % 8 || ldy % 8 || !aligned(x, 16) || !aligned(dy, 16)) return false;
  if ((int64_t)N * H * W < 128) return false;
  return encode_fn() != nullptr;
}

int conv_tc_wgrad(const void* x, int64_t ldx, const void* dy, int64_t ldy, float* dw, float* db, int N, int H, int W,
                  int Cin, int Cout, int K, float scale, cudaStream_t s) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return NERVECL_EUNSUPPORTED;
  const int bw_shift = tile_bw_shift(W);
  const int BW = 1 << bw_shift, BH = BM >> bw_shift;
  WgArgs a;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.K = K;
  a.NM = (Cout + 15) / 16 * 16;
  a.dy_row_bytes = Cout <= 32 ? 64 : 128;
  a.nb_blocks = Cout <= 64 ? 1 : (Cout + 63) / 64;
  a.nci = (Cin + 63) / 64;
  a.units = K * K * a.nci;
  a.pairs = (a.units + 1) / 2;
  a.G = (int)imin(a.pairs, 512 / a.NM);
  a.MG = (a.pairs + a.G - 1) / a.G;
  a.bw_shift = bw_shift;
  a.tiles_x = (W + BW - 1) / BW;
  a.tiles_y = (H + BH - 1) / BH;
  a.scale = scale;
  a.dw = dw;
  // in-kernel bias gradient: one 128-byte dY box per tile (Cout in (32, 64]); out-of-image pixels are zero-filled
  a.db = (db && a.dy_row_bytes == 128 && a.nb_blocks == 1) ? db : nullptr;

  CUtensorMap tx, td;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)W * ldx * 2, (cuuint64_t)H * W * ldx * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            (Cin % 64) ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,   // (no pad channels from DRAM)
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NERVECL_EUNSUPPORTED;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ldy * 2, (cuuint64_t)W * ldy * 2, (cuuint64_t)H * W * ldy * 2};
    cuuint32_t box[4] = {(cuuint32_t)(a.dy_row_bytes / 2), (cuuint32_t)BW, (cuuint32_t)BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&td, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE,
            a.dy_row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NERVECL_EUNSUPPORTED;
  }

  const size_t dy_bytes = (size_t)BM * a.dy_row_bytes * a.nb_blocks;
  // X stages: one per (tap, chunk) pair in flight.  A CTA with ONE pair per tile (1x1 convs over <= 128 channels)
  // runs through a tile per stage, so two dY buffers cap it at two tiles in flight: five there (0.53 -> 0.49 ms for the
  // extractor's pointwise / head weight gradients); with two pairs per tile the extra X stage is worth more.
  const int pairs_per_tile = a.G < a.pairs ? a.G : a.pairs;
  a.ndy = pairs_per_tile == 1 ? 5 : 2;
  int stages = (int)((208 * 1024 - a.ndy * dy_bytes) / (2 * XBOX_BYTES));
  if (stages > 6) stages = 6;
  if (stages < 2) { a.ndy = 2; stages = (int)((208 * 1024 - 2 * dy_bytes) / (2 * XBOX_BYTES)); if (stages > 6) stages = 6; }
  if (stages < 2) return NERVECL_EUNSUPPORTED;
  a.stages = stages;
  const size_t smem = 1024 + (size_t)stages * 2 * XBOX_BYTES + a.ndy * dy_bytes + (2 * stages + 2 * a.ndy + 1) * sizeof(uint64_t) + 16;

  const int64_t num_tiles = (int64_t)N * a.tiles_y * a.tiles_x;
  const int sms = sm_count();
  int nsplit = (int)imax(1, imin(num_tiles, sms / a.MG));
  const int grid = nsplit * a.MG;
  cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  conv_tc_wgrad_kernel<<<grid, kThreads, smem, s>>>(tx, td, a);
  int rc = launch_status();
  if (rc) return rc;
  if (!db || a.db) return NERVECL_OK;
  if (Cout % 4 == 0)
    return nervecl_chan_sum(dy, ldy, NERVECL_BF16, 1, (int64_t)N * H * W, Cout, scale, db, (nervecl_stream_t)s);
  const int64_t npix = (int64_t)N * H * W;
  colsum_small_kernel<<<(int)imin(cdiv(npix, 256 * 8), sms * 4), 256, 0, s>>>((const bf16*)dy, ldy, npix, Cout, scale, db);
  return launch_status();
}

}  // namespace nv
