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
      // region row qy = 2q + s goes through slab (s & 1); step i reads slabs i (lanes 0-15) and i + 1 (lanes 16-31)
      auto stage_row = [&](int s) {
        const uint32_t col = (uint32_t)((2 * q + s) * RX);
        uint32_t v[16], w[8];
        tmem_ld16(lane_addr + col, v);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "r"(lane_addr + col + 16u)
                     : "memory");
        tmem_ld_wait();
        float* dst = slab + (s & 1) * (RX * 32) + lane;
#pragma unroll
        for (int c = 0; c < 16; ++c) dst[c * 32] = __uint_as_float(v[c]);
#pragma unroll
        for (int c = 0; c < 8; ++c) dst[(16 + c) * 32] = __uint_as_float(w[c]);
      };
      stage_row(0);
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        __syncwarp();                                  // readers of the slab about to be overwritten are done
        stage_row(i + 1);
        __syncwarp();
        const float* src = slab + ((i + pyl) & 1) * (RX * 32) + px * 32 + lane;
#pragma unroll
        for (int j = 0; j < ND; ++j) res[i * ND + j] = src[j * 32] * inv_c;
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
