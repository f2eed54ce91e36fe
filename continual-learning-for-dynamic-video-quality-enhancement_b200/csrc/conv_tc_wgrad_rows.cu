// Grouped, row-resident tcgen05 weight gradient for 3x3 convolutions that share one input buffer.
//
//   dW_g[o, c, ky, kx] += scale * sum_p dY[p, col0_g + o] * X[p + (ky-1, kx-1), c]      c < cin_g
//
// The five 3x3 layers of a residual dense block all read prefixes of ONE 224-channel buffer and their
// output gradients are adjacent channel slices of ONE gradient buffer, so their weight gradients are one
// GEMM  D[c, (g,o)] = X^T dY  per tap with M = channels, N = 160 (all layers), K = pixels -- instead of five
// N = 32 GEMMs whose SS-mode MMAs are shared-memory bound at 40 % of the tensor rate.  (Entries with
// c >= cin_g are computed and dropped: 57 % of the dense product is used, at ~2.5x the MMA efficiency.)
//
// A CTA owns one (128-channel M tile, ky) class: its three taps kx = 0,1,2 accumulate in 3*N TMEM columns
// for the CTA's whole lifetime.  Per 128-pixel row tile TMA brings the halo'd X row y+ky-1 {64 ch x 130 px}
// per 64-channel chunk and the dY row {64 ch x 128 px} per chunk into shared memory once; both land as the
// canonical MN-major 128B-swizzled UMMA layout (K = pixel rows), and the three kx taps are three A
// descriptors offset by one pixel row (absolute-address swizzle, as in conv_tc_rows.cu).  M tiles that only
// matter to the later layers (channels >= 128 feed only layers with cin > 128) use the matching suffix of
// the dY columns.  At the end the accumulators are drained with fp32 atomics into the OIHW gradients.
#include "conv_internal.cuh"
#include "tc_common.cuh"

using namespace nv;
using namespace nv::tc;

namespace {

constexpr int kThreads = 192;                       // warp 0 TMA, warp 1 MMA, warps 2..5 drain
constexpr int KC = 64;
constexpr uint32_t ROWB = 128;
// One pipeline stage = KPX pixels of a row (the GEMM's K extent per stage).  Measured at the cfg-2 dense block
// (profiles/r02h_experiments.md): 128 px x 2 stages 1.65 ms, 64 x 5 1.99 ms, 32 x 10 3.17 ms -- smaller stages spread the
// classes over more rows at a time and the L2 hit rate drops (40 % -> 26 %), so the deeper pipeline loses.
#ifndef WG_KPX
#define WG_KPX 128
#endif
#ifndef WG_STAGES
#define WG_STAGES 2
#endif
constexpr int KPX = WG_KPX;
constexpr int PXB = KPX + 2;
constexpr uint32_t XCHUNK = (PXB * ROWB + 1023u) & ~1023u;    // 17408: one 64-channel chunk of a halo'd X row piece
constexpr uint32_t YCHUNK = KPX * ROWB;                         // 16384: one 64-channel chunk of a dY row piece
constexpr int kStages = WG_STAGES;
constexpr int kMaxGroups = 8;
constexpr int kMaxClasses = 12;

struct WgGroup { int col0, ncols, cin; float* dw; float* db; };

struct WgrArgs {
  int N, H, W, Cx, strips;
  int ngroups;
  WgGroup grp[kMaxGroups];
  int nclasses;
  // per class: M tile, ky, first dY column, UMMA N, dY chunks, first CTA of the class, CTAs in the class
  int c_mt[kMaxClasses], c_ky[kMaxClasses], c_col0[kMaxClasses], c_n16[kMaxClasses], c_nby[kMaxClasses],
      c_cta0[kMaxClasses], c_ctas[kMaxClasses];
  // an M tile with <= 64 live channels carries a SECOND kernel row in its upper 64 accumulator rows: the same
  // channels of input row y + ky2 - 1 (-1: none), so its three taps need two classes instead of three
  int c_ky2[kMaxClasses];
  int stages;
  int nby_max;         // dY chunks per stage of the widest class (sizes the stage)
  int bias;            // some group has a bias gradient: the drain warps of class 0 sum the dY tiles while the MMAs run
  float scale;
};

__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((8u * ROWB) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_mn(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_rows_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const WgrArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // which class does this CTA belong to?
  int cls = 0;
  for (int c = 1; c < a.nclasses; ++c)
    if ((int)blockIdx.x >= a.c_cta0[c]) cls = c;
  const int mt = a.c_mt[cls], ky = a.c_ky[cls], col0 = a.c_col0[cls], n16 = a.c_n16[cls], nby = a.c_nby[cls];
  const int ky2 = a.c_ky2[cls];
  const int idx = blockIdx.x - a.c_cta0[cls], nctas = a.c_ctas[cls];
  const uint32_t stage_bytes = 2u * XCHUNK + (uint32_t)a.nby_max * YCHUNK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + a.stages;
  uint64_t* done_bar = empty_bar + a.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row_tiles = (int64_t)a.N * a.H * a.strips;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_dy);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], (a.bias && cls == 0) ? 5 : 1);      // the MMA commit (+ the four column-sum warps)
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    // channel chunks of this M tile that exist (the second one may lie entirely beyond Cx: not loaded; its
    // accumulator rows are never drained)
    const int nxc = (mt * 128 + KC < a.Cx || ky2 >= 0) ? 2 : 1;
    const uint32_t bytes = (uint32_t)nxc * (PXB * ROWB) + (uint32_t)nby * YCHUNK;
    for (int64_t rt = idx; rt < row_tiles; rt += nctas) {
      if (lane == 0) {
        const int strip = (int)(rt % a.strips);
        const int64_t r = rt / a.strips;
        const int y = (int)(r % a.H), n = (int)(r / a.H);
        const int x0 = strip * KPX;
        mbar_wait(&empty_bar[stage], phase ^ 1);
#ifdef WG_NO_LOAD
        mbar_arrive(&full_bar[stage]);
#else
        mbar_expect_tx(&full_bar[stage], bytes);
        uint8_t* sx = smem + (size_t)stage * stage_bytes;
        uint8_t* sy = sx + 2 * XCHUNK;
        for (int c = 0; c < nxc; ++c)
          tma_load_4d(sx + (size_t)c * XCHUNK, &tmap_x, &full_bar[stage], mt * 128 + (ky2 >= 0 ? 0 : c * KC), x0 - 1,
                      y + (ky2 >= 0 && c ? ky2 : ky) - 1, n);
        for (int c = 0; c < nby; ++c)
          tma_load_4d(sy + (size_t)c * YCHUNK, &tmap_dy, &full_bar[stage], col0 + c * KC, x0, y, n);
#endif
      }
      if (++stage == a.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_mn((uint32_t)n16);
    const uint64_t hi = mn_desc(0, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lbo_a = ((XCHUNK >> 4) & 0x3FFF) << 16, lbo_b = ((YCHUNK >> 4) & 0x3FFF) << 16;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t started = 0;
    for (int64_t rt = idx; rt < row_tiles; rt += nctas) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sx = (smem_u32(smem + (size_t)stage * stage_bytes) & 0x3FFFFu) >> 4;
      const uint32_t sy = sx + ((2u * XCHUNK) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t d = tmem_base + (uint32_t)(kx * n16);
#pragma unroll
          for (int ks = 0; ks < KPX / 16; ++ks) {
            // K step = 16 pixel rows of 128 B; tap kx starts kx pixel rows into the halo'd X row
            const uint64_t da = hi | (uint64_t)(lbo_a | (sx + (uint32_t)((kx + 16 * ks) * 8)));
            const uint64_t db = hi | (uint64_t)(lbo_b | (sy + (uint32_t)(16 * ks * 8)));
#ifndef WG_NO_MMA
            umma_bf16(d, da, db, idesc, started | (uint32_t)ks);
#endif
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      started = 1;
      if (++stage == a.stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // drain: TMEM -> fp32 atomics into the OIHW gradients
    const int q = warp & 3;
    const bool upper = ky2 >= 0 && q >= 2;                   // this lane's accumulator row belongs to kernel row ky2
    const int c = mt * 128 + q * 32 + lane - (upper ? 64 : 0);   // input channel of this lane
    const int kyl = upper ? ky2 : ky;
    const bool any = idx < row_tiles;
    if (a.bias && cls == 0) {
      // Bias gradients db_g[o] = scale * sum_p dY[p, col0_g + o]: class 0 (M tile 0: every group feeds it) walks over
      // every dY row tile once, and these warps are idle until the drain -- so they sum the tiles in shared memory
      // instead of a second pass over dY in HBM.  Lane <-> channel pair of each 64-channel chunk, warp <-> 32 pixel rows.
      float bs[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t rt = idx; rt < row_tiles; rt += nctas) {
        mbar_wait(&full_bar[stage], phase);
        const uint8_t* sy = smem + (size_t)stage * stage_bytes + 2 * XCHUNK;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          if (cc < nby) {
#pragma unroll 8
            for (int r = q * (KPX / 4); r < (q + 1) * (KPX / 4); ++r) {
              const uint32_t w = *reinterpret_cast<const uint32_t*>(sy + (size_t)cc * YCHUNK + r * ROWB +
                                                                     ((((uint32_t)lane >> 2) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)lane & 3u) << 2));
              bs[cc][0] += __uint_as_float(w << 16);
              bs[cc][1] += __uint_as_float(w & 0xFFFF0000u);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      if (any) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          if (cc >= nby) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int col = col0 + cc * KC + 2 * lane + h;
            for (int g = 0; g < a.ngroups; ++g) {
              const int o = col - a.grp[g].col0;
              if (o >= 0 && o < a.grp[g].ncols) {
                if (a.grp[g].db) atomicAdd(a.grp[g].db + o, a.scale * bs[cc][h]);
                break;
              }
            }
          }
        }
      }
    }
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (any) {
      for (int kx = 0; kx < 3; ++kx) {
        const int tap = kyl * 3 + kx;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kx * n16);
        for (int c0 = 0; c0 < n16; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = col0 + c0 + j;
            for (int g = 0; g < a.ngroups; ++g) {
              const int o = col - a.grp[g].col0;
              if (o >= 0 && o < a.grp[g].ncols) {
                if (c < a.grp[g].cin)
                  atomicAdd(a.grp[g].dw + ((int64_t)o * a.grp[g].cin + c) * 9 + tap, a.scale * __uint_as_float(v[j]));
                break;
              }
            }
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
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

namespace nv {

bool wgrad_rows_supported(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype, int N, int H, int W, int Cx,
                          int Cy) {
  if (dtype != NERVECL_BF16) return false;
  if (Cx < 16 || Cy < 8 || Cy > 160) return false;
  if (ldx % 8 || ldy % 8 || !aligned(x, 16) || !aligned(dy, 16)) return false;
  if (W < 64 || (int64_t)N * H * W < 1024) return false;
  if ((Cx + 127) / 128 * 3 > kMaxClasses) return false;
  return encode_fn() != nullptr;
}

// groups must be sorted by ascending cin (so that the columns an M tile needs are a suffix)
int wgrad_rows(const void* x, int64_t ldx, const void* dy, int64_t ldy, int N, int H, int W, int Cx, int Cy, int ngroups,
               const int32_t* col0, const int32_t* ncols, const int32_t* cin, float* const* dw, float* const* db, float scale,
               cudaStream_t s) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return NERVECL_EUNSUPPORTED;
  if (ngroups < 1 || ngroups > kMaxGroups) return NERVECL_EINVAL;
  WgrArgs a;
  a.N = N; a.H = H; a.W = W; a.Cx = Cx; a.strips = (W + KPX - 1) / KPX;
  a.ngroups = ngroups;
  a.scale = scale;
  a.bias = 0;
  for (int g = 0; g < ngroups; ++g) {
    if (g && cin[g] < cin[g - 1]) return NERVECL_EINVAL;
    if (col0[g] < 0 || col0[g] + ncols[g] > Cy || cin[g] > Cx || !dw[g]) return NERVECL_EINVAL;
    a.grp[g] = WgGroup{col0[g], ncols[g], cin[g], dw[g], db ? db[g] : nullptr};
    a.bias = a.bias || (db && db[g]);
  }
  const int sms = sm_count();
  const int MT = (Cx + 127) / 128;
  // classes and their relative cost (MMA cycles per row tile vs bytes staged per row tile)
  double cost[kMaxClasses];
  double total_cost = 0;
  a.nclasses = 0;
  for (int mt = 0; mt < MT; ++mt) {
    // columns needed by this M tile: groups with cin > mt*128
    int lo = Cy, hi = 0;
    for (int g = 0; g < ngroups; ++g)
      if (cin[g] > mt * 128) { lo = lo < col0[g] ? lo : col0[g]; hi = hi > col0[g] + ncols[g] ? hi : col0[g] + ncols[g]; }
    if (hi <= lo) continue;
    lo = lo / 8 * 8;                                           // 16-byte aligned TMA start
    const int n16 = (hi - lo + 15) / 16 * 16;
    if (3 * n16 > 512) return NERVECL_EUNSUPPORTED;
    const bool half = Cx - mt * 128 <= KC;                     // <= 64 live channels: two kernel rows per class
    for (int ky = 0; ky < 3; ++ky) {
      if (half && ky == 1) continue;                           // (rides in the upper half of the ky = 0 class)
      const int c = a.nclasses++;
      a.c_ky2[c] = half && ky == 0 ? 1 : -1;
      a.c_mt[c] = mt; a.c_ky[c] = ky; a.c_col0[c] = lo; a.c_n16[c] = n16; a.c_nby[c] = (n16 + KC - 1) / KC;
      const double mma = (3.0 * KPX / 16.0) * n16 / 2.0;
      const double ld = (2.0 * PXB * ROWB + a.c_nby[c] * YCHUNK) / 48.0;
      cost[c] = mma > ld ? mma : ld;
      // Narrow classes (N <= 64) run 1.45x slower than this model says: their MMAs are shared-memory bound and nothing
      // they load is shared with a class in step with them.  Measured at the cfg-2 dense block (the share of SMs
      // they get, launch time): x1.0 1.70 ms, x1.2 1.51, x1.4 1.41, x1.6 1.52, x1.8 1.51, x2.1 1.62.
#ifndef WG_SMALL_SCALE
#define WG_SMALL_SCALE 1.45
#endif
      if (n16 <= 64) cost[c] *= WG_SMALL_SCALE;
      total_cost += cost[c];
    }
  }
  if (a.nclasses == 0) return NERVECL_OK;
  int used = 0;
  for (int c = 0; c < a.nclasses; ++c) {
    int n = (int)(sms * cost[c] / total_cost);
    if (n < 1) n = 1;
    a.c_ctas[c] = n;
    used += n;
  }
  for (int c = 0; used < sms; c = (c + 1) % a.nclasses) { ++a.c_ctas[c]; ++used; }   // hand out the remainder
  int acc = 0;
  for (int c = 0; c < a.nclasses; ++c) { a.c_cta0[c] = acc; acc += a.c_ctas[c]; }
  const int grid = acc;

  CUtensorMap tx, td;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cx, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)W * ldx * 2, (cuuint64_t)H * W * ldx * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)PXB, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            (Cx % KC) ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,   // (no pad channels from DRAM)
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NERVECL_EUNSUPPORTED;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cy, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ldy * 2, (cuuint64_t)W * ldy * 2, (cuuint64_t)H * W * ldy * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)KPX, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&td, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            (Cy % KC) ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NERVECL_EUNSUPPORTED;
  }
  // Stages sized for the widest class: launches with <= 128 gradient columns (every conv outside the dense blocks) get
  // 3-4 stages instead of 2 -- with two, the narrow classes' tensor work (1152 cycles per stage at N = 64) no longer
  // covered the load latency and they ran at half speed.
  a.nby_max = 1;
  for (int c = 0; c < a.nclasses; ++c) a.nby_max = a.c_nby[c] > a.nby_max ? a.c_nby[c] : a.nby_max;
  const size_t stage_sz = 2 * XCHUNK + (size_t)a.nby_max * YCHUNK;
#ifdef WG_FIXED_STAGES
  a.stages = kStages;
#else
  a.stages = (int)((220 * 1024) / stage_sz);
  if (a.stages > 4) a.stages = 4;
  if (a.stages < kStages) a.stages = kStages;
#endif
  const size_t smem = 1024 + (size_t)a.stages * stage_sz + (2 * a.stages + 1) * sizeof(uint64_t) + 16;
  cudaError_t e = cudaFuncSetAttribute(wgrad_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  wgrad_rows_kernel<<<grid, kThreads, smem, s>>>(tx, td, a);
  return launch_status();
}

}  // namespace nv

NV_API int nervecl_conv3x3_wgrad_grouped(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype, int N, int H,
                                         int W, int Cx, int Cy, int ngroups, const int32_t* col0_host,
                                         const int32_t* ncols_host, const int32_t* cin_host, float* const* dw_host,
                                         float* const* db_host, float scale, nervecl_stream_t stream) {
  if (!x || !dy || !col0_host || !ncols_host || !cin_host || !dw_host) return NERVECL_EINVAL;
  if (N <= 0 || H <= 0 || W <= 0 || Cx <= 0 || Cy <= 0 || ldx < Cx || ldy < Cy) return NERVECL_EINVAL;
  if (!wgrad_rows_supported(x, ldx, dy, ldy, dtype, N, H, W, Cx, Cy)) return NERVECL_EUNSUPPORTED;
  cudaStream_t s = as_stream(stream);
  return wgrad_rows(x, ldx, dy, ldy, N, H, W, Cx, Cy, ngroups, col0_host, ncols_host, cin_host, dw_host, db_host, scale, s);
}
