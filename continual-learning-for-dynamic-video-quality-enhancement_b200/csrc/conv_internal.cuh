// Internal interface between the conv ABI entry points (conv.cu) and the two engines.
#pragma once
#include "common.cuh"
#include <cstdlib>

namespace nv {

// every bf16 epilogue operand present allows 256-bit accesses (32-byte aligned base, pitch a multiple of 16 channels)
static inline int epi_v256(const nervecl_conv_params& a) {
  auto ok = [](const void* p, int64_t ld) { return !p || (ld % 16 == 0 && nv::aligned(p, 32)); };
  return a.out_dtype == NERVECL_BF16 && ok(a.out, a.ldo) && ok(a.res, a.ldres) && ok(a.mask, a.ldmask) &&
         ok(a.mask_sub, a.ldmask_sub) && !nv::tune_env("NERVECL_NO_V256");
}

int conv_simt_fwd(const nervecl_conv_params& a, cudaStream_t s);
int conv_simt_wgrad(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype, float* dw, float* db,
                    int N, int H, int W, int Cin, int Cout, int K, float scale, cudaStream_t s);

// tcgen05 / TMEM / TMA implicit GEMM (conv_tc.cu)
bool conv_tc_fwd_supported(const nervecl_conv_params& a);
int conv_tc_fwd(const nervecl_conv_params& a, cudaStream_t s);
// row-streaming 3x3 tcgen05 kernel (conv_tc_rows.cu)
bool conv_rows_supported(const nervecl_conv_params& a);
int conv_rows_fwd(const nervecl_conv_params& a, cudaStream_t s);
// CTA-pair (cta_group::2) row-streaming 3x3 kernel for bf16 outputs with a lean epilogue (conv_tc_rows2.cu)
bool conv_rows2_supported(const nervecl_conv_params& a);
int conv_rows2_fwd(const nervecl_conv_params& a, cudaStream_t s);
bool conv_tc_wgrad_supported(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype, int N, int H,
                             int W, int Cin, int Cout, int K);
int conv_tc_wgrad(const void* x, int64_t ldx, const void* dy, int64_t ldy, float* dw, float* db, int N, int H,
                  int W, int Cin, int Cout, int K, float scale, cudaStream_t s);

// grouped row-resident 3x3 weight gradient (conv_tc_wgrad_rows.cu)
bool wgrad_rows_supported(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype, int N, int H, int W, int Cx,
                          int Cy);
int wgrad_rows(const void* x, int64_t ldx, const void* dy, int64_t ldy, int N, int H, int W, int Cx, int Cy, int ngroups,
               const int32_t* col0, const int32_t* ncols, const int32_t* cin, float* const* dw, float* const* db, float scale,
               cudaStream_t s);

}  // namespace nv
