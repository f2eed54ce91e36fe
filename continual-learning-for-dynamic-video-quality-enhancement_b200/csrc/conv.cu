// ABI entry points for dense convolution: argument validation + engine dispatch.
#include "common.cuh"
#include "conv_internal.cuh"
#include <cstdlib>

using namespace nv;

NV_API int nervecl_has_tcgen05(void) { return 1; }

static int validate(const nervecl_conv_params* p) {
  if (!p || !p->x || !p->w || !p->out) return NERVECL_EINVAL;
  if (p->N <= 0 || p->H <= 0 || p->W <= 0 || p->Cin <= 0 || p->Cout <= 0) return NERVECL_EINVAL;
  if (p->K != 1 && p->K != 3 && p->K != 7) return NERVECL_EUNSUPPORTED;
  if (p->ldx < p->Cin || p->ldo < p->Cout || p->w_ld < p->Cin || p->w_rows < p->Cout) return NERVECL_EINVAL;
  if (p->res && p->ldres < (p->res_channels < p->Cout ? p->res_channels : p->Cout)) return NERVECL_EINVAL;
  if (p->dtype != NERVECL_F32 && p->dtype != NERVECL_BF16) return NERVECL_EDTYPE;
  if (p->out_dtype != NERVECL_F32 && p->out_dtype != NERVECL_BF16) return NERVECL_EDTYPE;
  if (p->dtype == NERVECL_F32 && p->out_dtype == NERVECL_BF16) return NERVECL_EDTYPE;
  return NERVECL_OK;
}

NV_API int nervecl_conv2d_fwd(const nervecl_conv_params* p, nervecl_stream_t stream) {
  int rc = validate(p);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  int engine = p->engine;
  const bool pair_ok = engine != NERVECL_CONV_TC_ROWS1;      // (ROWS1: the 1-CTA row kernel, for A/B comparisons and tests)
  if (engine == NERVECL_CONV_TC_ROWS1) engine = NERVECL_CONV_TC;
  if (p->sign_mode) {                            // packed ReLU signs: CTA-pair row kernel only
    if (p->sign_mode != 1 && p->sign_mode != 2) return NERVECL_EINVAL;
    if (!p->sign_bits || (p->Cout & 15)) return NERVECL_EINVAL;
    if (engine != NERVECL_CONV_AUTO && engine != NERVECL_CONV_TC) return NERVECL_EUNSUPPORTED;
    if (!pair_ok || !conv_rows2_supported(*p)) return NERVECL_EUNSUPPORTED;
    return conv_rows2_fwd(*p, s);
  }
  if (p->x2) {                                   // second input: row-streaming engines only
    if (p->Cin2 <= 0 || p->ldx2 < p->Cin2) return NERVECL_EINVAL;
    if (engine != NERVECL_CONV_AUTO && engine != NERVECL_CONV_TC) return NERVECL_EUNSUPPORTED;
    if (pair_ok && conv_rows2_supported(*p)) return conv_rows2_fwd(*p, s);
    if (!conv_rows_supported(*p)) return NERVECL_EUNSUPPORTED;
    return conv_rows_fwd(*p, s);
  }
  if (p->colsum) {                               // fused column sums: row-streaming engines only
    if (engine != NERVECL_CONV_AUTO && engine != NERVECL_CONV_TC) return NERVECL_EUNSUPPORTED;
    if (pair_ok && conv_rows2_supported(*p)) return conv_rows2_fwd(*p, s);
    if (!conv_rows_supported(*p)) return NERVECL_EUNSUPPORTED;
    return conv_rows_fwd(*p, s);
  }
  if (engine == NERVECL_CONV_AUTO) engine = conv_tc_fwd_supported(*p) ? NERVECL_CONV_TC : NERVECL_CONV_SIMT;
  if (engine == NERVECL_CONV_TC) {               // best tcgen05 kernel for the shape
    if (pair_ok && conv_rows2_supported(*p)) return conv_rows2_fwd(*p, s);
    // (1x1: the row kernel wins while the per-row MMA count stays small; wide inputs are at the HBM roofline
    //  with the per-tap kernel already)
    static const int k1_max_cin = nv::tune_env("NERVECL_ROWS_K1_MAXCIN") ? atoi(nv::tune_env("NERVECL_ROWS_K1_MAXCIN")) : 128;
    if ((p->K == 3 || p->Cin <= k1_max_cin) && conv_rows_supported(*p)) return conv_rows_fwd(*p, s);
    if (!conv_tc_fwd_supported(*p)) return NERVECL_EUNSUPPORTED;
    return conv_tc_fwd(*p, s);
  }
  if (engine == NERVECL_CONV_TC_TAPS) {          // per-tap kernel only (1x1, small shapes; A/B comparisons)
    if (!conv_tc_fwd_supported(*p)) return NERVECL_EUNSUPPORTED;
    return conv_tc_fwd(*p, s);
  }
  if (engine == NERVECL_CONV_SIMT) return conv_simt_fwd(*p, s);
  return NERVECL_EINVAL;
}

NV_API int nervecl_conv2d_wgrad(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype, float* dw,
                                float* db, int N, int H, int W, int Cin, int Cout, int K, float scale,
                                int engine, nervecl_stream_t stream) {
  if (!x || !dy || !dw) return NERVECL_EINVAL;
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || ldx < Cin || ldy < Cout) return NERVECL_EINVAL;
  if (K != 1 && K != 3 && K != 7) return NERVECL_EUNSUPPORTED;
  cudaStream_t s = as_stream(stream);
  bool tc_ok = conv_tc_wgrad_supported(x, ldx, dy, ldy, dtype, N, H, W, Cin, Cout, K);
  if (engine == NERVECL_CONV_AUTO) engine = tc_ok ? NERVECL_CONV_TC : NERVECL_CONV_SIMT;
  if (engine == NERVECL_CONV_TC && K == 3 && Cin >= 32 && Cout % 4 == 0 &&
      wgrad_rows_supported(x, ldx, dy, ldy, dtype, N, H, W, Cin, Cout)) {
    const int32_t col0 = 0, ncols = Cout, cin = Cin;
    float* dwp = dw;
    float* dbp = db;
    return nervecl_conv3x3_wgrad_grouped(x, ldx, dy, ldy, dtype, N, H, W, Cin, Cout, 1, &col0, &ncols, &cin, &dwp,
                                         db ? &dbp : nullptr, scale, stream);
  }
  if (engine == NERVECL_CONV_TC || engine == NERVECL_CONV_TC_TAPS) {
    if (!tc_ok) return NERVECL_EUNSUPPORTED;
    return conv_tc_wgrad(x, ldx, dy, ldy, dw, db, N, H, W, Cin, Cout, K, scale, s);
  }
  if (engine == NERVECL_CONV_SIMT) return conv_simt_wgrad(x, ldx, dy, ldy, dtype, dw, db, N, H, W, Cin, Cout, K, scale, s);
  return NERVECL_EINVAL;
}
