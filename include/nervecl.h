/*
 * nervecl.h -- C ABI of libnervecl.so, the sm_100a kernel library behind the drop-in
 * SuperResolutionNet / EWC modules (BASELINE.json north_star).
 *
 * The reference (manikya7022/Continual-Learning-for-Dynamic-Video-Quality-Enhancement) has no FFI
 * of its own: its hot path is a chain of ATen calls issued from Python.  Each entry point below
 * therefore cites the reference *call site(s)* whose arithmetic it replaces (paths relative to the
 * reference checkout).  A maintainer binds these with ctypes.
 *
 * Conventions
 *  - Every function returns 0 on success, a negative NERVECL_E* code for bad arguments, or a
 *    positive cudaError_t from the launch.  Nothing throws or exits across the ABI.
 *  - All pointers are DEVICE pointers unless the name ends in _host.  The caller owns every buffer;
 *    the library never allocates, frees or retains a pointer past return and keeps no mutable global
 *    state, so calls are thread-safe and CUDA-graph-capture-safe.  Work is enqueued on `stream`.
 *  - Activations are NHWC ("pixel-major"): element (n,y,x,c) of a tensor lives at
 *    base[((n*H + y)*W + x) * ld + c], where `ld` (the pixel pitch, in elements) may exceed the
 *    channel count so that an op can read or write a channel slice of a wider buffer (this is how
 *    torch.cat at super_resolution.py:249,252 and torch.stack at :194 are eliminated).
 *  - dtype: NERVECL_F32 or NERVECL_BF16 for activations; statistics, flow, logits, Fisher, and all
 *    parameter gradients are always fp32.
 */
#ifndef NERVECL_H_
#define NERVECL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* nervecl_stream_t; /* cudaStream_t */

enum { NERVECL_F32 = 0, NERVECL_BF16 = 1 };

enum {
  NERVECL_OK = 0,
  NERVECL_EINVAL = -1,      /* bad shape / null pointer */
  NERVECL_EALIGN = -2,      /* pointer or pitch not aligned as required */
  NERVECL_EDTYPE = -3,      /* unsupported dtype */
  NERVECL_EUNSUPPORTED = -4 /* shape not supported by the requested engine */
};

/* which convolution engine to use */
enum {
  NERVECL_CONV_AUTO = 0, /* tcgen05 when the shape qualifies, else SIMT */
  NERVECL_CONV_SIMT = 1, /* fp32-accumulate CUDA-core direct conv (any shape, f32 or bf16) */
  NERVECL_CONV_TC = 2,   /* tcgen05/TMEM/TMA implicit GEMM (bf16, Cin%8==0, Cout <= 256): the CTA-pair row-streaming
                            kernel for 3x3 / bf16 out / lean epilogue, the 1-CTA row kernel for other 3x3 / narrow
                            1x1 with Cin%16==0 and W>=64, else the per-tap kernel */
  NERVECL_CONV_TC_TAPS = 3, /* force the per-tap tcgen05 kernel */
  NERVECL_CONV_TC_ROWS1 = 4 /* like TC, but never the CTA-pair (cta_group::2) row kernel: A/B comparisons */
};

int nervecl_abi_version(void);
const char* nervecl_error_string(int code);
/* compile-time facts about the library: 1 if built with the tcgen05 path */
int nervecl_has_tcgen05(void);

/* ------------------------------------------------------------------------------------------
 * Layout
 * ---------------------------------------------------------------------------------------- */

/* (B,T,C,H,W) fp32 frames with arbitrary batch/frame/channel/row strides (unit column stride; a
 * stride-0 `expand`ed T as in experiments/train_baseline.py:82 is fine) -> frame-major NHWC
 * [T][B][H][W][ldd] in `dtype`; channels C..ldd-1 are written as zero (the head conv reads 8-channel
 * pixels so that its rows are 16-byte aligned for TMA).  Replaces the per-frame slicing
 * lr_frames[:, t] at super_resolution.py:346-349. */
int nervecl_pack_frames(const float* src, int64_t sB, int64_t sT, int64_t sC, int64_t sH,
                        void* dst, int64_t ldd, int dtype, int B, int T, int C, int H, int W,
                        nervecl_stream_t stream);

/* Same frames, unfolded for the 3x3 head convolution (feature_extractor.head.0, super_resolution.py:40):
 *   dst[t][b][y][x][c*9 + ky*3 + kx] = src[b][t][c][y+ky-1][x+kx-1]   (zero outside the frame),
 * channels 9C..ldd-1 zero.  The column order is the OIHW flattening of the filter, so the head conv is a
 * 1x1 convolution with weight.view(Cout, 9C) over this buffer and its weight gradient a 1x1 weight gradient
 * (a 3-channel 3x3 filter wastes > 80 % of every 16-channel MMA k-step; 27 of 32 unfolded channels are live). */
int nervecl_pack_frames_unfold3(const float* src, int64_t sB, int64_t sT, int64_t sC, int64_t sH,
                                void* dst, int64_t ldd, int dtype, int B, int T, int C, int H, int W,
                                nervecl_stream_t stream);

/* Gradient-side 3x3 unfold of a narrow NHWC tensor (C <= 3 live channels of pitch lds):
 *   dst[n][y][x][o*9 + ky*3 + kx] = src[n][y-(ky-1)][x-(kx-1)][o]   (zero outside the image), columns 9C..ldd-1 zero.
 * With it the weight gradient of a 3x3 conv with 2-3 OUTPUT channels (flow_net.6, attention.4) is one 1x1
 * weight-gradient GEMM  dW[o][c][tap] = sum_q x[q][c] * dst[q][o*9+tap]  instead of nine taps of an MMA whose
 * N dimension is 87 % padding; column o*9+4 (the centre tap) sums to the bias gradient. */
int nervecl_unfold3_grad(const void* src, int64_t lds, int src_dtype, int C, void* dst, int64_t ldd,
                         int dst_dtype, int N, int H, int W, nervecl_stream_t stream);

/* NHWC (dtype) channel slice -> NCHW fp32 contiguous.  Used only to hand intermediates back to
 * Python for return_intermediate=True (super_resolution.py:384-389) and by tests. */
int nervecl_nhwc_to_nchw(const void* src, int64_t ld, int dtype, float* dst,
                         int N, int C, int H, int W, nervecl_stream_t stream);
/* NCHW fp32 contiguous -> NHWC slice in dtype (tests, and dY entry for per-op backward). */
int nervecl_nchw_to_nhwc(const float* src, void* dst, int64_t ld, int dtype,
                         int N, int C, int H, int W, nervecl_stream_t stream);

/* OIHW fp32 master weight -> packed [tap][rows_pad][cols_pad] in `dtype`, rows = O, cols = I, zero
 * filled beyond.  transpose_flip=1 packs the data-gradient operator instead: rows = I, cols = O and
 * taps rotated by 180 degrees, so conv2d_fwd on dY with these weights is conv2d's input gradient. */
int nervecl_pack_conv_weight(const float* w_oihw, void* dst, int dtype, int O, int I, int KH,
                             int KW, int rows_pad, int cols_pad, int transpose_flip,
                             nervecl_stream_t stream);

/* The same for n weights in (n + 47) / 48 launches: host arrays of length n (dst_host[i] is [K*K][rows_pad][cols_pad]
 * in `dtype`).  Used once per step for all forward and data-gradient operators. */
int nervecl_pack_conv_weights_batched(int n, const float* const* w_host, void* const* dst_host,
                                      const int32_t* O_host, const int32_t* I_host, const int32_t* K_host,
                                      const int32_t* rows_pad_host, const int32_t* cols_pad_host,
                                      const int32_t* flip_host, int dtype, nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense convolution (stride 1, "same" zero padding, odd square kernel) with fused epilogue.
 * Replaces every nn.Conv2d on the path: super_resolution.py:40-43,74-82,167-174,234-242,308-311;
 * efficient_layers.py:50-56 (pointwise), :94-100 (upsampler conv), :196-197 (7x7) -- and, called
 * on dY with transpose_flip weights, ATen's convolution_backward input gradient.
 *
 *   v = sum_taps sum_ci x[p+tap, ci] * w[tap, co, ci]   (+ sum over the x2 channels when x2 != NULL)
 *   v += bias[co]                      (bias != NULL)
 *   v  = max(v, 0)                     (relu == 1)
 *   v *= alpha
 *   v += res[p, co]                    (res != NULL and co < res_channels)
 *   v += out[p, co]                    (accumulate)
 *   v  = max(v, 0)                     (relu == 2: ReLU after the residual, efficient_layers.py:148-150)
 *   v  = 0 if mask[p, co] - (mask_sub ? mask_sub[p, co] : 0) <= 0   (mask != NULL and co >= mask_c0)
 *   out[p, co] = v
 * ---------------------------------------------------------------------------------------- */
typedef struct nervecl_conv_params {
  int32_t N, H, W, Cin, Cout, K;       /* K = kernel size (1,3,7) */
  int32_t w_ld;                        /* packed weight row length (cols_pad of pack_conv_weight) */
  int32_t w_rows;                      /* packed weight rows per tap (rows_pad >= Cout) */
  int32_t dtype;                       /* dtype of x, w, res, mask */
  int32_t out_dtype;                   /* dtype of out (may be F32 while dtype is BF16) */
  int32_t engine;                      /* NERVECL_CONV_* */
  int32_t relu;
  int32_t accumulate;
  int32_t res_channels;
  int32_t mask_c0;
  float alpha;
  const void* x;   int64_t ldx;
  const void* w;                       /* packed by nervecl_pack_conv_weight */
  const float* bias;
  const void* res; int64_t ldres;
  const void* mask; int64_t ldmask;
  const void* mask_sub; int64_t ldmask_sub;
  void* out;       int64_t ldo;
  /* ABI v2: optional second input, a virtual channel concat [x | x2] (no copy).  Its weights are columns
   * [ceil(Cin/64)*64, +Cin2) of the packed weight rows.  With x2_center != 0 only the centre tap of those
   * columns is used (a fused 1x1 branch).  Row-streaming tcgen05 engine only (3x3, bf16). */
  const void* x2;  int64_t ldx2;
  int32_t Cin2;
  int32_t x2_center;
  /* ABI v3: optional fp32 [Cout] vector that receives (+=, atomically) the per-channel sum over all pixels of
   * the values written to `out` -- when `out` is the output gradient of the previous layer (a data-gradient
   * conv gated by that layer's ReLU mask) this IS that layer's bias gradient, so no separate reduction pass over
   * the gradient is needed.  Row-streaming engine, mask-gated bf16 outputs with <= 32 channels per CTA only
   * (NERVECL_EUNSUPPORTED otherwise). */
  float* colsum;
  /* ABI v7: packed ReLU sign bits, one uint16 per (16-channel group, pixel), group-major: bit j of
   * sign_bits[g * N*H*W + p] <-> channel 16 g + c with j = (c odd ? 15 - c/2 : 7 - c/2) (a warp's 32 pixels are 64
   * contiguous bytes; the bit order is what eight packed bf16x2 words give in three instructions each).  sign_mode 1: WRITE the signs of the written outputs (out > 0;
   * a forward conv with relu == 1).  sign_mode 2: READ them as the ReLU mask (v = 0 where the bit is clear; `mask` must
   * be NULL) -- the mask of relu'(y) of super_resolution.py:237 costs 2 bytes per pixel and 16 channels instead of 32.
   * CTA-pair row kernel (3x3, bf16, Cout % 16 == 0); NERVECL_EUNSUPPORTED otherwise. */
  void* sign_bits;
  int32_t sign_mode;
  int32_t reserved0;
} nervecl_conv_params;

int nervecl_conv2d_fwd(const nervecl_conv_params* p, nervecl_stream_t stream);

/* Weight (+bias) gradient of the same convolution, ACCUMULATED into fp32 OIHW `dw` (and `db`):
 *   dw[co,ci,ky,kx] += scale * sum_p dy[p,co] * x[p+tap,ci] ;  db[co] += scale * sum_p dy[p,co]
 * i.e. ATen convolution_backward's weight/bias outputs, written straight into param.grad layout.
 * workspace: at least nervecl_conv2d_wgrad_workspace() bytes (may be 0). */
int nervecl_conv2d_wgrad(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype,
                         float* dw, float* db, int N, int H, int W, int Cin, int Cout, int K,
                         float scale, int engine, nervecl_stream_t stream);

/* Weight (+bias) gradients of `ngroups` 3x3 convolutions that read channel PREFIXES [0, cin_g) of one input
 * buffer x (Cx channels) and whose output gradients are the channel slices [col0_g, col0_g + ncols_g) of one
 * buffer dy (Cy <= 160 channels) -- the five layers of a ResidualDenseBlock (super_resolution.py:234-252;
 * ATen convolution_backward's weight/bias outputs for each of them) in one tcgen05 GEMM:
 *   dw_g[o, c, ky, kx] += scale * sum_p dy[p, col0_g + o] * x[p + (ky-1, kx-1), c] ;  db_g[o] += scale * sum_p dy
 * Host arrays have ngroups entries, sorted by ascending cin; dw_g is fp32 OIHW [ncols_g][cin_g][3][3];
 * db_host may be NULL (or hold NULL entries).  bf16 only. */
int nervecl_conv3x3_wgrad_grouped(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype,
                                  int N, int H, int W, int Cx, int Cy, int ngroups,
                                  const int32_t* col0_host, const int32_t* ncols_host,
                                  const int32_t* cin_host, float* const* dw_host, float* const* db_host,
                                  float scale, nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Depthwise 3x3 (groups=C, no bias): efficient_layers.py:37-46,63.
 * w is the parameter itself, fp32 [C][1][3][3].  flip=1 applies the 180-degree rotated filter
 * (input gradient).  accumulate=1 adds into y.
 * ---------------------------------------------------------------------------------------- */
int nervecl_dwconv3x3_fwd(const void* x, int64_t ldx, const float* w, void* y, int64_t ldy,
                          int dtype, int N, int H, int W, int C, int flip, int accumulate,
                          nervecl_stream_t stream);
int nervecl_dwconv3x3_wgrad(const void* x, int64_t ldx, const void* dy, int64_t ldy, int dtype,
                            float* dw, int N, int H, int W, int C, nervecl_stream_t stream);
/* ABI v7: y = (dwconv3x3(x) + add) where mask > 0, else 0 -- the depthwise data gradient of the extractor's first body
 * layer, the skip gradient of super_resolution.py:53 (feat = body(head) + head) and the ReLU of the head conv (:40-43)
 * in one pass (bf16, TMA-tiled kernel only; y may not alias add or mask). */
int nervecl_dwconv3x3_fwd_masked(const void* x, int64_t ldx, const float* w, const void* add, int64_t ldadd,
                                 const void* mask, int64_t ldmask, void* y, int64_t ldy, int dtype,
                                 int N, int H, int W, int C, int flip, nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm2d + ReLU over `groups` consecutive image groups that each have their own batch
 * statistics (the T per-frame calls of the shared extractor, super_resolution.py:346-349 with
 * efficient_layers.py:59,65-66).  npix = pixels per group (B*H*W).
 * ---------------------------------------------------------------------------------------- */
/* sums[g][c][0..1] += (sum x, sum x^2) in float64; caller zeroes `sums`. */
int nervecl_bn_stats(const void* x, int64_t ldx, int dtype, int C, int64_t npix, int groups,
                     double* sums, nervecl_stream_t stream);
/* training=1: stat[g][c] = (mean, invstd) from sums; running stats updated in group order with
 * `momentum` (unbiased variance), *num_batches_tracked += groups.
 * training=0: stat[g][c] = (running_mean, rsqrt(running_var+eps)) for every g; buffers untouched. */
int nervecl_bn_finalize(const double* sums, float* stat, float* running_mean, float* running_var,
                        int64_t* num_batches_tracked, int C, int64_t npix, int groups,
                        float momentum, float eps, int training, nervecl_stream_t stream);
/* y = relu((x-mean)*invstd*gamma+beta) (+ res) */
int nervecl_bn_relu_fwd(const void* x, int64_t ldx, const float* stat, const float* gamma,
                        const float* beta, const void* res, int64_t ldres, void* y, int64_t ldy,
                        int dtype, int C, int64_t npix, int groups, nervecl_stream_t stream);
/* g = dy * [bn(x) > 0];  bsums[g][c] += (sum g, sum g*xhat) in float64; caller zeroes. */
int nervecl_bn_relu_bwd_reduce(const void* x, int64_t ldx, const void* dy, int64_t lddy,
                               const float* stat, const float* gamma, const float* beta,
                               int dtype, int C, int64_t npix, int groups, double* bsums,
                               nervecl_stream_t stream);
/* dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) (training) or gamma*invstd*g (eval);
 * dgamma[c] += sum_g bsums[g][c][1], dbeta[c] += sum_g bsums[g][c][0]. */
int nervecl_bn_relu_bwd_apply(const void* x, int64_t ldx, const void* dy, int64_t lddy,
                              const float* stat, const float* gamma, const float* beta,
                              const double* bsums, void* dx, int64_t lddx, float* dgamma,
                              float* dbeta, int dtype, int C, int64_t npix, int groups,
                              int training, nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 81-displacement correlation (LiteFlowNetCorrelation.forward, efficient_layers.py:328-343):
 *   out[p, i*9+j] = (1/C) sum_c x1[p,c] * x2[p + (i-4, j-4), c]   (zero outside the image)
 * channels 81..Cout_pad-1 of out are written as zero.
 * ---------------------------------------------------------------------------------------- */
int nervecl_corr_fwd(const void* x1, int64_t ld1, const void* x2, int64_t ld2, void* out,
                     int64_t ldo, int dtype, int N, int H, int W, int C, int cout_pad,
                     nervecl_stream_t stream);
/* dx1 / dx2 are fp32-accumulated and then stored (acc1/acc2 = 0) or added (= 1) in dtype.
 * workspace (nullable, ABI v6): N*H*W*96 elements of dtype, 16-byte aligned, for the second operand's view of the
 * gradient (dout gathered from the neighbouring pixels); with it the bf16 / C = 64 case runs both gradients as
 * tcgen05 GEMMs over 8 x 16 pixel tiles, without it as banded mma.sync products. */
int nervecl_corr_bwd(const void* x1, int64_t ld1, const void* x2, int64_t ld2, const void* dout,
                     int64_t lddo, void* dx1, int64_t lddx1, int acc1, void* dx2, int64_t lddx2,
                     int acc2, int dtype, int N, int H, int W, int C, void* workspace,
                     int64_t workspace_bytes, nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Flow warp = grid build + bilinear grid_sample(zeros, align_corners=True)
 * (warp_features, super_resolution.py:104-143).  flow is fp32 NHWC [N][H][W][2] = (dx, dy).
 * The coordinate arithmetic replays the reference's fp32 op sequence without contraction:
 *   g  = (2*(x + fx)) {* (1/(W-1)) | / (W-1)} - 1      div_mode 0: ATen-CUDA, 1: ATen-CPU
 *   ix = ((g + 1) / 2) * (W-1);  x0 = floor(ix)
 * so the four neighbour indices are bit-identical to ATen's.  idx_out (nullable) receives
 * int32 [N][H][W][2] = (x0, y0).
 * ---------------------------------------------------------------------------------------- */
int nervecl_warp_fwd(const void* feat, int64_t ldf, const float* flow, void* out, int64_t ldo,
                     int dtype, int N, int H, int W, int C, int div_mode, int32_t* idx_out,
                     nervecl_stream_t stream);
/* dfeat is fp32 [N][H][W][C] (pitch lddf) and is ACCUMULATED with atomics (caller zeroes);
 * dflow fp32 [N][H][W][2] is overwritten. */
int nervecl_warp_bwd(const void* feat, int64_t ldf, const float* flow, const void* dout,
                     int64_t lddo, float* dfeat, int64_t lddf, float* dflow, int dtype, int N,
                     int H, int W, int C, int div_mode, nervecl_stream_t stream);
/* Same, with dfeat in the ACTIVATION dtype: for bf16 the scatter uses packed 8 x bf16 reductions (half the L2
 * reduction traffic of the fp32 form; every partial sum is rounded to bf16).  fp32 activations: identical to
 * nervecl_warp_bwd. */
int nervecl_warp_bwd_lp(const void* feat, int64_t ldf, const float* flow, const void* dout,
                        int64_t lddo, void* dfeat, int64_t lddf, float* dflow, int dtype, int N,
                        int H, int W, int C, int div_mode, nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Temporal fusion: softmax over the T logits + attention-weighted sum of the T aligned feature
 * maps (super_resolution.py:173-174,200-204).  feats is the frame-major concat buffer: frame t's
 * channels are [t*C, (t+1)*C) of a pitch-ldf buffer.  attn (fp32 [npix][T]) is saved for backward.
 * ---------------------------------------------------------------------------------------- */
int nervecl_tfuse_fwd(const void* feats, int64_t ldf, const float* logits, float* attn, void* out,
                      int64_t ldo, int dtype, int64_t npix, int T, int C, nervecl_stream_t stream);
/* dfeats[p, t*C+c] = attn[p,t]*d[p,c] (stored);  dlogits = softmax backward of <d, feat_t>;
 * d = dout (+ nc_bias[n][c] when nc_bias != NULL, pix_per_image pixels per n). */
int nervecl_tfuse_bwd(const void* feats, int64_t ldf, const float* attn, const void* dout,
                      int64_t lddo, const float* nc_bias, int64_t pix_per_image, void* dfeats,
                      int64_t lddf, float* dlogits, int dtype, int64_t npix, int T, int C,
                      nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * CBAM (efficient_layers.py:154-228): channel attention then spatial attention.
 * ---------------------------------------------------------------------------------------- */
/* out[n][c] += scale * sum_p x[n,p,c]  (caller zeroes; global average pool, :177) */
int nervecl_chan_sum(const void* x, int64_t ldx, int dtype, int N, int64_t pix_per_image, int C,
                     float scale, float* out, nervecl_stream_t stream);
/* hidden = relu(pool @ w1^T) [N][R]; gate = sigmoid(hidden @ w2^T) [N][C]   (:170-179) */
int nervecl_ca_gate_fwd(const float* pool, const float* w1, const float* w2, float* hidden,
                        float* gate, int N, int C, int R, nervecl_stream_t stream);
/* dpool[n][c], dw1 += , dw2 += from dgate   */
int nervecl_ca_gate_bwd(const float* pool, const float* w1, const float* w2, const float* hidden,
                        const float* gate, const float* dgate, float* dpool, float* dw1,
                        float* dw2, int N, int C, int R, nervecl_stream_t stream);
/* stats[p] = (mean_c, max_c) of x[p,c]*gate[n,c]  (:200-203) */
int nervecl_cbam_stats_fwd(const void* x, int64_t ldx, const float* gate, float* stats, int dtype,
                           int N, int64_t pix_per_image, int C, nervecl_stream_t stream);
/* sgate[p] = sigmoid(conv7x7(stats)[p]);  out[p,c] = x[p,c]*gate[n,c]*sgate[p]   (:204-205) */
int nervecl_cbam_apply_fwd(const void* x, int64_t ldx, const float* gate, const float* stats,
                           const float* w7, float* sgate, void* out, int64_t ldo, int dtype, int N,
                           int H, int W, int C, nervecl_stream_t stream);
/* dz[p] = sgate(1-sgate) * sum_c dy[p,c]*x[p,c]*gate[n,c] */
int nervecl_cbam_bwd_dz(const void* x, int64_t ldx, const float* gate, const float* sgate,
                        const void* dy, int64_t lddy, float* dz, int dtype, int N,
                        int64_t pix_per_image, int C, nervecl_stream_t stream);
/* dstats = conv7x7^T(dz);  dw7[2][7][7] += sum_p dz[p]*stats[p+tap] */
int nervecl_cbam_bwd_spatial(const float* dz, const float* stats, const float* w7, float* dstats,
                             float* dw7, int N, int H, int W, nervecl_stream_t stream);
/* dxs = dy*sgate + dstats.mean/C + dstats.max*[c==argmax];  dx = dxs*gate (stored);
 * dgate[n][c] += sum_p dxs*x  (caller zeroes) */
int nervecl_cbam_bwd_dx(const void* x, int64_t ldx, const float* gate, const float* sgate,
                        const float* stats, const float* dstats, const void* dy, int64_t lddy,
                        void* dx, int64_t lddx, float* dgate, int dtype, int N,
                        int64_t pix_per_image, int C, nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Output stage: PixelShuffle + bicubic skip + clamp (efficient_layers.py:101-105,
 * super_resolution.py:378-382).  conv_out is the upsampler conv result, fp32 NHWC [N][H][W][3*s*s];
 * lr is the centre LR frame (fp32, strides in elements, unit column stride);
 * out is fp32 NCHW contiguous [N][3][s*H][s*W].
 * ---------------------------------------------------------------------------------------- */
int nervecl_upfinish_fwd(const float* conv_out, const float* lr, int64_t sN, int64_t sC, int64_t sH,
                         float* out, int N, int C, int H, int W, int s, nervecl_stream_t stream);
/* dconv[p, c*s*s + i*s + j] = dout[n,c,y*s+i,x*s+j] * [0 <= bicubic + conv_out <= 1] */
int nervecl_upfinish_bwd(const float* conv_out, const float* lr, int64_t sN, int64_t sC,
                         int64_t sH, const float* dout, float* dconv, int N, int C, int H, int W,
                         int s, nervecl_stream_t stream);

/* out (N,C,sH,sW fp32, in place) = strength * out + (1 - strength) * bicubic(lr)  -- the EnhancementEngine's
 * strength blend, nerve_cl/models/enhancement_engine.py:172-182 (ATen bicubic, A = -0.75, align_corners=False);
 * lr is (N,C,H,W) fp32 with element strides sN, sC, sH and unit column stride. */
int nervecl_bicubic_blend(float* out, const float* lr, int64_t sN, int64_t sC, int64_t sH, int N, int C,
                          int H, int W, int s, float strength, nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * FrameRecoveryNet trunk (nerve_cl/models/frame_recovery.py), inference path.  Its stride-1 3x3 / 1x1
 * convolutions (94 % of the FLOPs: TemporalConv3D's (1,3,3) spatial and (3,1,1) temporal convs,
 * efficient_layers.py:231-294, the pointwise convs of ResidualBlock :109-151, FusionModule.align, and
 * ConvTranspose2d(4,2,1) rewritten as a 3x3 conv with 4*Cout outputs) run through nervecl_conv2d_fwd with
 * the eval-mode BatchNorm folded into weights and bias; the entry points below are the rest.
 * ---------------------------------------------------------------------------------------- */
/* Direct convolution with stride / padding, OIHW fp32 weights read as they are (no packing), + bias, ReLU:
 * the encoder stem Conv2d(4,64,7,2,3) + folded BN + ReLU (frame_recovery.py:43-47) and the 1x1 stride-2
 * shortcuts of SpatialEncoder._make_stage (:70-74).  out is [N][OH][OW][ldo], OH = (H + 2 pad - K)/stride + 1. */
int nervecl_conv2d_direct(const void* x, int64_t ldx, int dtype, const float* w_oihw, const float* bias,
                          void* out, int64_t ldo, int out_dtype, int N, int H, int W, int Cin, int Cout,
                          int K, int stride, int pad, int relu, nervecl_stream_t stream);
/* nn.MaxPool2d(k, stride, pad) (frame_recovery.py:47) and F.max_pool3d(x, (1,2,2)) (:150,153) over NHWC frames */
int nervecl_maxpool2d(const void* x, int64_t ldx, void* y, int64_t ldy, int dtype, int N, int H, int W,
                      int C, int k, int stride, int pad, nervecl_stream_t stream);
/* y[n][s*h+i][s*w+j][c] = x[n][h][w][(i*s+j)*C + c]: second half of ConvTranspose2d(4,2,1) (:281-303) */
int nervecl_depth_to_space(const void* x, int64_t ldx, void* y, int64_t ldy, int dtype, int N, int H, int W,
                           int C, int s, nervecl_stream_t stream);
/* F.interpolate(mode='bilinear', align_corners=False) to (OH, OW) (frame_recovery.py:224-230) */
int nervecl_resize_bilinear(const void* x, int64_t ldx, void* y, int64_t ldy, int dtype, int N, int H, int W,
                            int C, int OH, int OW, nervecl_stream_t stream);
/* FusionModule (:232-256): a = softmax(logits[p][0:2]); out[p][c] = aligned[p][c] + a0 * mean_c spatial[p] +
 * a1 * mean_c temporal[p]  (the two all-ones/C 1x1 convolutions of :243-250 are channel means). */
int nervecl_fusion_blend(const void* aligned, int64_t lda, const float* logits, int64_t ldl,
                         const void* spatial, int64_t lds, int Cs, const void* temporal, int64_t ldt, int Ct,
                         void* out, int64_t ldo, int dtype, int C, int64_t npix, nervecl_stream_t stream);
/* Tail (:425-442): recovered = tanh(conv_out [N][Hd][Wd][ldc] fp32), bilinearly resized to (H, W) when the sizes
 * differ; out = frame * (1 - mask) + recovered * mask.  frame / out (N,C,H,W) fp32, mask (N,1,H,W) fp32 or NULL. */
int nervecl_recovery_finish(const float* conv_out, int64_t ldc, int Hd, int Wd, const float* frame,
                            const float* mask, float* out, int N, int C, int H, int W,
                            nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Elementwise helpers (gradient routing that autograd does implicitly in the reference)
 * ---------------------------------------------------------------------------------------- */
/* out = (accumulate ? out : 0) + alpha * x      over npix x C with pitches */
int nervecl_axpy(const void* x, int64_t ldx, int x_dtype, void* out, int64_t ldo, int out_dtype,
                 int64_t npix, int C, float alpha, int accumulate, nervecl_stream_t stream);
/* out = dy * [(y - (y_sub ? y_sub : 0)) > 0]  (threshold_backward of the in-place ReLUs) */
int nervecl_relu_bwd(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* y_sub,
                     int64_t ldys, void* out, int64_t ldo, int dtype, int64_t npix, int C,
                     nervecl_stream_t stream);
int nervecl_fill_zero(void* p, size_t bytes, nervecl_stream_t stream);
/* loss += scale * sum (a-b)^2 over n fp32 elements (fp32 result, caller zeroes); dgrad (nullable)
 * receives 2*scale*(a-b): nn.MSELoss fwd+bwd of experiments/train_baseline.py:64,86-87 in one pass */
int nervecl_mse_fwd_bwd(const float* a, const float* b, float* dgrad, float* loss, int64_t n,
                        float scale, nervecl_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * EWC (nerve_cl/continual/ewc.py).  All state is flat fp32; tensors of the model are addressed
 * through host arrays of device pointers so arbitrary nn.Modules work without a copy.
 * ---------------------------------------------------------------------------------------- */
/* fisher[off_i + k] += scale * g_i[k]^2   (ewc.py:139-141; the final 1/N of :146-147 is `scale`
 * on the last call or a separate nervecl_ewc_scale) */
int nervecl_ewc_fisher_accum(float* fisher, const float* const* grads_host,
                             const int64_t* numel_host, int ntensors, float scale,
                             nervecl_stream_t stream);
/* v[k] = a*v[k] + b*w[k]   (online consolidation, ewc.py:186-190; w may be NULL => pure scale) */
int nervecl_ewc_axpby(float* v, const float* w, int64_t n, float a, float b,
                      nervecl_stream_t stream);
/* *out += coef * sum_i sum_k F[off_i+k] * (theta_i[k] - star[off_i+k])^2   (ewc.py:226-232);
 * caller zeroes *out; partial sums are accumulated in float64 inside the kernel. */
int nervecl_ewc_penalty_fwd(const float* const* theta_host, const int64_t* numel_host,
                            int ntensors, const float* fisher, const float* star, float coef,
                            float* out, nervecl_stream_t stream);
/* grad_i[k] += (*gscale) * coef2 * F*(theta-theta*)  -- autograd of the penalty; gscale is the
 * device scalar grad_output (NULL => 1). */
int nervecl_ewc_penalty_bwd(const float* const* theta_host, float* const* grad_host,
                            const int64_t* numel_host, int ntensors, const float* fisher,
                            const float* star, float coef2, const float* gscale,
                            nervecl_stream_t stream);

/* flat[offset_i + k] = src_i[k] for every non-NULL src_i: gathers per-parameter tensors (e.g. .grad tensors that autograd
 * summed out of place) into one flat buffer with explicit slot offsets, so the fused optimiser step stays ONE launch. */
int nervecl_flat_gather(const float* const* src_host, const int64_t* numel_host,
                        const int64_t* offset_host, int ntensors, float* flat,
                        nervecl_stream_t stream);

/* Synaptic Intelligence (nerve_cl/continual/ewc.py:306-379) on flat fp32 state laid out like the EWC
 * buffers (tensor i at offset sum of the numels before it).
 * update (ewc.py:342-352, after every optimiser step): for every tensor whose grad pointer is non-NULL
 *   W[off_i+k] += -g_i[k] * (theta_i[k] - p_old[off_i+k]);  p_old[off_i+k] = theta_i[k]
 * (a NULL grad -- param.grad is None -- leaves both untouched, like the reference). */
int nervecl_si_update(const float* const* theta_host, const float* const* grad_host,
                      const int64_t* numel_host, int ntensors, float* W, float* p_old,
                      nervecl_stream_t stream);
/* register_task (ewc.py:354-366): omega += W / ((theta - p_old)^2 + damping); W = 0; p_old = theta.
 * The SI penalty  si_lambda * sum omega (theta - p_old)^2  (ewc.py:368-379) is
 * nervecl_ewc_penalty_fwd/bwd with fisher = omega, star = p_old, coef = si_lambda, coef2 = 2 si_lambda. */
int nervecl_si_register(const float* const* theta_host, const int64_t* numel_host, int ntensors,
                        float* W, float* p_old, float* omega, float damping,
                        nervecl_stream_t stream);

/* Fused AdamW over a flat fp32 parameter buffer (torch.optim.AdamW semantics,
 * experiments/train_baseline.py:62).  step is 1-based; step_dev (nullable, ABI v6) is a device int32 holding the
 * step count instead -- the bias corrections are then computed in the kernel, so a CUDA graph that contains the
 * launch (and the increment of that counter) replays correctly. */
int nervecl_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                       int64_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int step, float grad_scale, const int32_t* step_dev,
                       nervecl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NERVECL_H_ */
