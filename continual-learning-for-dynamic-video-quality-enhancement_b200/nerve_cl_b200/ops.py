"""``torch.ops.nervecl.*`` -- the thin op layer between PyTorch tensors and the C ABI.

Each op is declared with ``torch.library`` (schema with mutation annotations), implemented for the
CUDA dispatch key by a function that extracts raw pointers / pitches and calls ``libnervecl.so`` on
the current stream, and given a Meta kernel (all ops write into caller-provided outputs, so the fake
implementations are no-ops).  No CPU kernels are registered: calling an op on CPU tensors raises.

Activation tensors are NHWC views ``[N, H, W, C]`` with unit channel stride; the pixel pitch
(``stride(2)``) may exceed ``C`` -- i.e. ``buf[..., 64:96]`` of a 224-channel buffer is a valid
operand -- which is how the reference's ``torch.cat``/``torch.stack`` copies disappear.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import BF16, CONV_AUTO, CONV_SIMT, CONV_TC, CONV_TC_ROWS1, CONV_TC_TAPS, F32, ConvParams  # noqa: F401

Tensor = torch.Tensor

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dt(t: Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"nervecl: unsupported dtype {t.dtype}") from None


def _cuda(t: Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"nervecl: `{name}` must be a CUDA tensor -- this path has no CPU implementation")


def _nhwc(t: Tensor, name: str):
    """(ptr, pitch, N, H, W, C) of an NHWC view whose pixels are uniformly pitched."""
    _cuda(t, name)
    if t.dim() != 4:
        raise RuntimeError(f"nervecl: `{name}` must be [N,H,W,C], got {tuple(t.shape)}")
    n, h, w, c = t.shape
    ld = t.stride(2)
    if t.stride(3) != 1 or (h > 1 and t.stride(1) != w * ld) or (n > 1 and t.stride(0) != h * w * ld) or ld < c:
        raise RuntimeError(f"nervecl: `{name}` is not a pitched NHWC view (shape {tuple(t.shape)}, "
                           f"strides {t.stride()})")
    return t.data_ptr(), ld, n, h, w, c


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _flat(t: Tensor, name: str, dtype=torch.float32) -> int:
    _cuda(t, name)
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"nervecl: `{name}` must be a contiguous {dtype} tensor")
    return t.data_ptr()


lib = torch.library.Library("nervecl", "DEF")
_impls = {}
LAUNCHES = [0]          # number of kernel-launching op calls (bench.py reports it as gpu_launches)
_NO_KERNEL = {"fill_zero"}   # cudaMemsetAsync, not one of our kernels


def _op(schema: str):
    """Declare ``schema`` and register the decorated function as its CUDA kernel (+ no-op Meta)."""
    name = schema.split("(", 1)[0]

    def deco(fn):
        lib.define(schema)
        if name in _NO_KERNEL:
            impl = fn
        else:
            def impl(*a, **k):
                LAUNCHES[0] += 1
                return fn(*a, **k)
        lib.impl(name, impl, "CUDA")
        lib.impl(name, lambda *a, **k: None, "Meta")
        _impls[name] = fn
        return fn
    return deco


# --------------------------------------------------------------------------------------------
# layout
# --------------------------------------------------------------------------------------------
@_op("pack_frames(Tensor src, Tensor(a!) dst) -> ()")
def _pack_frames(src: Tensor, dst: Tensor) -> None:
    """src (B,T,C,H,W) fp32 (any batch/frame/channel/row strides) -> dst [T*B,H,W,Cpad>=C], padding zeroed."""
    _cuda(src, "src")
    if src.dim() != 5 or src.dtype != torch.float32:
        raise RuntimeError("nervecl.pack_frames: src must be (B,T,C,H,W) float32")
    if src.stride(4) != 1:
        src = src.contiguous()
    b, t, c, h, w = src.shape
    dp, ld, n, dh, dw, dc = _nhwc(dst, "dst")
    if (n, dh, dw) != (t * b, h, w) or dc < c or ld != dc:
        raise RuntimeError("nervecl.pack_frames: dst must be contiguous [T*B,H,W,Cpad] with Cpad >= C")
    _lib.check(_lib.load().nervecl_pack_frames(src.data_ptr(), src.stride(0), src.stride(1), src.stride(2),
                                              src.stride(3), dp, ld, _dt(dst), b, t, c, h, w, _stream()),
               "pack_frames")


@_op("pack_frames_unfold3(Tensor src, Tensor(a!) dst) -> ()")
def _pack_frames_unfold3(src: Tensor, dst: Tensor) -> None:
    """src (B,T,C,H,W) fp32 -> dst [T*B,H,W,Cpad>=9C]: the 3x3 neighbourhood of every pixel unfolded into
    channels c*9 + ky*3 + kx (zero outside the frame), padding zeroed (see ``nervecl_pack_frames_unfold3``)."""
    _cuda(src, "src")
    if src.dim() != 5 or src.dtype != torch.float32:
        raise RuntimeError("nervecl.pack_frames_unfold3: src must be (B,T,C,H,W) float32")
    if src.stride(4) != 1:
        src = src.contiguous()
    b, t, c, h, w = src.shape
    dp, ld, n, dh, dw, dc = _nhwc(dst, "dst")
    if (n, dh, dw) != (t * b, h, w) or dc < 9 * c or ld != dc:
        raise RuntimeError("nervecl.pack_frames_unfold3: dst must be contiguous [T*B,H,W,Cpad] with Cpad >= 9*C")
    _lib.check(_lib.load().nervecl_pack_frames_unfold3(src.data_ptr(), src.stride(0), src.stride(1), src.stride(2),
                                                      src.stride(3), dp, ld, _dt(dst), b, t, c, h, w, _stream()),
               "pack_frames_unfold3")


@_op("unfold3_grad(Tensor src, Tensor(a!) dst) -> ()")
def _unfold3_grad(src: Tensor, dst: Tensor) -> None:
    """dst[p, o*9 + tap] = src[p - tap offset, o] for a narrow NHWC ``src`` (<= 3 channels), see
    ``nervecl_unfold3_grad``; ``dst`` is contiguous [N,H,W,Cpad >= 9*C], padding zeroed."""
    sp, lds, n, h, w, c = _nhwc(src, "src")
    dp, ldd, dn, dh, dw, dc = _nhwc(dst, "dst")
    if (dn, dh, dw) != (n, h, w) or ldd != dc:
        raise RuntimeError("nervecl.unfold3_grad: dst must be contiguous [N,H,W,Cpad] matching src")
    _lib.check(_lib.load().nervecl_unfold3_grad(sp, lds, _dt(src), c, dp, ldd, _dt(dst), n, h, w, _stream()),
               "unfold3_grad")


@_op("nhwc_to_nchw(Tensor src, Tensor(a!) dst) -> ()")
def _nhwc_to_nchw(src: Tensor, dst: Tensor) -> None:
    sp, ld, n, h, w, c = _nhwc(src, "src")
    if tuple(dst.shape) != (n, c, h, w):
        raise RuntimeError("nervecl.nhwc_to_nchw: shape mismatch")
    _lib.check(_lib.load().nervecl_nhwc_to_nchw(sp, ld, _dt(src), _flat(dst, "dst"), n, c, h, w, _stream()),
               "nhwc_to_nchw")


@_op("nchw_to_nhwc(Tensor src, Tensor(a!) dst) -> ()")
def _nchw_to_nhwc(src: Tensor, dst: Tensor) -> None:
    dp, ld, n, h, w, c = _nhwc(dst, "dst")
    if tuple(src.shape) != (n, c, h, w):
        raise RuntimeError("nervecl.nchw_to_nhwc: shape mismatch")
    _lib.check(_lib.load().nervecl_nchw_to_nhwc(_flat(src, "src"), dp, ld, _dt(dst), n, c, h, w, _stream()),
               "nchw_to_nhwc")


@_op("pack_conv_weight(Tensor w, Tensor(a!) dst, bool transpose_flip) -> ()")
def _pack_conv_weight(w: Tensor, dst: Tensor, transpose_flip: bool) -> None:
    """w OIHW fp32 -> dst [K*K, rows_pad, cols_pad] (dst's own shape gives the padding)."""
    o, i, kh, kw = w.shape
    if dst.dim() != 3 or dst.shape[0] != kh * kw or not dst.is_contiguous():
        raise RuntimeError("nervecl.pack_conv_weight: dst must be contiguous [K*K, rows, cols]")
    _lib.check(_lib.load().nervecl_pack_conv_weight(_flat(w, "w"), dst.data_ptr(), _dt(dst), o, i, kh, kw,
                                                   dst.shape[1], dst.shape[2], int(transpose_flip), _stream()),
               "pack_conv_weight")


@_op("pack_conv_weights_batched(Tensor[] w, Tensor(a!)[] dst, bool[] transpose_flip) -> ()")
def _pack_conv_weights_batched(w, dst, transpose_flip) -> None:
    """``pack_conv_weight`` for a list of weights in a handful of launches (all ``dst`` share one dtype)."""
    n = len(w)
    if len(dst) != n or len(transpose_flip) != n:
        raise RuntimeError("nervecl.pack_conv_weights_batched: list lengths differ")
    if n == 0:
        return
    for wi, di in zip(w, dst):
        if di.dim() != 3 or di.shape[0] != wi.shape[2] * wi.shape[3] or not di.is_contiguous() or di.dtype != dst[0].dtype:
            raise RuntimeError("nervecl.pack_conv_weights_batched: dst must be contiguous [K*K, rows, cols] of one dtype")
    vps, i32 = C.c_void_p * n, C.c_int32 * n
    _lib.check(_lib.load().nervecl_pack_conv_weights_batched(
        n, vps(*[_flat(t, "w") for t in w]), vps(*[t.data_ptr() for t in dst]), i32(*[t.shape[0] for t in w]),
        i32(*[t.shape[1] for t in w]), i32(*[t.shape[2] for t in w]), i32(*[t.shape[1] for t in dst]),
        i32(*[t.shape[2] for t in dst]), i32(*[int(f) for f in transpose_flip]), _dt(dst[0]), _stream()),
        "pack_conv_weights_batched")


# --------------------------------------------------------------------------------------------
# dense convolution
# --------------------------------------------------------------------------------------------
@_op("conv2d_fwd(Tensor x, Tensor w, Tensor? bias, Tensor? res, Tensor? mask, Tensor? mask_sub, "
     "Tensor(a!) out, int cout, int relu, bool accumulate, int res_channels, int mask_c0, float alpha, "
     "int engine, Tensor? x2=None, bool x2_center=False, Tensor(b!)? colsum=None, Tensor(c!)? sign_bits=None, "
     "int sign_mode=0) -> ()")
def _conv2d_fwd(x, w, bias, res, mask, mask_sub, out, cout, relu, accumulate, res_channels, mask_c0, alpha,
                engine, x2=None, x2_center=False, colsum=None, sign_bits=None, sign_mode=0) -> None:
    """Fused conv (see ``nervecl_conv2d_fwd``).  ``w`` is a packed [K*K, rows, cols] tensor; ``cout`` is the
    number of output channels actually computed (<= rows)."""
    xp, ldx, n, h, wd, cin = _nhwc(x, "x")
    op, ldo, on, oh, ow, oc = _nhwc(out, "out")
    if (on, oh, ow) != (n, h, wd) or oc != cout:
        raise RuntimeError(f"nervecl.conv2d_fwd: out {tuple(out.shape)} does not match x {tuple(x.shape)} / cout {cout}")
    kk = w.shape[0]
    k = int(round(kk ** 0.5))
    p = ConvParams()
    p.N, p.H, p.W, p.Cin, p.Cout, p.K = n, h, wd, cin, cout, k
    # packed weight: [K*K, rows, cols]; a row-slice view of a larger packed tensor is fine (rows per tap and the
    # row length come from the strides)
    if w.stride(2) != 1 or w.stride(0) % max(w.stride(1), 1):
        raise RuntimeError("nervecl.conv2d_fwd: packed weight must have unit column stride")
    p.w_ld, p.w_rows = w.stride(1), (w.stride(0) // w.stride(1) if kk > 1 else w.shape[1])
    p.dtype, p.out_dtype = _dt(x), _dt(out)
    if w.dtype != x.dtype:
        raise RuntimeError("nervecl.conv2d_fwd: weight dtype must equal activation dtype")
    p.engine, p.relu, p.accumulate = engine, int(relu), int(accumulate)
    p.res_channels, p.mask_c0, p.alpha = res_channels, mask_c0, alpha
    p.x, p.ldx, p.w = xp, ldx, w.data_ptr()
    p.bias = _flat(bias, "bias") if bias is not None else None
    if res is not None:
        rp, ldr, *_ = _nhwc(res, "res")
        p.res, p.ldres = rp, ldr
    if mask is not None:
        mp, ldm, *_ = _nhwc(mask, "mask")
        p.mask, p.ldmask = mp, ldm
    if mask_sub is not None:
        sp, lds, *_ = _nhwc(mask_sub, "mask_sub")
        p.mask_sub, p.ldmask_sub = sp, lds
    p.out, p.ldo = op, ldo
    if x2 is not None:
        x2p, ldx2, n2, h2, w2, c2 = _nhwc(x2, "x2")
        if (n2, h2, w2) != (n, h, wd) or x2.dtype != x.dtype:
            raise RuntimeError("nervecl.conv2d_fwd: x2 must match x in batch, size and dtype")
        p.x2, p.ldx2, p.Cin2, p.x2_center = x2p, ldx2, c2, int(x2_center)
    if colsum is not None:                       # fused per-channel sum of the written values (+=, fp32 [cout])
        if colsum.numel() < cout:
            raise RuntimeError("nervecl.conv2d_fwd: colsum must have at least cout elements")
        p.colsum = _flat(colsum, "colsum")
    if sign_mode:                                # packed ReLU signs: int16 [cout / 16, N, H, W], written (1) or read as mask (2)
        if (sign_bits is None or sign_bits.dtype != torch.int16 or not sign_bits.is_contiguous()
                or sign_bits.numel() != n * h * wd * (cout // 16) or cout % 16):
            raise RuntimeError("nervecl.conv2d_fwd: sign_bits must be a contiguous int16 [cout / 16, N, H, W] tensor")
        p.sign_bits, p.sign_mode = sign_bits.data_ptr(), int(sign_mode)
    _lib.check(_lib.load().nervecl_conv2d_fwd(C.byref(p), _stream()), "conv2d_fwd")


@_op("conv2d_wgrad(Tensor x, Tensor dy, Tensor(a!) dw, Tensor(b!)? db, float scale, int engine) -> ()")
def _conv2d_wgrad(x, dy, dw, db, scale, engine) -> None:
    """dw (OIHW fp32) += scale * dy^T x ;  db += scale * sum dy."""
    xp, ldx, n, h, w, cin = _nhwc(x, "x")
    yp, ldy, yn, yh, yw, cout = _nhwc(dy, "dy")
    if (yn, yh, yw) != (n, h, w) or x.dtype != dy.dtype:
        raise RuntimeError("nervecl.conv2d_wgrad: x / dy mismatch")
    o, i, kh, kw = dw.shape
    if (o, i) != (cout, cin) or kh != kw:
        raise RuntimeError(f"nervecl.conv2d_wgrad: dw {tuple(dw.shape)} vs Cout {cout} Cin {cin}")
    _lib.check(_lib.load().nervecl_conv2d_wgrad(xp, ldx, yp, ldy, _dt(x), _flat(dw, "dw"),
                                               _flat(db, "db") if db is not None else None, n, h, w, cin, cout, kh,
                                               scale, engine, _stream()), "conv2d_wgrad")


@_op("conv3x3_wgrad_grouped(Tensor x, Tensor dy, Tensor(a!)[] dw, Tensor(b!)[] db, int[] col0, float scale) -> ()")
def _conv3x3_wgrad_grouped(x, dy, dw, db, col0, scale) -> None:
    """Weight/bias gradients of several 3x3 convs reading channel prefixes of ``x`` whose output gradients are
    the slices ``dy[..., col0[g] : col0[g] + dw[g].shape[0]]`` (see ``nervecl_conv3x3_wgrad_grouped``).
    ``dw[g]`` is OIHW fp32 (its shape gives ncols and cin); ``db`` is empty or one fp32 vector per group."""
    xp, ldx, n, h, w, cx = _nhwc(x, "x")
    yp, ldy, yn, yh, yw, cy = _nhwc(dy, "dy")
    if (yn, yh, yw) != (n, h, w) or x.dtype != dy.dtype:
        raise RuntimeError("nervecl.conv3x3_wgrad_grouped: x / dy mismatch")
    ng = len(dw)
    if len(col0) != ng or (len(db) not in (0, ng)):
        raise RuntimeError("nervecl.conv3x3_wgrad_grouped: list lengths differ")
    for t in dw:
        if t.dim() != 4 or t.shape[2:] != (3, 3):
            raise RuntimeError("nervecl.conv3x3_wgrad_grouped: dw must be OIHW with 3x3 taps")
    i32 = C.c_int32 * ng
    vps = C.c_void_p * ng
    c0 = i32(*[int(v) for v in col0])
    nc = i32(*[int(t.shape[0]) for t in dw])
    ci = i32(*[int(t.shape[1]) for t in dw])
    dwp = vps(*[_flat(t, "dw") for t in dw])
    dbp = vps(*[_flat(t, "db") for t in db]) if len(db) else None
    _lib.check(_lib.load().nervecl_conv3x3_wgrad_grouped(xp, ldx, yp, ldy, _dt(x), n, h, w, cx, cy, ng, c0, nc, ci, dwp,
                                                        dbp, scale, _stream()), "conv3x3_wgrad_grouped")


# --------------------------------------------------------------------------------------------
# feature-extractor body
# --------------------------------------------------------------------------------------------
@_op("dwconv3x3_fwd(Tensor x, Tensor w, Tensor(a!) y, bool flip, bool accumulate) -> ()")
def _dwconv3x3_fwd(x, w, y, flip, accumulate) -> None:
    xp, ldx, n, h, wd, c = _nhwc(x, "x")
    yp, ldy, *_ = _nhwc(y, "y")
    _lib.check(_lib.load().nervecl_dwconv3x3_fwd(xp, ldx, _flat(w, "w"), yp, ldy, _dt(x), n, h, wd, c, int(flip),
                                                int(accumulate), _stream()), "dwconv3x3_fwd")


@_op("dwconv3x3_fwd_masked(Tensor x, Tensor w, Tensor add, Tensor mask, Tensor(a!) y, bool flip) -> ()")
def _dwconv3x3_fwd_masked(x, w, add, mask, y, flip) -> None:
    """y = (dwconv3x3(x) + add) where mask > 0, else 0 (``nervecl_dwconv3x3_fwd_masked``; bf16)."""
    xp, ldx, n, h, wd, c = _nhwc(x, "x")
    ap, lda, *_ = _nhwc(add, "add")
    mp, ldm, *_ = _nhwc(mask, "mask")
    yp, ldy, *_ = _nhwc(y, "y")
    _lib.check(_lib.load().nervecl_dwconv3x3_fwd_masked(xp, ldx, _flat(w, "w"), ap, lda, mp, ldm, yp, ldy, _dt(x), n, h, wd,
                                                       c, int(flip), _stream()), "dwconv3x3_fwd_masked")


@_op("dwconv3x3_wgrad(Tensor x, Tensor dy, Tensor(a!) dw) -> ()")
def _dwconv3x3_wgrad(x, dy, dw) -> None:
    xp, ldx, n, h, wd, c = _nhwc(x, "x")
    yp, ldy, *_ = _nhwc(dy, "dy")
    _lib.check(_lib.load().nervecl_dwconv3x3_wgrad(xp, ldx, yp, ldy, _dt(x), _flat(dw, "dw"), n, h, wd, c,
                                                  _stream()), "dwconv3x3_wgrad")


def _grouped(x: Tensor, groups: int, name: str):
    xp, ldx, n, h, w, c = _nhwc(x, name)
    if n % groups:
        raise RuntimeError(f"nervecl: `{name}` batch {n} not divisible into {groups} groups")
    return xp, ldx, c, (n // groups) * h * w


@_op("bn_stats(Tensor x, int groups, Tensor(a!) sums) -> ()")
def _bn_stats(x, groups, sums) -> None:
    xp, ldx, c, npix = _grouped(x, groups, "x")
    _lib.check(_lib.load().nervecl_bn_stats(xp, ldx, _dt(x), c, npix, groups, _flat(sums, "sums", torch.float64),
                                           _stream()), "bn_stats")


@_op("bn_finalize(Tensor? sums, Tensor(a!) stat, Tensor(b!) running_mean, Tensor(c!) running_var, "
     "Tensor(d!)? num_batches_tracked, int npix, int groups, float momentum, float eps, bool training) -> ()")
def _bn_finalize(sums, stat, running_mean, running_var, nbt, npix, groups, momentum, eps, training) -> None:
    c = running_mean.numel()
    _lib.check(_lib.load().nervecl_bn_finalize(
        _flat(sums, "sums", torch.float64) if sums is not None else None, _flat(stat, "stat"),
        _flat(running_mean, "running_mean"), _flat(running_var, "running_var"),
        _flat(nbt, "num_batches_tracked", torch.int64) if nbt is not None else None, c, npix, groups, momentum, eps,
        int(training), _stream()), "bn_finalize")


@_op("bn_relu_fwd(Tensor x, Tensor stat, Tensor gamma, Tensor beta, Tensor? res, Tensor(a!) y, int groups) -> ()")
def _bn_relu_fwd(x, stat, gamma, beta, res, y, groups) -> None:
    xp, ldx, c, npix = _grouped(x, groups, "x")
    yp, ldy, *_ = _nhwc(y, "y")
    rp, ldr = (None, 0)
    if res is not None:
        rp, ldr, *_ = _nhwc(res, "res")
    _lib.check(_lib.load().nervecl_bn_relu_fwd(xp, ldx, _flat(stat, "stat"), _flat(gamma, "gamma"),
                                              _flat(beta, "beta"), rp, ldr, yp, ldy, _dt(x), c, npix, groups,
                                              _stream()), "bn_relu_fwd")


@_op("bn_relu_bwd_reduce(Tensor x, Tensor dy, Tensor stat, Tensor gamma, Tensor beta, int groups, "
     "Tensor(a!) bsums) -> ()")
def _bn_relu_bwd_reduce(x, dy, stat, gamma, beta, groups, bsums) -> None:
    xp, ldx, c, npix = _grouped(x, groups, "x")
    dp, ldd, *_ = _nhwc(dy, "dy")
    _lib.check(_lib.load().nervecl_bn_relu_bwd_reduce(xp, ldx, dp, ldd, _flat(stat, "stat"), _flat(gamma, "gamma"),
                                                     _flat(beta, "beta"), _dt(x), c, npix, groups,
                                                     _flat(bsums, "bsums", torch.float64), _stream()),
               "bn_relu_bwd_reduce")


@_op("bn_relu_bwd_apply(Tensor x, Tensor dy, Tensor stat, Tensor gamma, Tensor beta, Tensor bsums, "
     "Tensor(a!) dx, Tensor(b!)? dgamma, Tensor(c!)? dbeta, int groups, bool training) -> ()")
def _bn_relu_bwd_apply(x, dy, stat, gamma, beta, bsums, dx, dgamma, dbeta, groups, training) -> None:
    xp, ldx, c, npix = _grouped(x, groups, "x")
    dp, ldd, *_ = _nhwc(dy, "dy")
    op, ldo, *_ = _nhwc(dx, "dx")
    _lib.check(_lib.load().nervecl_bn_relu_bwd_apply(
        xp, ldx, dp, ldd, _flat(stat, "stat"), _flat(gamma, "gamma"), _flat(beta, "beta"),
        _flat(bsums, "bsums", torch.float64), op, ldo, _ptr(dgamma), _ptr(dbeta), _dt(x), c, npix, groups,
        int(training), _stream()), "bn_relu_bwd_apply")


# --------------------------------------------------------------------------------------------
# motion
# --------------------------------------------------------------------------------------------
@_op("corr_fwd(Tensor x1, Tensor x2, Tensor(a!) out) -> ()")
def _corr_fwd(x1, x2, out) -> None:
    p1, ld1, n, h, w, c = _nhwc(x1, "x1")
    p2, ld2, *_ = _nhwc(x2, "x2")
    po, ldo, _, _, _, cp = _nhwc(out, "out")
    _lib.check(_lib.load().nervecl_corr_fwd(p1, ld1, p2, ld2, po, ldo, _dt(x1), n, h, w, c, cp, _stream()),
               "corr_fwd")


@_op("corr_bwd(Tensor x1, Tensor x2, Tensor dout, Tensor(a!) dx1, bool acc1, Tensor(b!) dx2, bool acc2, "
     "Tensor(c!)? workspace=None) -> ()")
def _corr_bwd(x1, x2, dout, dx1, acc1, dx2, acc2, workspace=None) -> None:
    """``workspace``: optional [N,H,W,96] buffer of the activation dtype (enables the tcgen05 gradient kernels)."""
    p1, ld1, n, h, w, c = _nhwc(x1, "x1")
    p2, ld2, *_ = _nhwc(x2, "x2")
    pg, ldg, *_ = _nhwc(dout, "dout")
    q1, l1, *_ = _nhwc(dx1, "dx1")
    q2, l2, *_ = _nhwc(dx2, "dx2")
    wp, wb = (None, 0)
    if workspace is not None:
        _cuda(workspace, "workspace")
        if not workspace.is_contiguous() or workspace.dtype != x1.dtype:
            raise RuntimeError("nervecl.corr_bwd: workspace must be contiguous and of the activation dtype")
        wp, wb = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    _lib.check(_lib.load().nervecl_corr_bwd(p1, ld1, p2, ld2, pg, ldg, q1, l1, int(acc1), q2, l2, int(acc2),
                                           _dt(x1), n, h, w, c, wp, wb, _stream()), "corr_bwd")


@_op("warp_fwd(Tensor feat, Tensor flow, Tensor(a!) out, int div_mode, Tensor(b!)? idx_out) -> ()")
def _warp_fwd(feat, flow, out, div_mode, idx_out) -> None:
    pf, ldf, n, h, w, c = _nhwc(feat, "feat")
    po, ldo, *_ = _nhwc(out, "out")
    if tuple(flow.shape) != (n, h, w, 2):
        raise RuntimeError("nervecl.warp_fwd: flow must be [N,H,W,2] float32")
    ip = _flat(idx_out, "idx_out", torch.int32) if idx_out is not None else None
    _lib.check(_lib.load().nervecl_warp_fwd(pf, ldf, _flat(flow, "flow"), po, ldo, _dt(feat), n, h, w, c, div_mode,
                                           ip, _stream()), "warp_fwd")


@_op("warp_bwd(Tensor feat, Tensor flow, Tensor dout, Tensor(a!) dfeat, Tensor(b!) dflow, int div_mode) -> ()")
def _warp_bwd(feat, flow, dout, dfeat, dflow, div_mode) -> None:
    pf, ldf, n, h, w, c = _nhwc(feat, "feat")
    pg, ldg, *_ = _nhwc(dout, "dout")
    pd, ldd, *_ = _nhwc(dfeat, "dfeat")
    if dfeat.dtype != torch.float32:
        raise RuntimeError("nervecl.warp_bwd: dfeat must be float32 (atomic accumulation)")
    _lib.check(_lib.load().nervecl_warp_bwd(pf, ldf, _flat(flow, "flow"), pg, ldg, pd, ldd, _flat(dflow, "dflow"),
                                           _dt(feat), n, h, w, c, div_mode, _stream()), "warp_bwd")


@_op("warp_bwd_lp(Tensor feat, Tensor flow, Tensor dout, Tensor(a!) dfeat, Tensor(b!) dflow, int div_mode) -> ()")
def _warp_bwd_lp(feat, flow, dout, dfeat, dflow, div_mode) -> None:
    """``warp_bwd`` accumulating into ``dfeat`` of the activation dtype (packed bf16 reductions)."""
    pf, ldf, n, h, w, c = _nhwc(feat, "feat")
    pg, ldg, *_ = _nhwc(dout, "dout")
    pd, ldd, *_ = _nhwc(dfeat, "dfeat")
    if dfeat.dtype != feat.dtype:
        raise RuntimeError("nervecl.warp_bwd_lp: dfeat must have the activation dtype")
    _lib.check(_lib.load().nervecl_warp_bwd_lp(pf, ldf, _flat(flow, "flow"), pg, ldg, pd, ldd, _flat(dflow, "dflow"),
                                              _dt(feat), n, h, w, c, div_mode, _stream()), "warp_bwd_lp")


# --------------------------------------------------------------------------------------------
# temporal fusion + CBAM
# --------------------------------------------------------------------------------------------
@_op("tfuse_fwd(Tensor feats, Tensor logits, Tensor(a!) attn, Tensor(b!) out) -> ()")
def _tfuse_fwd(feats, logits, attn, out) -> None:
    pf, ldf, n, h, w, tc = _nhwc(feats, "feats")
    po, ldo, _, _, _, c = _nhwc(out, "out")
    t = tc // c
    _lib.check(_lib.load().nervecl_tfuse_fwd(pf, ldf, _flat(logits, "logits"), _flat(attn, "attn"), po, ldo,
                                            _dt(feats), n * h * w, t, c, _stream()), "tfuse_fwd")


@_op("tfuse_bwd(Tensor feats, Tensor attn, Tensor dout, Tensor? nc_bias, Tensor(a!) dfeats, "
     "Tensor(b!) dlogits) -> ()")
def _tfuse_bwd(feats, attn, dout, nc_bias, dfeats, dlogits) -> None:
    pf, ldf, n, h, w, tc = _nhwc(feats, "feats")
    pg, ldg, _, _, _, c = _nhwc(dout, "dout")
    pd, ldd, *_ = _nhwc(dfeats, "dfeats")
    t = tc // c
    _lib.check(_lib.load().nervecl_tfuse_bwd(pf, ldf, _flat(attn, "attn"), pg, ldg,
                                            _flat(nc_bias, "nc_bias") if nc_bias is not None else None, h * w, pd,
                                            ldd, _flat(dlogits, "dlogits"), _dt(feats), n * h * w, t, c, _stream()),
               "tfuse_bwd")


@_op("chan_sum(Tensor x, float scale, Tensor(a!) out) -> ()")
def _chan_sum(x, scale, out) -> None:
    xp, ldx, n, h, w, c = _nhwc(x, "x")
    _lib.check(_lib.load().nervecl_chan_sum(xp, ldx, _dt(x), n, h * w, c, scale, _flat(out, "out"), _stream()),
               "chan_sum")


@_op("ca_gate_fwd(Tensor pool, Tensor w1, Tensor w2, Tensor(a!) hidden, Tensor(b!) gate) -> ()")
def _ca_gate_fwd(pool, w1, w2, hidden, gate) -> None:
    n, c = pool.shape
    r = w1.shape[0]
    _lib.check(_lib.load().nervecl_ca_gate_fwd(_flat(pool, "pool"), _flat(w1, "w1"), _flat(w2, "w2"),
                                              _flat(hidden, "hidden"), _flat(gate, "gate"), n, c, r, _stream()),
               "ca_gate_fwd")


@_op("ca_gate_bwd(Tensor pool, Tensor w1, Tensor w2, Tensor hidden, Tensor gate, Tensor dgate, "
     "Tensor(a!) dpool, Tensor(b!) dw1, Tensor(c!) dw2) -> ()")
def _ca_gate_bwd(pool, w1, w2, hidden, gate, dgate, dpool, dw1, dw2) -> None:
    n, c = pool.shape
    r = w1.shape[0]
    _lib.check(_lib.load().nervecl_ca_gate_bwd(_flat(pool, "pool"), _flat(w1, "w1"), _flat(w2, "w2"),
                                              _flat(hidden, "hidden"), _flat(gate, "gate"), _flat(dgate, "dgate"),
                                              _flat(dpool, "dpool"), _flat(dw1, "dw1"), _flat(dw2, "dw2"), n, c, r,
                                              _stream()), "ca_gate_bwd")


@_op("cbam_stats_fwd(Tensor x, Tensor gate, Tensor(a!) stats) -> ()")
def _cbam_stats_fwd(x, gate, stats) -> None:
    xp, ldx, n, h, w, c = _nhwc(x, "x")
    _lib.check(_lib.load().nervecl_cbam_stats_fwd(xp, ldx, _flat(gate, "gate"), _flat(stats, "stats"), _dt(x), n,
                                                 h * w, c, _stream()), "cbam_stats_fwd")


@_op("cbam_apply_fwd(Tensor x, Tensor gate, Tensor stats, Tensor w7, Tensor(a!) sgate, Tensor(b!) out) -> ()")
def _cbam_apply_fwd(x, gate, stats, w7, sgate, out) -> None:
    xp, ldx, n, h, w, c = _nhwc(x, "x")
    op, ldo, *_ = _nhwc(out, "out")
    _lib.check(_lib.load().nervecl_cbam_apply_fwd(xp, ldx, _flat(gate, "gate"), _flat(stats, "stats"),
                                                 _flat(w7, "w7"), _flat(sgate, "sgate"), op, ldo, _dt(x), n, h, w,
                                                 c, _stream()), "cbam_apply_fwd")


@_op("cbam_bwd_dz(Tensor x, Tensor gate, Tensor sgate, Tensor dy, Tensor(a!) dz) -> ()")
def _cbam_bwd_dz(x, gate, sgate, dy, dz) -> None:
    xp, ldx, n, h, w, c = _nhwc(x, "x")
    dp, ldd, *_ = _nhwc(dy, "dy")
    _lib.check(_lib.load().nervecl_cbam_bwd_dz(xp, ldx, _flat(gate, "gate"), _flat(sgate, "sgate"), dp, ldd,
                                              _flat(dz, "dz"), _dt(x), n, h * w, c, _stream()), "cbam_bwd_dz")


@_op("cbam_bwd_spatial(Tensor dz, Tensor stats, Tensor w7, Tensor(a!) dstats, Tensor(b!) dw7) -> ()")
def _cbam_bwd_spatial(dz, stats, w7, dstats, dw7) -> None:
    n, h, w = dz.shape
    _lib.check(_lib.load().nervecl_cbam_bwd_spatial(_flat(dz, "dz"), _flat(stats, "stats"), _flat(w7, "w7"),
                                                   _flat(dstats, "dstats"), _flat(dw7, "dw7"), n, h, w, _stream()),
               "cbam_bwd_spatial")


@_op("cbam_bwd_dx(Tensor x, Tensor gate, Tensor sgate, Tensor stats, Tensor dstats, Tensor dy, Tensor(a!) dx, "
     "Tensor(b!) dgate) -> ()")
def _cbam_bwd_dx(x, gate, sgate, stats, dstats, dy, dx, dgate) -> None:
    xp, ldx, n, h, w, c = _nhwc(x, "x")
    dp, ldd, *_ = _nhwc(dy, "dy")
    op, ldo, *_ = _nhwc(dx, "dx")
    _lib.check(_lib.load().nervecl_cbam_bwd_dx(xp, ldx, _flat(gate, "gate"), _flat(sgate, "sgate"),
                                              _flat(stats, "stats"), _flat(dstats, "dstats"), dp, ldd, op, ldo,
                                              _flat(dgate, "dgate"), _dt(x), n, h * w, c, _stream()), "cbam_bwd_dx")


# --------------------------------------------------------------------------------------------
# output stage
# --------------------------------------------------------------------------------------------
def _lr_view(lr: Tensor):
    _cuda(lr, "lr")
    if lr.dim() != 4 or lr.dtype != torch.float32 or lr.stride(3) != 1:
        raise RuntimeError("nervecl: lr must be (N,C,H,W) float32 with unit column stride")
    return lr.data_ptr(), lr.stride(0), lr.stride(1), lr.stride(2)


@_op("upfinish_fwd(Tensor conv_out, Tensor lr, Tensor(a!) out, int scale) -> ()")
def _upfinish_fwd(conv_out, lr, out, scale) -> None:
    n, c, h, w = lr.shape
    lp, sn, sc, sh = _lr_view(lr)
    _lib.check(_lib.load().nervecl_upfinish_fwd(_flat(conv_out, "conv_out"), lp, sn, sc, sh, _flat(out, "out"), n, c,
                                               h, w, scale, _stream()), "upfinish_fwd")


@_op("upfinish_bwd(Tensor conv_out, Tensor lr, Tensor dout, Tensor(a!) dconv, int scale) -> ()")
def _upfinish_bwd(conv_out, lr, dout, dconv, scale) -> None:
    n, c, h, w = lr.shape
    lp, sn, sc, sh = _lr_view(lr)
    _lib.check(_lib.load().nervecl_upfinish_bwd(_flat(conv_out, "conv_out"), lp, sn, sc, sh, _flat(dout, "dout"),
                                               _flat(dconv, "dconv"), n, c, h, w, scale, _stream()), "upfinish_bwd")


@_op("bicubic_blend(Tensor(a!) out, Tensor lr, int scale, float strength) -> ()")
def _bicubic_blend(out, lr, scale, strength) -> None:
    """out = strength * out + (1 - strength) * bicubic(lr) (``nervecl_bicubic_blend``)."""
    n, c, h, w = lr.shape
    lp, sn, sc, sh = _lr_view(lr)
    if tuple(out.shape) != (n, c, h * scale, w * scale):
        raise RuntimeError("nervecl.bicubic_blend: out must be (N,C,s*H,s*W)")
    _lib.check(_lib.load().nervecl_bicubic_blend(_flat(out, "out"), lp, sn, sc, sh, n, c, h, w, scale, strength, _stream()),
               "bicubic_blend")


# --------------------------------------------------------------------------------------------
# FrameRecoveryNet trunk
# --------------------------------------------------------------------------------------------
@_op("conv2d_direct(Tensor x, Tensor w, Tensor? bias, Tensor(a!) out, int stride, int pad, bool relu) -> ()")
def _conv2d_direct(x, w, bias, out, stride, pad, relu) -> None:
    """Strided direct convolution, OIHW fp32 weights (``nervecl_conv2d_direct``)."""
    xp, ldx, n, h, wd, cin = _nhwc(x, "x")
    op, ldo, on, oh, ow, oc = _nhwc(out, "out")
    o, i, kh, kw = w.shape
    if i != cin or o != oc or kh != kw or on != n:
        raise RuntimeError("nervecl.conv2d_direct: weight / tensor shapes do not match")
    if (oh, ow) != ((h + 2 * pad - kh) // stride + 1, (wd + 2 * pad - kh) // stride + 1):
        raise RuntimeError("nervecl.conv2d_direct: out has the wrong spatial size")
    _lib.check(_lib.load().nervecl_conv2d_direct(xp, ldx, _dt(x), _flat(w, "w"), _flat(bias, "bias") if bias is not None else None,
                                                op, ldo, _dt(out), n, h, wd, cin, o, kh, stride, pad, int(relu), _stream()),
               "conv2d_direct")


@_op("maxpool2d(Tensor x, Tensor(a!) y, int k, int stride, int pad) -> ()")
def _maxpool2d(x, y, k, stride, pad) -> None:
    xp, ldx, n, h, w, c = _nhwc(x, "x")
    yp, ldy, yn, oh, ow, yc = _nhwc(y, "y")
    if (yn, yc) != (n, c) or (oh, ow) != ((h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1) or x.dtype != y.dtype:
        raise RuntimeError("nervecl.maxpool2d: output shape / dtype mismatch")
    _lib.check(_lib.load().nervecl_maxpool2d(xp, ldx, yp, ldy, _dt(x), n, h, w, c, k, stride, pad, _stream()), "maxpool2d")


@_op("depth_to_space(Tensor x, Tensor(a!) y, int s) -> ()")
def _depth_to_space(x, y, s) -> None:
    xp, ldx, n, h, w, cs = _nhwc(x, "x")
    yp, ldy, yn, oh, ow, c = _nhwc(y, "y")
    if (yn, oh, ow) != (n, h * s, w * s) or cs != c * s * s or x.dtype != y.dtype:
        raise RuntimeError("nervecl.depth_to_space: output shape / dtype mismatch")
    _lib.check(_lib.load().nervecl_depth_to_space(xp, ldx, yp, ldy, _dt(x), n, h, w, c, s, _stream()), "depth_to_space")


@_op("resize_bilinear(Tensor x, Tensor(a!) y) -> ()")
def _resize_bilinear(x, y) -> None:
    xp, ldx, n, h, w, c = _nhwc(x, "x")
    yp, ldy, yn, oh, ow, yc = _nhwc(y, "y")
    if (yn, yc) != (n, c) or x.dtype != y.dtype:
        raise RuntimeError("nervecl.resize_bilinear: output shape / dtype mismatch")
    _lib.check(_lib.load().nervecl_resize_bilinear(xp, ldx, yp, ldy, _dt(x), n, h, w, c, oh, ow, _stream()), "resize_bilinear")


@_op("fusion_blend(Tensor aligned, Tensor logits, Tensor spatial, Tensor temporal, Tensor(a!) out) -> ()")
def _fusion_blend(aligned, logits, spatial, temporal, out) -> None:
    ap, lda, n, h, w, c = _nhwc(aligned, "aligned")
    lp, ldl, *_ = _nhwc(logits, "logits")
    sp, lds, _, _, _, cs = _nhwc(spatial, "spatial")
    tp, ldt, _, _, _, ct = _nhwc(temporal, "temporal")
    op, ldo, *_ = _nhwc(out, "out")
    if logits.dtype != torch.float32 or logits.shape[-1] < 2:
        raise RuntimeError("nervecl.fusion_blend: logits must be float32 [N,H,W,>=2]")
    _lib.check(_lib.load().nervecl_fusion_blend(ap, lda, lp, ldl, sp, lds, cs, tp, ldt, ct, op, ldo, _dt(aligned), c, n * h * w,
                                               _stream()), "fusion_blend")


@_op("recovery_finish(Tensor conv_out, Tensor frame, Tensor? mask, Tensor(a!) out) -> ()")
def _recovery_finish(conv_out, frame, mask, out) -> None:
    cp, ldc, n, hd, wd, c = _nhwc(conv_out, "conv_out")
    fn, fc, h, w = frame.shape
    if conv_out.dtype != torch.float32 or (fn, fc) != (n, c) or out.shape != frame.shape:
        raise RuntimeError("nervecl.recovery_finish: shape / dtype mismatch")
    _lib.check(_lib.load().nervecl_recovery_finish(cp, ldc, hd, wd, _flat(frame, "frame"),
                                                  _flat(mask, "mask") if mask is not None else None, _flat(out, "out"), n, c,
                                                  h, w, _stream()), "recovery_finish")


# --------------------------------------------------------------------------------------------
# elementwise
# --------------------------------------------------------------------------------------------
@_op("axpy(Tensor x, Tensor(a!) out, float alpha, bool accumulate) -> ()")
def _axpy(x, out, alpha, accumulate) -> None:
    xp, ldx, n, h, w, c = _nhwc(x, "x")
    op, ldo, *_ = _nhwc(out, "out")
    _lib.check(_lib.load().nervecl_axpy(xp, ldx, _dt(x), op, ldo, _dt(out), n * h * w, c, alpha, int(accumulate),
                                       _stream()), "axpy")


@_op("relu_bwd(Tensor dy, Tensor y, Tensor? y_sub, Tensor(a!) out) -> ()")
def _relu_bwd(dy, y, y_sub, out) -> None:
    dp, ldd, n, h, w, c = _nhwc(dy, "dy")
    yp, ldy, *_ = _nhwc(y, "y")
    op, ldo, *_ = _nhwc(out, "out")
    sp, lds = (None, 0)
    if y_sub is not None:
        sp, lds, *_ = _nhwc(y_sub, "y_sub")
    _lib.check(_lib.load().nervecl_relu_bwd(dp, ldd, yp, ldy, sp, lds, op, ldo, _dt(dy), n * h * w, c, _stream()),
               "relu_bwd")


@_op("fill_zero(Tensor(a!) t) -> ()")
def _fill_zero(t) -> None:
    _cuda(t, "t")
    if not t.is_contiguous():
        raise RuntimeError("nervecl.fill_zero: tensor must be contiguous")
    _lib.check(_lib.load().nervecl_fill_zero(t.data_ptr(), t.numel() * t.element_size(), _stream()), "fill_zero")


@_op("mse_fwd_bwd(Tensor a, Tensor b, Tensor(a!)? dgrad, Tensor(b!) loss, float scale) -> ()")
def _mse_fwd_bwd(a, b, dgrad, loss, scale) -> None:
    _lib.check(_lib.load().nervecl_mse_fwd_bwd(_flat(a, "a"), _flat(b, "b"),
                                              _flat(dgrad, "dgrad") if dgrad is not None else None,
                                              _flat(loss, "loss"), a.numel(), scale, _stream()), "mse_fwd_bwd")


# --------------------------------------------------------------------------------------------
# EWC + optimiser
# --------------------------------------------------------------------------------------------
def _table(tensors: Sequence[Optional[Tensor]]):
    n = len(tensors)
    ptrs = (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in tensors])
    return ptrs, n


def _numels(tensors: Sequence[Tensor]):
    return (C.c_int64 * len(tensors))(*[t.numel() for t in tensors])


def _check_flat_list(tensors, name):
    for t in tensors:
        if t is None:
            continue
        _cuda(t, name)
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"nervecl: every tensor in `{name}` must be contiguous float32")


@_op("ewc_fisher_accum(Tensor(a!) fisher, Tensor?[] grads, int[] numels, float scale) -> ()")
def _ewc_fisher_accum(fisher, grads, numels, scale) -> None:
    _check_flat_list(grads, "grads")
    ptrs, n = _table(grads)
    _lib.check(_lib.load().nervecl_ewc_fisher_accum(_flat(fisher, "fisher"), ptrs, (C.c_int64 * n)(*numels), n, scale,
                                                   _stream()), "ewc_fisher_accum")


@_op("ewc_axpby(Tensor(a!) v, Tensor? w, float a, float b) -> ()")
def _ewc_axpby(v, w, a, b) -> None:
    _lib.check(_lib.load().nervecl_ewc_axpby(_flat(v, "v"), _flat(w, "w") if w is not None else None, v.numel(), a, b,
                                            _stream()), "ewc_axpby")


@_op("ewc_penalty_fwd(Tensor[] theta, Tensor fisher, Tensor star, float coef, Tensor(a!) out) -> ()")
def _ewc_penalty_fwd(theta, fisher, star, coef, out) -> None:
    _check_flat_list(theta, "theta")
    ptrs, n = _table(theta)
    _lib.check(_lib.load().nervecl_ewc_penalty_fwd(ptrs, _numels(theta), n, _flat(fisher, "fisher"),
                                                  _flat(star, "star"), coef, _flat(out, "out"), _stream()),
               "ewc_penalty_fwd")


@_op("ewc_penalty_bwd(Tensor[] theta, Tensor(a!)[] grads, Tensor fisher, Tensor star, float coef2, "
     "Tensor? gscale) -> ()")
def _ewc_penalty_bwd(theta, grads, fisher, star, coef2, gscale) -> None:
    _check_flat_list(theta, "theta")
    _check_flat_list(grads, "grads")
    tp, n = _table(theta)
    gp, _ = _table(grads)
    _lib.check(_lib.load().nervecl_ewc_penalty_bwd(tp, gp, _numels(theta), n, _flat(fisher, "fisher"),
                                                  _flat(star, "star"), coef2,
                                                  _flat(gscale, "gscale") if gscale is not None else None,
                                                  _stream()), "ewc_penalty_bwd")


@_op("flat_gather(Tensor?[] src, int[] offsets, Tensor(a!) flat) -> ()")
def _flat_gather(src, offsets, flat) -> None:
    """flat[offsets[i] : offsets[i] + src[i].numel()] = src[i] for every non-None entry (``nervecl_flat_gather``)."""
    _check_flat_list(src, "src")
    ptrs, n = _table(src)
    numels = (C.c_int64 * n)(*[0 if t is None else t.numel() for t in src])
    _lib.check(_lib.load().nervecl_flat_gather(ptrs, numels, (C.c_int64 * n)(*offsets), n, _flat(flat, "flat"), _stream()),
               "flat_gather")


@_op("si_update(Tensor[] theta, Tensor?[] grads, Tensor(a!) W, Tensor(b!) p_old) -> ()")
def _si_update(theta, grads, W, p_old) -> None:
    """Synaptic Intelligence running importance (``nervecl_si_update``, reference ewc.py:342-352)."""
    _check_flat_list(theta, "theta")
    _check_flat_list(grads, "grads")
    tp, n = _table(theta)
    gp, _ = _table(grads)
    _lib.check(_lib.load().nervecl_si_update(tp, gp, _numels(theta), n, _flat(W, "W"), _flat(p_old, "p_old"), _stream()),
               "si_update")


@_op("si_register(Tensor[] theta, Tensor(a!) W, Tensor(b!) p_old, Tensor(c!) omega, float damping) -> ()")
def _si_register(theta, W, p_old, omega, damping) -> None:
    """Synaptic Intelligence consolidation (``nervecl_si_register``, reference ewc.py:354-366)."""
    _check_flat_list(theta, "theta")
    tp, n = _table(theta)
    _lib.check(_lib.load().nervecl_si_register(tp, _numels(theta), n, _flat(W, "W"), _flat(p_old, "p_old"),
                                              _flat(omega, "omega"), damping, _stream()), "si_register")


@_op("adamw_step(Tensor(a!) param, Tensor grad, Tensor(b!) exp_avg, Tensor(c!) exp_avg_sq, float lr, float beta1, "
     "float beta2, float eps, float weight_decay, int step, float grad_scale, Tensor? step_dev=None) -> ()")
def _adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                step_dev=None) -> None:
    """``step_dev``: optional device int32 scalar holding the step count (CUDA-graph replays), see ``nervecl_adamw_step``."""
    _lib.check(_lib.load().nervecl_adamw_step(_flat(param, "param"), _flat(grad, "grad"), _flat(exp_avg, "exp_avg"),
                                             _flat(exp_avg_sq, "exp_avg_sq"), param.numel(), lr, beta1, beta2, eps,
                                             weight_decay, step,
                                             grad_scale, _flat(step_dev, "step_dev", torch.int32) if step_dev is not None else None,
                                             _stream()), "adamw_step")


nv = torch.ops.nervecl
OP_NAMES = tuple(_impls)
