"""ctypes binding of ``libnervecl.so`` (the C ABI declared in ``include/nervecl.h``).

There is deliberately no fallback: if the shared library is missing or a call fails, a
``RuntimeError`` is raised.  The library is built in-tree by ``__graft_entry__.build()`` /
``make -C csrc`` with ``nvcc -gencode arch=compute_100a,code=sm_100a``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnervecl.so")

F32, BF16 = 0, 1
CONV_AUTO, CONV_SIMT, CONV_TC, CONV_TC_TAPS, CONV_TC_ROWS1 = 0, 1, 2, 3, 4

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class ConvParams(C.Structure):
    """Mirror of ``nervecl_conv_params``."""
    _fields_ = [
        ("N", c_i32), ("H", c_i32), ("W", c_i32), ("Cin", c_i32), ("Cout", c_i32), ("K", c_i32),
        ("w_ld", c_i32), ("w_rows", c_i32), ("dtype", c_i32), ("out_dtype", c_i32),
        ("engine", c_i32), ("relu", c_i32), ("accumulate", c_i32), ("res_channels", c_i32),
        ("mask_c0", c_i32), ("alpha", c_f32),
        ("x", c_vp), ("ldx", c_i64),
        ("w", c_vp),
        ("bias", c_vp),
        ("res", c_vp), ("ldres", c_i64),
        ("mask", c_vp), ("ldmask", c_i64),
        ("mask_sub", c_vp), ("ldmask_sub", c_i64),
        ("out", c_vp), ("ldo", c_i64),
        ("x2", c_vp), ("ldx2", c_i64), ("Cin2", c_i32), ("x2_center", c_i32),
        ("colsum", c_vp),
        ("sign_bits", c_vp), ("sign_mode", c_i32), ("reserved0", c_i32),
    ]


# name -> argtypes (every function returns int unless listed in _RESTYPES)
_SIGNATURES = {
    "nervecl_abi_version": [],
    "nervecl_error_string": [c_i32],
    "nervecl_has_tcgen05": [],
    "nervecl_pack_frames": [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                            c_vp],
    "nervecl_pack_frames_unfold3": [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                            c_vp],
    "nervecl_unfold3_grad": [c_vp, c_i64, c_i32, c_i32, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_nhwc_to_nchw": [c_vp, c_i64, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_nchw_to_nhwc": [c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_pack_conv_weight": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_pack_conv_weights_batched": [c_i32, C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_i32), C.POINTER(c_i32),
                                          C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_i32), c_i32,
                                          c_vp],
    "nervecl_conv2d_fwd": [C.POINTER(ConvParams), c_vp],
    "nervecl_conv2d_wgrad": [c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32,
                             c_i32, c_f32, c_i32, c_vp],
    "nervecl_conv3x3_wgrad_grouped": [c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                                      C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_vp),
                                      C.POINTER(c_vp), c_f32, c_vp],
    "nervecl_dwconv3x3_fwd": [c_vp, c_i64, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_dwconv3x3_fwd_masked": [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32,
                                     c_i32, c_i32, c_vp],
    "nervecl_dwconv3x3_wgrad": [c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_bn_stats": [c_vp, c_i64, c_i32, c_i32, c_i64, c_i32, c_vp, c_vp],
    "nervecl_bn_finalize": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i64, c_i32, c_f32, c_f32, c_i32, c_vp],
    "nervecl_bn_relu_fwd": [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i64, c_i32, c_vp],
    "nervecl_bn_relu_bwd_reduce": [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, c_i32, c_vp, c_vp],
    "nervecl_bn_relu_bwd_apply": [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp,
                                  c_i32, c_i32, c_i64, c_i32, c_i32, c_vp],
    "nervecl_corr_fwd": [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_corr_bwd": [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_i64, c_i32,
                         c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp],
    "nervecl_warp_fwd": [c_vp, c_i64, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp],
    "nervecl_warp_bwd": [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32,
                         c_i32, c_vp],
    "nervecl_warp_bwd_lp": [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32,
                         c_i32, c_vp],
    "nervecl_tfuse_fwd": [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i32, c_i64, c_i32, c_i32, c_vp],
    "nervecl_tfuse_bwd": [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i32, c_i64, c_i32,
                          c_i32, c_vp],
    "nervecl_chan_sum": [c_vp, c_i64, c_i32, c_i32, c_i64, c_i32, c_f32, c_vp, c_vp],
    "nervecl_ca_gate_fwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp],
    "nervecl_ca_gate_bwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp],
    "nervecl_cbam_stats_fwd": [c_vp, c_i64, c_vp, c_vp, c_i32, c_i32, c_i64, c_i32, c_vp],
    "nervecl_cbam_apply_fwd": [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32,
                               c_i32, c_vp],
    "nervecl_cbam_bwd_dz": [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, c_i32, c_i64, c_i32, c_vp],
    "nervecl_cbam_bwd_spatial": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp],
    "nervecl_cbam_bwd_dx": [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i32, c_i32,
                            c_i64, c_i32, c_vp],
    "nervecl_upfinish_fwd": [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_upfinish_bwd": [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_bicubic_blend": [c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp],
    "nervecl_conv2d_direct": [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                              c_i32, c_i32, c_i32, c_vp],
    "nervecl_maxpool2d": [c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_depth_to_space": [c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_resize_bilinear": [c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_fusion_blend": [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_i32,
                             c_i64, c_vp],
    "nervecl_recovery_finish": [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp],
    "nervecl_axpy": [c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_i64, c_i32, c_f32, c_i32, c_vp],
    "nervecl_relu_bwd": [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i32, c_i64, c_i32, c_vp],
    "nervecl_fill_zero": [c_vp, C.c_size_t, c_vp],
    "nervecl_mse_fwd_bwd": [c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_vp],
    "nervecl_ewc_fisher_accum": [c_vp, C.POINTER(c_vp), C.POINTER(c_i64), c_i32, c_f32, c_vp],
    "nervecl_ewc_axpby": [c_vp, c_vp, c_i64, c_f32, c_f32, c_vp],
    "nervecl_ewc_penalty_fwd": [C.POINTER(c_vp), C.POINTER(c_i64), c_i32, c_vp, c_vp, c_f32, c_vp, c_vp],
    "nervecl_ewc_penalty_bwd": [C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_i64), c_i32, c_vp, c_vp, c_f32,
                                c_vp, c_vp],
    "nervecl_flat_gather": [C.POINTER(c_vp), C.POINTER(c_i64), C.POINTER(c_i64), c_i32, c_vp, c_vp],
    "nervecl_si_update": [C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_i64), c_i32, c_vp, c_vp, c_vp],
    "nervecl_si_register": [C.POINTER(c_vp), C.POINTER(c_i64), c_i32, c_vp, c_vp, c_vp, c_f32, c_vp],
    "nervecl_adamw_step": [c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_i32, c_f32, c_vp, c_vp],
}
_RESTYPES = {"nervecl_error_string": C.c_char_p}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    """Load the kernel library (once).  Fails loudly -- there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"nerve_cl_b200: {LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C <pkg>/csrc`). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI mismatch with include/nervecl.h
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_i32)
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().nervecl_error_string(rc).decode()
        raise RuntimeError(f"nervecl: {what} failed with code {rc}: {msg}")
