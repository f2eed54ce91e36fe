#!/usr/bin/env python
"""Micro-benchmark of the dense-conv kernels at the cfg-2 shapes (B=16, 360x640): CUDA-event time and
algorithmic TFLOP/s per layer shape, row-streaming/auto tcgen05 engine vs the per-tap kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "continual-learning-for-dynamic-video-quality-enhancement_b200"))
from nerve_cl_b200 import ops  # noqa: E402

nv = ops.nv
B, H, W = int(os.environ.get("B", 16)), 360, 640
dev = "cuda"


def bench(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def fwd_case(cin, cout, k, ld_in, ld_out, engines, accumulate=False, mask=False, c0=0):
    x = torch.randn((B, H, W, ld_in), device=dev, dtype=torch.bfloat16)
    out = torch.zeros((B, H, W, ld_out), device=dev, dtype=torch.bfloat16)
    w = torch.randn((k * k, cout, (cin + 7) // 8 * 8), device=dev, dtype=torch.bfloat16) * 0.05
    bias = torch.randn(cout, device=dev)
    m = torch.randn((B, H, W, ld_out), device=dev, dtype=torch.bfloat16) if mask else None
    flops = 2.0 * B * H * W * cin * cout * k * k
    res = []
    for eng in engines:
        try:
            plain = not accumulate and not mask
            ms = bench(lambda: nv.conv2d_fwd(x[..., c0:c0 + cin], w, bias if plain else None, None,
                                             m[..., :cout] if mask else None, None, out[..., :cout], cout,
                                             plain, accumulate, 0, 0, 1.0, eng))
            res.append(f"{ms:7.3f} ms {flops / ms / 1e9:7.1f} TF")
        except RuntimeError as e:
            res.append(f"unsupported ({str(e)[-30:]})")
    print(f"fwd {cin:3d}->{cout:3d} k{k} acc={int(accumulate)} c0={c0}: " + " | ".join(res), flush=True)


def wgrad_case(cin, cout, k, ld_in, ld_dy):
    x = torch.randn((B, H, W, ld_in), device=dev, dtype=torch.bfloat16)
    dy = torch.randn((B, H, W, ld_dy), device=dev, dtype=torch.bfloat16)
    dw = torch.zeros((cout, cin, k, k), device=dev)
    db = torch.zeros(cout, device=dev)
    flops = 2.0 * B * H * W * cin * cout * k * k
    if k == 3:
        ms = bench(lambda: nv.conv3x3_wgrad_grouped(x[..., :cin], dy[..., :cout], [dw], [db], [0], 1.0))
    else:
        ms = bench(lambda: nv.conv2d_wgrad(x[..., :cin], dy[..., :cout], dw, db, 1.0, ops.CONV_TC))
    print(f"wgrad {cin:3d}->{cout:3d} k{k}: {ms:7.3f} ms {flops / ms / 1e9:7.1f} TF", flush=True)


def grouped_case():
    """The dense block's five 3x3 weight gradients as ONE grouped launch (engine._rdb_backward_fused)."""
    x = torch.randn((B, H, W, 256), device=dev, dtype=torch.bfloat16)
    g = torch.randn((B, H, W, 256), device=dev, dtype=torch.bfloat16)
    dws = [torch.zeros((32, 64 + 32 * i, 3, 3), device=dev) for i in range(5)]
    flops = sum(2.0 * B * H * W * 9 * (64 + 32 * i) * 32 for i in range(5))
    ms = bench(lambda: nv.conv3x3_wgrad_grouped(x[..., :192], g[..., 64:224], dws, [], [32 * i for i in range(5)], 1.0))
    print(f"wgrad grouped dense block: {ms:7.3f} ms {flops / ms / 1e9:7.1f} TF", flush=True)


def slice_case(cin, with_x2, bits):
    """A slice gradient of the fused dense-block backward: 3x3 over `cin` later-layer gradient channels (+ the block
    gradient through the centre tap) -> 32 channels, gated by a ReLU mask (bf16 activation or packed sign bits)."""
    g = torch.randn((B, H, W, 256), device=dev, dtype=torch.bfloat16)
    db = torch.randn((B, H, W, 64), device=dev, dtype=torch.bfloat16)
    act = torch.randn((B, H, W, 256), device=dev, dtype=torch.bfloat16)
    out = torch.zeros((B, H, W, 256), device=dev, dtype=torch.bfloat16)
    cpad = (cin + 63) // 64 * 64
    w = torch.randn((9, 32, cpad + 64), device=dev, dtype=torch.bfloat16) * 0.05
    cs = torch.zeros(32, device=dev)
    sb = (torch.randint(0, 65536, (2, B, H, W), device=dev, dtype=torch.int32) - 32768).to(torch.int16)
    x2 = db if with_x2 else None
    if bits:
        fn = lambda: nv.conv2d_fwd(g[..., 64:64 + cin], w, None, None, None, None, out[..., 192:224], 32, False, False, 0, 0,
                                   1.0, ops.CONV_TC, x2, with_x2, cs, sb, 2)
    else:
        fn = lambda: nv.conv2d_fwd(g[..., 64:64 + cin], w, None, None, act[..., 96:128], None, out[..., 192:224], 32, False,
                                   False, 0, 0, 1.0, ops.CONV_TC, x2, with_x2, cs)
    ms = bench(fn)
    nbytes = (cin + (64 if with_x2 else 0) + 32 + (0 if bits else 32)) * 2.0 * B * H * W
    print(f"slice {cin:3d}{'+64c' if with_x2 else '    '} -> 32 {'bits' if bits else 'mask'}: {ms:7.3f} ms  {nbytes / ms / 1e6:6.0f} GB/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    engines = [ops.CONV_TC, ops.CONV_TC_TAPS]
    if which in ("fwd", "all"):
        print("engines: rows/auto | per-tap")
        for cin in (64, 96, 128, 160, 192):
            fwd_case(cin, 32, 3, 224, 224, engines)
        fwd_case(96, 128, 3, 96, 128, engines)
        fwd_case(128, 64, 3, 128, 64, engines)
        fwd_case(192, 64, 3, 192, 64, engines)
        fwd_case(64, 64, 3, 64, 64, engines)
        fwd_case(224, 64, 1, 224, 64, engines)
    if which == "pair":
        # CTA-pair (cta_group::2) row kernel vs the 1-CTA row kernel on the dense-block / flow / attention shapes
        print("engines: pair (auto) | rows1")
        eng2 = [ops.CONV_TC, ops.CONV_TC_ROWS1]
        for cin in (64, 96, 128, 160, 192):
            fwd_case(cin, 32, 3, 256, 256, eng2)
        for cin in (32, 64, 96, 128):
            fwd_case(cin, 32, 3, 256, 256, eng2, mask=True)
        fwd_case(160, 64, 3, 256, 256, eng2)
        fwd_case(96, 128, 3, 96, 128, eng2)
        fwd_case(128, 64, 3, 128, 64, eng2)
        fwd_case(192, 64, 3, 192, 64, eng2)
        fwd_case(64, 64, 3, 64, 64, eng2)
    if which == "slices":
        for bits in (False, True):
            slice_case(64, False, bits)
            for cin in (32, 64, 96, 128):
                slice_case(cin, True, bits)
    if which == "align":
        # input channel slice starting on / off a 128-byte line of the 256-channel-pitch buffer (slice gradients)
        for cin, c0 in ((64, 128), (64, 160), (64, 96), (128, 64), (128, 96), (128, 32)):
            fwd_case(cin, 32, 3, 256, 256, [ops.CONV_TC], mask=True, c0=c0)
    if which == "layout":
        # same conv, input/output pixel pitch 64/32 (contiguous pixels) vs 256 (channel slices of a wide buffer)
        for cin in (64, 128, 192):
            fwd_case(cin, 32, 3, cin, 32, [ops.CONV_TC])
            fwd_case(cin, 32, 3, cin, 256, [ops.CONV_TC])
            fwd_case(cin, 32, 3, 256, 32, [ops.CONV_TC])
            fwd_case(cin, 32, 3, 256, 256, [ops.CONV_TC])
    if which == "rows":
        # (component timing: build with -DNERVECL_TUNING and set NERVECL_ROWS_DBG)
        fwd_case(64, 32, 3, 256, 256, [ops.CONV_TC])
        fwd_case(192, 32, 3, 256, 256, [ops.CONV_TC])
        fwd_case(64, 64, 1, 64, 64, [ops.CONV_TC])
        fwd_case(32, 128, 3, 256, 256, [ops.CONV_TC], accumulate=False, mask=True)
    if which == "k1":
        # 1x1 shapes of the step: LFF data gradient of the last dense layer (B images), extractor pointwise (T*B),
        # LFF forward
        print("engines: rows/auto | per-tap")
        fwd_case(64, 32, 1, 64, 256, engines, accumulate=False, mask=True)
        fwd_case(224, 64, 1, 256, 64, engines)
        B_save = B
        globals()["B"] = 3 * B_save
        fwd_case(64, 64, 1, 64, 64, engines)
        globals()["B"] = B_save
    if which in ("dgrad", "all"):
        for cout in (64, 96, 128, 160, 192):
            fwd_case(32, cout, 3, 224, 224, engines, accumulate=True, mask=True)
        fwd_case(64, 224, 1, 64, 224, engines)
    if which in ("grouped", "wgrad", "all"):
        grouped_case()
    if which in ("wgrad", "all"):
        for cin in (64, 96, 128, 160, 192):
            wgrad_case(cin, 32, 3, 224, 224)
        wgrad_case(224, 64, 1, 224, 64)
    if which == "wgrad3":
        # 3x3 weight gradients outside the dense blocks (flow_net, attention, gff)
        wgrad_case(96, 128, 3, 96, 128)
        wgrad_case(128, 64, 3, 128, 64)
        wgrad_case(192, 64, 3, 192, 64)
        wgrad_case(64, 64, 3, 64, 64)
        wgrad_case(64, 32, 3, 64, 32)
    if which == "wgrad1":
        # 1x1 weight gradients of the step: extractor pointwise and head conv (48 images), LFF (16 images)
        B_save = B
        globals()["B"] = 3 * B_save
        wgrad_case(64, 64, 1, 64, 64)
        wgrad_case(27, 64, 1, 32, 64)
        globals()["B"] = B_save
        wgrad_case(224, 64, 1, 256, 64)
