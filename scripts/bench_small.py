#!/usr/bin/env python
"""Micro-benchmark of the small bandwidth-bound kernels of one cfg-2 training step (B=16, 360x640, 64 features):
CUDA-event time and algorithmic GB/s against the measured copy bandwidth (development helper)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "continual-learning-for-dynamic-video-quality-enhancement_b200"))
from nerve_cl_b200 import ops  # noqa: E402

nv = ops.nv
B, T, H, W, F, S = 16, 3, 360, 640, 64, 2
dev = "cuda"
PX = B * H * W
bf = torch.bfloat16


def bench(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, nbytes):
    print(f"{name:24s} {ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s  ({nbytes / ms / 1e6 / 6546.9:.2f} of copy bandwidth)", flush=True)


def act(n, c, dtype=bf):
    return torch.randn((n, H, W, c), device=dev, dtype=dtype)


flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

# output stage
co = act(B, 3 * S * S, torch.float32) * 0.3
lr = torch.rand((B, T, 3, H, W), device=dev)
out = torch.empty((B, 3, H * S, W * S), device=dev)
dy = torch.randn_like(out)
dconv = torch.empty_like(co)
report("upfinish_fwd", bench(lambda: nv.upfinish_fwd(co, lr[:, 1], out, S)), co.numel() * 4 + out.numel() * 4 + PX * 12)
report("upfinish_bwd", bench(lambda: nv.upfinish_bwd(co, lr[:, 1], dy, dconv, S)), co.numel() * 8 + out.numel() * 4 + PX * 12)

# CBAM
x = act(B, F)
gate = torch.rand((B, F), device=dev)
stats = torch.empty((B, H, W, 2), device=dev)
w7 = torch.randn((1, 2, 7, 7), device=dev) * 0.2
sg = torch.empty((B, H, W), device=dev)
o = torch.empty_like(x)
dz = torch.empty((B, H, W), device=dev)
dstats, dw7 = torch.empty_like(stats), torch.zeros_like(w7)
dx, dgate = torch.empty_like(x), torch.zeros((B, F), device=dev)
g = act(B, F)
report("cbam_stats_fwd", bench(lambda: nv.cbam_stats_fwd(x, gate, stats)), PX * (F * 2 + 8))
report("cbam_apply_fwd", bench(lambda: nv.cbam_apply_fwd(x, gate, stats, w7, sg, o)), PX * (F * 4 + 12))
report("cbam_bwd_dz", bench(lambda: nv.cbam_bwd_dz(x, gate, sg, g, dz)), PX * (F * 4 + 8))
report("cbam_bwd_spatial", bench(lambda: nv.cbam_bwd_spatial(dz, stats, w7, dstats, dw7)), PX * 20)
report("cbam_bwd_dx", bench(lambda: nv.cbam_bwd_dx(x, gate, sg, stats, dstats, g, dx, dgate)), PX * (F * 6 + 20))

# temporal fusion
cat = act(B, T * F)
lg = torch.randn((B, H, W, T), device=dev)
at = torch.empty_like(lg)
report("tfuse_fwd", bench(lambda: nv.tfuse_fwd(cat, lg, at, o)), PX * ((T + 1) * F * 2 + 8 * T))
dcat, dlg = torch.empty_like(cat), torch.empty_like(lg)
nb = torch.randn((B, F), device=dev)
report("tfuse_bwd", bench(lambda: nv.tfuse_bwd(cat, at, g, nb, dcat, dlg)), PX * ((2 * T + 1) * F * 2 + 8 * T))

# axpy (the engine's three big ones) and the head's unfold
feat = act(B, F)
report("axpy copy -> cat slice", bench(lambda: nv.axpy(feat, cat[..., F:2 * F], 1.0, False)), PX * F * 4)
report("axpy cat slice -> copy", bench(lambda: nv.axpy(cat[..., F:2 * F], feat, 1.0, False)), PX * F * 4)
report("axpy accumulate", bench(lambda: nv.axpy(g, feat, 1.0, True)), PX * F * 6)
unf = torch.empty((T * B, H, W, 32), device=dev, dtype=bf)
report("pack_frames_unfold3", bench(lambda: nv.pack_frames_unfold3(lr, unf)), T * PX * (12 + 64))
y = act(B, F)
report("relu_bwd", bench(lambda: nv.relu_bwd(g, y, feat, o)), PX * F * 8)
