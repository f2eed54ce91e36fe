#!/usr/bin/env python
"""Micro-benchmark of the bandwidth-bound feature-extractor / motion kernels at the cfg-2 shapes
(T*B = 48 frames of 360x640, C = 64, bf16): CUDA-event time and achieved GB/s of algorithmic bytes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "continual-learning-for-dynamic-video-quality-enhancement_b200"))
from nerve_cl_b200 import ops  # noqa: E402

nv = ops.nv
B, T, H, W, C = int(os.environ.get("B", 16)), 3, 360, 640, 64
dev = "cuda"
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6546.9)
except Exception:
    PEAK = 6546.9


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


def report(name, nbytes, fn):
    ms = bench(fn)
    gbs = nbytes / ms / 1e6
    print(f"{name:22s} {ms:7.3f} ms  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of {PEAK:.0f}", flush=True)


def main():
    which = sys.argv[1:] or ["fe", "motion"]
    N = B * T
    e = 2
    tensor = N * H * W * C * e
    if "fe" in which:
        x = torch.randn((N, H, W, C), device=dev, dtype=torch.bfloat16)
        y = torch.empty_like(x)
        d = torch.randn_like(x)
        wdw = torch.randn((C, 1, 3, 3), device=dev)
        dwg = torch.zeros_like(wdw)
        sums = torch.zeros((T, C, 2), device=dev, dtype=torch.float64)
        stat = torch.zeros((T, C, 2), device=dev)
        stat[..., 1] = 1.0
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        report("dwconv3x3_fwd", 2 * tensor, lambda: nv.dwconv3x3_fwd(x, wdw, y, False, False))
        report("dwconv3x3_fwd(flip,acc)", 3 * tensor, lambda: nv.dwconv3x3_fwd(x, wdw, y, True, True))
        report("dwconv3x3_wgrad", 2 * tensor, lambda: nv.dwconv3x3_wgrad(x, d, dwg))
        report("bn_stats", tensor, lambda: nv.bn_stats(x, T, sums))
        report("bn_relu_fwd", 2 * tensor, lambda: nv.bn_relu_fwd(x, stat, gamma, beta, None, y, T))
        report("bn_relu_fwd(+res)", 3 * tensor, lambda: nv.bn_relu_fwd(x, stat, gamma, beta, d, y, T))
        report("bn_relu_bwd_reduce", 2 * tensor, lambda: nv.bn_relu_bwd_reduce(x, d, stat, gamma, beta, T, sums))
        report("bn_relu_bwd_apply", 3 * tensor,
               lambda: nv.bn_relu_bwd_apply(x, d, stat, gamma, beta, sums, y, dg, db, T, True))
        report("relu_bwd", 3 * tensor, lambda: nv.relu_bwd(d, x, None, y))
        report("axpy", 2 * tensor, lambda: nv.axpy(x, y, 1.0, False))
        del x, y, d
    if "motion" in which:
        px = B * H * W
        f1 = torch.randn((B, H, W, C), device=dev, dtype=torch.bfloat16)
        f2 = torch.randn_like(f1)
        corr = torch.empty((B, H, W, 96), device=dev, dtype=torch.bfloat16)
        dcorr = torch.randn_like(corr)
        d1, d2 = torch.zeros_like(f1), torch.zeros_like(f1)
        flow = torch.randn((B, H, W, 2), device=dev) * 3
        dflow = torch.zeros_like(flow)
        out = torch.empty_like(f1)
        report("corr_fwd", (2 * C + 96) * e * px, lambda: nv.corr_fwd(f1, f2, corr))
        report("corr_bwd", (96 + 4 * C) * e * px, lambda: nv.corr_bwd(f1, f2, dcorr, d1, False, d2, False))
        report("corr_bwd(acc)", (96 + 6 * C) * e * px, lambda: nv.corr_bwd(f1, f2, dcorr, d1, True, d2, True))
        report("warp_fwd", (2 * C * e + 8) * px, lambda: nv.warp_fwd(f1, flow, out, 0, None))
        d32 = torch.zeros((B, H, W, C), device=dev)
        report("warp_bwd", (3 * C * e + 16) * px, lambda: nv.warp_bwd(f1, flow, f2, d32, dflow, 0))


if __name__ == "__main__":
    main()
