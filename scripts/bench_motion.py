"""Micro-benchmark of the motion kernels at the cfg-2 shape (16 x 360 x 640, 64 channels, bf16): achieved GB/s of the
algorithmic bytes against the measured HBM copy bandwidth."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "continual-learning-for-dynamic-video-quality-enhancement_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from nerve_cl_b200 import ops  # noqa: E402

nv = ops.nv
N, H, W, C = 16, 360, 640, 64
dev = "cuda"
peak = 6546.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def ms(fn, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


bf = torch.bfloat16
x1 = torch.randn(N, H, W, C, device=dev).to(bf)
x2 = torch.randn(N, H, W, C, device=dev).to(bf)
corr = torch.empty(N, H, W, 96, device=dev, dtype=bf)
g = torch.randn(N, H, W, 96, device=dev).to(bf)
d1, d2 = torch.empty_like(x1), torch.empty_like(x2)
flow = (torch.rand(N, H, W, 2, device=dev) - 0.5) * 6
flow_smooth = 0.3 + 0.05 * torch.randn(N, H, W, 2, device=dev)       # what flow_net produces: neighbours share their corners
out = torch.empty_like(x1)
dfeat = torch.zeros_like(x1)
dflow = torch.empty_like(flow)
ws = torch.empty_like(corr)
px = N * H * W
e = 2
rows = {
    "corr_fwd": ((2 * C * e + 96 * e) * px, lambda: nv.corr_fwd(x1, x2, corr)),
    "corr_bwd": ((96 * e + 4 * C * e) * px, lambda: nv.corr_bwd(x1, x2, g, d1, False, d2, False, ws)),
    "corr_bwd_mma": ((96 * e + 4 * C * e) * px, lambda: nv.corr_bwd(x1, x2, g, d1, False, d2, False)),
    "warp_fwd": ((2 * C * e + 8) * px, lambda: nv.warp_fwd(x1, flow, out, 0, None)),
    "warp_bwd_lp": ((3 * C * e + 16) * px, lambda: nv.warp_bwd_lp(x1, flow, g[..., :C], dfeat, dflow, 0)),
    "warp_fwd_smooth": ((2 * C * e + 8) * px, lambda: nv.warp_fwd(x1, flow_smooth, out, 0, None)),
    "warp_bwd_lp_smooth": ((3 * C * e + 16) * px, lambda: nv.warp_bwd_lp(x1, flow_smooth, g[..., :C], dfeat, dflow, 0)),
}
only = sys.argv[1:] or list(rows)
for k in only:
    nbytes, fn = rows[k]
    t = ms(fn)
    print(f"{k:14s} {t:7.3f} ms  {nbytes / t / 1e6:8.1f} GB/s  {nbytes / t / 1e6 / peak:5.3f} of {peak:.0f}")
