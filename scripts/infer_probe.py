import os, sys, time, torch
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "continual-learning-for-dynamic-video-quality-enhancement_b200")):
    sys.path.insert(0, p)
from nerve_cl_b200.models import EnhancementConfig, EnhancementEngine
dev = torch.device("cuda", 0)
eng4 = EnhancementEngine(EnhancementConfig(frame_recovery_enabled=False, scale_factor=4, sr_num_features=64,
                                           sr_num_residual_blocks=8, sr_temporal_window=2)).to(dev).eval()
eng4.super_resolution.compute_dtype = torch.bfloat16
clip_host = torch.rand(32, 3, 180, 320).pin_memory()
hr_host = torch.empty(32, 3, 720, 1280).pin_memory()
v = clip_host.to(dev)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
hr = eng4.enhance_video(v, batch_size=16)
print("enhance_video device-resident ms/clip", t(lambda: eng4.enhance_video(v, batch_size=16)))
print("D2H 354MB ms", t(lambda: hr_host.copy_(hr, non_blocking=True)))
print("H2D ms", t(lambda: clip_host.to(dev, non_blocking=True)))
t0 = time.perf_counter()
for _ in range(5): eng4.enhance_video(v, batch_size=16)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host time per enhance_video call ms", (t1 - t0) / 5 * 1e3, "drain", (t2 - t1) * 1e3)
lr4 = torch.rand(16, 5, 3, 180, 320, device=dev)
m4 = eng4.super_resolution
with torch.no_grad():
    print("model fwd 16 windows ms", t(lambda: m4(lr4)))
