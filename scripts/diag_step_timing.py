import sys, time, os
sys.argv = ["bench.py", "--steps", "10", "--warmup", "3"]
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "continual-learning-for-dynamic-video-quality-enhancement_b200"))
import torch, bench
from nerve_cl_b200 import ops
from nerve_cl_b200.models import SuperResolutionNet
from nerve_cl_b200.optim import FlatAdamW
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = SuperResolutionNet(scale_factor=2, num_features=64, num_residual_blocks=8, temporal_window=1).to(dev).train()
model.compute_dtype = torch.bfloat16
opt = FlatAdamW(model, lr=1e-3, weight_decay=1e-5)
lr, hr = bench.synth_batch(16, 3, 360, 640, 2, 1234, dev)
def step():
    opt.zero_grad()
    out = model(lr)
    loss = torch.nn.functional.mse_loss(out, hr)
    loss.backward()
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
for rep in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.LAUNCHES[0]
    t0 = time.perf_counter(); e0.record()
    losses = []
    for _ in range(10):
        losses.append(step())
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"rep {rep}: event {e0.elapsed_time(e1):.1f} ms wall {wall:.1f} ms launches {ops.LAUNCHES[0]-n0} loss {float(losses[-1]):.5f}", flush=True)
    # per-step synchronised timing
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("   per-step synced:", [round(t, 1) for t in ts], flush=True)
