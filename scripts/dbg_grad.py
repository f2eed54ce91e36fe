import sys, os, torch
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/continual-learning-for-dynamic-video-quality-enhancement_b200")
from nerve_cl_b200.models import SuperResolutionNet
from nerve_cl_b200 import ops
scale, feats, blocks, tw, b, h, w = 2, 64, 2, 1, 2, 40, 160
torch.manual_seed(11)
model = SuperResolutionNet(scale_factor=scale, num_features=feats, num_residual_blocks=blocks, temporal_window=tw).cuda().train()
sd = {k: v.clone() for k, v in model.state_dict().items()}
t = model.num_frames
g = torch.Generator().manual_seed(12)
base = torch.rand(b, 3, h, w, generator=g)
x = torch.stack([torch.roll(base, (i - t // 2, 2 * (i - t // 2)), (2, 3)) for i in range(t)], 1).cuda()
tgt = torch.rand(b, 3, h * scale, w * scale, generator=g).cuda()
res = {}
for tag, dt, eng in (("fp32", torch.float32, ops.CONV_AUTO), ("bf16", torch.bfloat16, ops.CONV_AUTO), ("bf16_simt", torch.bfloat16, ops.CONV_SIMT)):
    model.load_state_dict(sd); model.zero_grad(); model.compute_dtype = dt; model.conv_engine = eng
    model._plans.clear()
    out = model(x); torch.nn.functional.mse_loss(out, tgt).backward()
    res[tag] = {n: p.grad.detach().double().clone() for n, p in model.named_parameters()}
for n, ref in res["fp32"].items():
    d = float(ref.norm())
    if d < 1e-10: continue
    e1 = float((res["bf16"][n] - ref).norm()) / d
    e2 = float((res["bf16_simt"][n] - ref).norm()) / d
    e3 = float((res["bf16"][n] - res["bf16_simt"][n]).norm()) / float(res["bf16_simt"][n].norm())
    if e3 > 1e-2: print(f"{n:60s} tc-fp32 {e1:.3e}  simt-fp32 {e2:.3e}  tc-simt {e3:.3e}")
print("done")
