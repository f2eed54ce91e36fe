"""Ad-hoc GPU debugging helper (not a test): per-tensor error report for one golden case."""
import sys
import torch
from conftest import build_case, load_golden, relerr

name = sys.argv[1] if len(sys.argv) > 1 else "sr_tiny_x2_train.npz"
dtype = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
g = load_golden(name)
model, scale, training = build_case(g["meta"], "cuda")
model.compute_dtype = dtype
x = torch.from_numpy(g["lr_frames"]).cuda()
target = torch.from_numpy(g["target"]).cuda()
out, inter = model(x, return_intermediate=True)
t = model.num_frames
print("feat_centre", relerr(inter["features"][t // 2], torch.from_numpy(g["feat_centre"])))
print("aligned0   ", relerr(inter["aligned"][0], torch.from_numpy(g["aligned0"])))
print("aggregated ", relerr(inter["aggregated"], torch.from_numpy(g["aggregated"])))
print("out        ", relerr(out, torch.from_numpy(g["out"])))
loss = torch.nn.functional.mse_loss(out, target)
print("loss", float(loss), float(g["loss"]))
loss.backward()
torch.cuda.synchronize()
if "g/gff.0.weight" in g:
    for n, p in model.named_parameters():
        print(f"{relerr(p.grad, torch.from_numpy(g['g/' + n])):10.3e}  {n}")
