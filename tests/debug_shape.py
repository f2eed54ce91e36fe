"""Ad-hoc GPU debugging helper (not a test): CUDA path vs CPU oracle on an arbitrary shape."""
import sys
import torch
from conftest import relerr
from oracle import sr_oracle
from nerve_cl_b200.models import SuperResolutionNet

F, NB, B, H, W = [int(v) for v in sys.argv[1:6]]
dtype = torch.bfloat16 if len(sys.argv) > 6 and sys.argv[6] == "bf16" else torch.float32
torch.manual_seed(0)
model = SuperResolutionNet(scale_factor=2, num_features=F, num_residual_blocks=NB).cuda().train()
sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
x = torch.rand(B, 3, 3, H, W)
tgt = torch.rand(B, 3, 2 * H, 2 * W)
o_out, o_loss, o_grads = sr_oracle.train_step_grads(sd, x, tgt, 2, True)
model.compute_dtype = dtype
import os
model.warp_div_mode = int(os.environ.get('DIVMODE', '0'))
out = model(x.cuda())
torch.nn.functional.mse_loss(out, tgt.cuda()).backward()
print("out", relerr(out, o_out))
for n, p in model.named_parameters():
    e = relerr(p.grad, o_grads[n])
    if e > 1e-4:
        print(f"{e:10.3e}  {n}")
print("done")
# flip diagnosis: are the mismatches confined to single output channels (one ReLU/cell flip)?
for n, p in model.named_parameters():
    ref = o_grads[n]
    d = (p.grad.cpu() - ref).abs()
    if relerr(p.grad, ref) > 1e-4 and d.dim() == 4:
        per_out = d.flatten(1).max(1)[0] / ref.abs().max()
        bad = (per_out > 1e-4).nonzero().flatten().tolist()
        print(n, "bad out-channels:", bad, "of", d.shape[0])
