"""Worker of tests/test_dp_gpu.py: one data-parallel rank (launched by torch.distributed.run, backend from
NERVECL_DP_BACKEND: gloo puts both ranks on cuda:0 of a 1-GPU box, nccl uses one GPU per rank).  Writes its findings
as JSON to $NERVECL_DP_OUT.<rank>."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "continual-learning-for-dynamic-video-quality-enhancement_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from nerve_cl_b200 import distributed as nd
    from nerve_cl_b200.continual import EWC
    from nerve_cl_b200.models import SuperResolutionNet
    from nerve_cl_b200.optim import FlatAdamW
    backend = os.environ.get("NERVECL_DP_BACKEND", "gloo")
    rank, local_rank, world = nd.init_from_env(backend)
    dev = torch.device("cuda", local_rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    res = {"rank": rank, "world": world}

    def batch(seed, b=2):
        g = torch.Generator().manual_seed(seed)
        return torch.rand(b, 3, 3, 24, 72, generator=g).to(dev), torch.rand(b, 3, 48, 144, generator=g).to(dev)

    torch.manual_seed(100 + rank)            # DIFFERENT initial weights per rank: data_parallel must broadcast rank 0's
    model = SuperResolutionNet(num_features=16, num_residual_blocks=1).to(dev).train()
    model.compute_dtype = torch.float32
    nd.data_parallel(model, bucket_bytes=64 << 10)
    res["divergence_after_broadcast"] = nd.param_divergence(model)

    # (1) the synchronised gradient == the mean of the ranks' local gradients
    x, t = batch(10 + rank)
    sync = model._grad_sync
    model._grad_sync = None
    model.zero_grad()
    torch.nn.functional.mse_loss(model(x), t).backward()
    local = torch.cat([p.grad.flatten() for p in model.parameters()]).clone()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean_local = torch.stack(gathered).mean(0)
    model._grad_sync = sync
    model.zero_grad()
    torch.nn.functional.mse_loss(model(x), t).backward()
    synced = torch.cat([p.grad.flatten() for p in model.parameters()])
    res["sync_vs_mean_local"] = float((synced - mean_local).abs().max() / mean_local.abs().max())
    res["buckets"] = len(sync.launch_log)

    # (2) K optimiser steps on different data per rank: parameters stay identical on every rank
    opt = FlatAdamW(model, lr=1e-3, weight_decay=1e-5)
    for step in range(4):
        x, t = batch(1000 + 10 * step + rank)
        opt.zero_grad()
        torch.nn.functional.mse_loss(model(x), t).backward()
        opt.step()
    res["divergence_after_steps"] = nd.param_divergence(model)
    m = model.feature_extractor.body[0].bn.running_mean
    ref = m.clone()
    dist.broadcast(ref, 0)
    res["bn_drift_before_sync"] = float((m - ref).abs().max())
    nd.sync_buffers(model)
    ref = m.clone()
    dist.broadcast(ref, 0)
    res["bn_drift_after_sync"] = float((m - ref).abs().max())

    # (3) EWC Fisher over ranks == the single-process Fisher over the union of the ranks' batches
    ewc = EWC(model, ewc_lambda=10.0)
    ewc.process_group = dist.group.WORLD
    mine = [batch(500 + 2 * rank + i) for i in range(2)]
    fisher = ewc.compute_fisher(mine)
    got = torch.cat([fisher[n].flatten() for n, _ in model.named_parameters()]).clone()
    res["grad_sync_restored"] = model._grad_sync is sync
    single = EWC(model, ewc_lambda=10.0)
    model._grad_sync = None
    union = [batch(500 + 2 * r + i) for r in range(world) for i in range(2)]
    f1 = single.compute_fisher(union)
    want = torch.cat([f1[n].flatten() for n, _ in model.named_parameters()])
    model._grad_sync = sync
    res["fisher_vs_single_process"] = float((got - want).abs().max() / want.abs().max())
    with open(os.environ["NERVECL_DP_OUT"] + f".{rank}", "w") as f:
        json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
