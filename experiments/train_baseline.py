#!/usr/bin/env python
"""Baseline super-resolution training on the B200 path -- the reference's `experiments/train_baseline.py`
flow (same model hyper-parameters, AdamW + cosine schedule, MSE loss, T = 3 identical frames per sample,
best-PSNR checkpoint with the reference's keys) with torchrun data parallelism added.

    python experiments/train_baseline.py --synthetic                       # 1 GPU
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 experiments/train_baseline.py --synthetic

Each rank trains on its shard of the dataset with `--batch-size` samples per step; gradients are all-reduced
in buckets while backward is still running (nerve_cl_b200.distributed.GradSync); BatchNorm statistics stay
per rank, as in the single-process reference.  Rank 0 validates and writes `checkpoints/best_model.pt`.
"""
import argparse
import math
import time
from pathlib import Path

import _common  # noqa: F401  (sets sys.path)
import torch


def load_dataset(data_dir: str, synthetic: bool, size: int, n_train: int, n_val: int):
    """{data_dir}/{train,val}/data.pt with keys 'lr', 'hr' (the reference's format), or seeded synthetic data."""
    if not synthetic:
        tr, va = torch.load(f"{data_dir}/train/data.pt"), torch.load(f"{data_dir}/val/data.pt")
        return (tr["lr"], tr["hr"]), (va["lr"], va["hr"])
    g = torch.Generator().manual_seed(1234)
    def make(n):
        hr = torch.rand(n, 3, 2 * size, 2 * size, generator=g)
        lr = torch.nn.functional.avg_pool2d(hr, 2)
        return lr, hr
    return make(n_train), make(n_val)


def psnr(pred: torch.Tensor, target: torch.Tensor) -> float:
    mse = torch.mean((pred - target) ** 2)
    return float("inf") if float(mse) == 0 else float(20 * torch.log10(1.0 / torch.sqrt(mse)))


def train(args):
    from nerve_cl_b200 import distributed as nd
    from nerve_cl_b200.models import SuperResolutionNet
    from nerve_cl_b200.optim import FlatAdamW
    rank, local_rank, world, device = _common.setup_distributed()
    (tr_lr, tr_hr), (va_lr, va_hr) = load_dataset(args.data_dir, args.synthetic, args.size, args.train_samples,
                                                   args.val_samples)
    _common.log(rank, f"ranks {world}; train {len(tr_lr)} / val {len(va_lr)} samples; LR {tuple(tr_lr.shape[1:])}")

    torch.manual_seed(0)
    model = SuperResolutionNet(scale_factor=2, num_features=args.features, num_residual_blocks=args.blocks,
                               temporal_window=1).to(device)
    model.compute_dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    nd.data_parallel(model)
    _common.log(rank, f"parameters: {model.get_num_parameters():,}")
    opt = FlatAdamW(model, lr=args.lr, weight_decay=1e-5)

    mine = list(_common.shard(len(tr_lr), rank, world))
    steps_per_epoch = max(len(mine) // args.batch_size, 1)
    best, t0 = 0.0, time.time()
    Path("checkpoints").mkdir(exist_ok=True)
    for epoch in range(args.epochs):
        opt.param_groups[0]["lr"] = 0.5 * args.lr * (1 + math.cos(math.pi * epoch / args.epochs))   # CosineAnnealingLR
        model.train()
        perm = torch.randperm(len(mine), generator=torch.Generator().manual_seed(epoch)).tolist()
        loss_sum = torch.zeros((), device=device)
        for s in range(steps_per_epoch):
            idx = [mine[i] for i in perm[s * args.batch_size:(s + 1) * args.batch_size]]
            lr = tr_lr[idx].to(device, non_blocking=True)
            hr = tr_hr[idx].to(device, non_blocking=True)
            opt.zero_grad()
            out = model(lr.unsqueeze(1).expand(-1, 3, -1, -1, -1))      # T identical frames, as the reference does
            loss = torch.nn.functional.mse_loss(out, hr)
            loss.backward()
            opt.step()
            loss_sum += loss.detach()                                     # no host sync inside the loop
        train_loss = float(loss_sum) / steps_per_epoch
        nd.sync_buffers(model)       # validation / checkpoint use rank 0's BatchNorm statistics on every rank (DDP semantics)
        if rank == 0:
            model.eval()
            vl, vp, nb = 0.0, 0.0, 0
            with torch.no_grad():
                for s in range(0, len(va_lr), args.batch_size):
                    lr, hr = va_lr[s:s + args.batch_size].to(device), va_hr[s:s + args.batch_size].to(device)
                    out = model(lr.unsqueeze(1).expand(-1, 3, -1, -1, -1))
                    vl += float(torch.nn.functional.mse_loss(out, hr))
                    vp += psnr(out, hr)
                    nb += 1
            vl, vp = vl / max(nb, 1), vp / max(nb, 1)
            print(f"Epoch {epoch + 1:3d}/{args.epochs} | Train Loss: {train_loss:.4f} | Val Loss: {vl:.4f} | "
                  f"Val PSNR: {vp:.2f} dB | Time: {time.time() - t0:.1f}s", flush=True)
            if vp > best:
                best = vp
                torch.save({"epoch": epoch, "model_state_dict": model.state_dict(),
                            "optimizer_state_dict": opt.state_dict(), "psnr": best}, "checkpoints/best_model.pt")
        if world > 1:
            torch.distributed.barrier()
    _common.log(rank, f"Training complete. Best PSNR: {best:.2f} dB, total {time.time() - t0:.1f}s")


def main():
    ap = argparse.ArgumentParser(description="Train the NERVE-CL SR baseline on B200")
    ap.add_argument("--data-dir", default="data")
    ap.add_argument("--batch-size", type=int, default=16, help="samples per step PER GPU")
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--features", type=int, default=32)        # the reference launcher's 32 / 4 configuration
    ap.add_argument("--blocks", type=int, default=4)
    ap.add_argument("--dtype", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--synthetic", action="store_true", help="seeded random data instead of {data-dir}/*/data.pt")
    ap.add_argument("--size", type=int, default=64, help="synthetic LR size")
    ap.add_argument("--train-samples", type=int, default=512)
    ap.add_argument("--val-samples", type=int, default=64)
    train(ap.parse_args())


if __name__ == "__main__":
    main()
