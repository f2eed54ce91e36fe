#!/usr/bin/env python
"""Clip-sharded video enhancement on the B200 path (BASELINE.json configs[2] / configs[4]):
clip k runs on rank k mod world, no collective on the data path; within a rank the sliding windows of
`EnhancementEngine.enhance_video` are batched `--batch-size` windows per network call.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 experiments/infer_video.py --clips 64 --frames 16 \
        --scale 4 --sr-window 2 --height 180 --width 320                      # cfg 3: x4, T=5, 320x180 -> 1280x720
    python experiments/infer_video.py --recovery --height 540 --width 960     # cfg 5: inpainting -> x2 SR, 540p -> 1080p

Clips are synthetic here (seeded per clip id); `--checkpoint` loads an engine state_dict (the reference's
`checkpoints/continual_model.pt` format).  Every clip goes host (pinned) -> device -> `enhance_video` -> host (pinned);
the reported rate is output frames per second over the whole job, device-timed, max over ranks.
"""
import argparse
import json

import _common  # noqa: F401
import torch


def synth_clip(clip_id: int, frames: int, h: int, w: int, recovery: bool):
    g = torch.Generator().manual_seed(10_000 + clip_id)
    base = torch.rand(3, h + 16, w + 16, generator=g)
    vid = torch.stack([base[:, 8 + (t % 5) - 2: 8 + (t % 5) - 2 + h, 8 + (2 * t) % 7 - 3: 8 + (2 * t) % 7 - 3 + w]
                       for t in range(frames)])
    vid = (vid + 0.01 * torch.randn(vid.shape, generator=g)).clamp_(0, 1)
    masks = None
    if recovery:
        masks = torch.zeros(frames, 1, h, w)
        for t in range(1, frames, 3):                          # every third frame has a corrupted rectangle
            y0, x0 = (37 * t) % (h // 2), (53 * t) % (w // 2)
            masks[t, :, y0:y0 + h // 3, x0:x0 + w // 3] = 1
    return vid, masks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=8)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--scale", type=int, default=2)
    ap.add_argument("--sr-window", type=int, default=1)
    ap.add_argument("--recovery", action="store_true", help="full pipeline: FrameRecoveryNet inpainting -> SR")
    ap.add_argument("--lightweight", action="store_true")
    ap.add_argument("--batch-size", type=int, default=16, help="windows per network call")
    ap.add_argument("--dtype", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--save", default=None, help="directory for the enhanced clips (clip_<id>.pt)")
    args = ap.parse_args()

    from nerve_cl_b200 import distributed as nd
    from nerve_cl_b200.models import EnhancementConfig, EnhancementEngine
    rank, local_rank, world, device = _common.setup_distributed()
    torch.manual_seed(0)
    engine = EnhancementEngine(EnhancementConfig(frame_recovery_enabled=args.recovery, scale_factor=args.scale,
                                                 sr_temporal_window=args.sr_window, use_lightweight_sr=args.lightweight))
    if args.checkpoint:
        engine.load_state_dict(torch.load(args.checkpoint, map_location="cpu"))
    engine = engine.to(device).eval()
    adt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    for m in (engine.super_resolution, engine.frame_recovery):
        if m is not None:
            m.compute_dtype = adt

    mine = nd.shard_clips(args.clips, rank, world)
    clips = {k: synth_clip(k, args.frames, args.height, args.width, args.recovery) for k in mine}
    pinned = {k: (v.pin_memory(), None if m is None else m.pin_memory()) for k, (v, m) in clips.items()}
    s = args.scale
    out_hosts = [torch.empty((args.frames, 3, args.height * s, args.width * s)).pin_memory() for _ in range(2)]
    d2h = torch.cuda.Stream(device=device)
    state = {"k": 0}

    def run_clip(k):
        """clip k is enhanced on the compute stream while clip k-1's frames travel back on a side stream"""
        v, m = pinned[k]
        vd = v.to(device, non_blocking=True)
        md = None if m is None else m.to(device, non_blocking=True)
        out = engine.enhance_video(vd, md, batch_size=args.batch_size)
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(d2h):
            d2h.wait_event(done)
            out_hosts[state["k"] & 1].copy_(out, non_blocking=True)
            out.record_stream(d2h)
        state["k"] += 1
        return out_hosts[(state["k"] - 1) & 1]

    if mine:
        run_clip(mine[0])                                      # warm-up: shape plans, packed weights
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in mine:
        host = run_clip(k)
        if args.save:
            torch.cuda.synchronize()
            from pathlib import Path
            Path(args.save).mkdir(parents=True, exist_ok=True)
            torch.save(host.clone(), f"{args.save}/clip_{k}.pt")
    torch.cuda.current_stream().wait_stream(d2h)           # the last clip's frames are back before the clock stops
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        total = args.clips * args.frames
        print(json.dumps({"metric": "enhanced_frames_per_sec", "value": total / (float(ms) / 1e3), "clips": args.clips,
                          "frames_per_clip": args.frames, "n_gpus": world, "ms": float(ms),
                          "config": f"{'recovery+' if args.recovery else ''}SR x{s}, {args.width}x{args.height}, "
                                    f"sr_window {args.sr_window}, {args.dtype}, batch {args.batch_size}"}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
