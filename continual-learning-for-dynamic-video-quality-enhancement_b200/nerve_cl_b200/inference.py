"""Batched clip inference: the sliding-window loop of the reference ``EnhancementEngine.enhance_video``
(`nerve_cl/models/enhancement_engine.py:187-245`, SR-only engine) without its one-window-per-Python-iteration
structure.

The reference walks the clip frame by frame, slices a window, picks the SR frames out of it (`:141-166`) and calls the
network once per frame with two host synchronisations.  Here the window -> source-frame table is computed once on
the host (`sr_window_indices`, pinned to the reference by `tests/golden/enhance_windows.npz`), the windows are
gathered on the device and the network runs on `batch_size` windows per call under ``no_grad`` -- which also
takes the BatchNorm-folded, activation-free inference path of the engine.

Reference behaviours kept on purpose: windows clipped at the clip borders are padded by repeating the LAST frame of
the clipped window, so near the leading edge the network's centre slot does not hold frame ``t``; the optional
``enhancement_strength < 1`` blend uses the bicubic upsample of frame ``t`` itself.
"""
from __future__ import annotations

from typing import List, Optional

import torch

Tensor = torch.Tensor


def sr_window_indices(num_frames: int, sr_temporal_window: int = 1, recovery_temporal_window: int = 2) -> Tensor:
    """int64 [num_frames, 2*sr_temporal_window+1]: source frame of every slot of every output frame's SR window,
    exactly as `enhance_video` (`enhancement_engine.py:221-240`) followed by `forward` (`:141-160`) select them."""
    half = (2 * max(recovery_temporal_window, sr_temporal_window) + 1) // 2
    want = 2 * sr_temporal_window + 1
    rows: List[List[int]] = []
    for t in range(num_frames):
        start, end = max(0, t - half), min(num_frames, t + half + 1)          # the engine's frame window
        c = t - start                                                         # centre inside that window
        s0, e0 = max(0, c - sr_temporal_window), min(end - start, c + sr_temporal_window + 1)
        idx = [start + i for i in range(s0, e0)]
        idx += [idx[-1]] * (want - len(idx))                                  # pad by repeating the last frame
        rows.append(idx)
    return torch.tensor(rows, dtype=torch.int64)


@torch.no_grad()
def enhance_video(model: torch.nn.Module, video: Tensor, batch_size: int = 16, recovery_temporal_window: int = 2,
                  enhancement_strength: Optional[float] = None) -> Tensor:
    """Super-resolve every frame of ``video`` ((T,C,H,W) or (B,T,C,H,W)) with ``model`` (a ``SuperResolutionNet``).

    Returns (T,C,sH,sW) / (B,T,C,sH,sW).  Equivalent to the reference engine with ``frame_recovery_enabled=False``;
    ``enhancement_strength`` (default 1: no blend) mixes in the bicubic upsample of the frame as the engine does.
    Clips are independent: shard them over ranks for multi-GPU inference (no collective)."""
    squeeze = video.dim() == 4
    if squeeze:
        video = video.unsqueeze(0)
    if video.dim() != 5:
        raise RuntimeError("enhance_video: video must be (T,C,H,W) or (B,T,C,H,W)")
    B, T, C, H, W = video.shape
    idx = sr_window_indices(T, model.temporal_window, recovery_temporal_window).to(video.device)
    windows = video[:, idx]                                                   # (B, T, T', C, H, W), one device gather
    windows = windows.reshape(B * T, idx.shape[1], C, H, W)
    # inference means eval mode: in train mode BatchNorm would normalise with the statistics of whichever windows share
    # a batch (and update the running statistics), so the result would depend on batch_size -- the reference calls the
    # network per window.  The mode is restored afterwards.  (The tail batch, if smaller, gets its own shape plan.)
    was_training = model.training
    model.eval()
    try:
        outs = [model(windows[i:i + batch_size]) for i in range(0, B * T, batch_size)]
    finally:
        model.train(was_training)
    out = torch.cat(outs, 0)
    if enhancement_strength is not None and enhancement_strength < 1.0:
        from . import ops as _ops
        _ops.nv.bicubic_blend(out, video.reshape(B * T, C, H, W).float().contiguous(), out.shape[-1] // W,
                              float(enhancement_strength))
    out = out.view(B, T, *out.shape[1:])
    return out.squeeze(0) if squeeze else out
