#!/usr/bin/env python
"""Golden fixture for the batched ``enhance_video``: which source frames the reference's sliding-window loop
(`nerve_cl/models/enhancement_engine.py:187-245` + the SR window selection at `:141-166`) hands to the SR network
for every output frame -- including its edge behaviour (windows clipped at the clip borders are padded by repeating
the LAST frame, so near the leading edge the network's centre slot does not hold frame t).

Run in the build container (needs /root/reference):  python tests/golden/make_enhance_windows.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, "/root/reference")
from nerve_cl.models.enhancement_engine import EnhancementConfig, EnhancementEngine  # noqa: E402


class Recorder(nn.Module):
    """Stands in for SuperResolutionNet: records the frame ids of every window it is given."""

    def __init__(self):
        super().__init__()
        self.seen = []

    def forward(self, frames):                     # (B, T', C, H, W); pixel value == frame id
        self.seen.append(frames[0, :, 0, 0, 0].round().long().tolist())
        return frames[:, frames.shape[1] // 2]


out = {}
for sr_w in (1, 2):
    for rec_w in (1, 2, 3):
        for T in (1, 2, 3, 4, 7, 9):
            cfg = EnhancementConfig(frame_recovery_enabled=False, super_resolution_enabled=True,
                                    recovery_temporal_window=rec_w, sr_temporal_window=sr_w)
            eng = EnhancementEngine(cfg)
            rec = Recorder()
            eng.super_resolution = rec
            video = torch.arange(T, dtype=torch.float32).view(1, T, 1, 1, 1).expand(1, T, 3, 4, 4).clone()
            with torch.no_grad():
                eng.enhance_video(video)
            out[f"sr{sr_w}_rec{rec_w}_T{T}"] = np.asarray(rec.seen, dtype=np.int64)

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "enhance_windows.npz")
np.savez_compressed(path, **out)
print("wrote", path, len(out), "cases")
