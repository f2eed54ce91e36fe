"""Batched clip inference: window table pinned to the reference's enhance_video, batched run."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import psnr  # noqa: F401  (path setup)


def test_sr_window_indices_match_reference_enhance_video():
    from nerve_cl_b200.inference import sr_window_indices
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "enhance_windows.npz"))
    assert len(g.files) == 36
    for key in g.files:
        sr_w, rec_w, T = map(int, re.match(r"sr(\d+)_rec(\d+)_T(\d+)", key).groups())
        got = sr_window_indices(T, sr_w, rec_w).numpy()
        assert np.array_equal(got, g[key]), key


def test_sr_window_indices_equal_the_engine_forward_selection():
    """The SR-only path of EnhancementEngine.enhance_video batches every frame through `sr_window_indices`; the general
    path cuts the engine window (`window_table`) and lets `forward` slice / pad the SR frames out of it.  Same frames,
    for every clip length and window configuration (host logic only: no GPU)."""
    from nerve_cl_b200.inference import sr_window_indices
    from nerve_cl_b200.models import EnhancementConfig, EnhancementEngine
    for sr_w in (1, 2, 3):
        for rec_w in (1, 2, 3):
            eng = EnhancementEngine.__new__(EnhancementEngine)          # window_table only reads the config
            eng.config = EnhancementConfig(sr_temporal_window=sr_w, recovery_temporal_window=rec_w)
            for T in (1, 2, 3, 5, 8, 13):
                idx = sr_window_indices(T, sr_w, rec_w)
                want = 2 * sr_w + 1
                assert idx.shape == (T, want)
                for t, (start, end, c) in enumerate(eng.window_table(T)):
                    s0, e0 = max(0, c - sr_w), min(end - start, c + sr_w + 1)      # EnhancementEngine.forward
                    sel = list(range(start + s0, start + e0))
                    sel += [sel[-1]] * (want - len(sel))
                    assert idx[t].tolist() == sel, (sr_w, rec_w, T, t)


@pytest.mark.gpu
def test_enhance_video_batched_equals_per_frame_loop():
    from nerve_cl_b200.inference import enhance_video, sr_window_indices
    from nerve_cl_b200.models import SuperResolutionNet
    torch.manual_seed(3)
    model = SuperResolutionNet(scale_factor=2, num_features=16, num_residual_blocks=1).cuda().eval()
    video = torch.rand(2, 5, 3, 16, 72, device="cuda")
    out = enhance_video(model, video, batch_size=4)
    assert out.shape == (2, 5, 3, 32, 144)
    idx = sr_window_indices(5, 1, 2)
    with torch.no_grad():
        for t in range(5):
            ref = model(video[:, idx[t].tolist()])
            assert float((out[:, t] - ref).abs().max()) <= 1e-5, t     # same kernels, same inputs (atomics order aside)
    single = enhance_video(model, video[0], batch_size=16)
    assert float((single - out[0]).abs().max()) <= 1e-5
    blend = enhance_video(model, video, batch_size=16, enhancement_strength=0.25)
    assert float((blend - out).abs().max()) > 0
