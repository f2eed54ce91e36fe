"""Golden fixtures for the rows either side of the SR hot path, generated from the LIVE
reference (``nerve_cl`` imported from /root/reference) and checked against ``oracle/recovery_oracle.py``:

    python tests/golden/make_engine_golden.py          # build container only

* recovery_b16.npz      FrameRecoveryNet(base_channels=16) eval forward: 64x96 (decoder output = input size) and
                        70x90 (bilinear resize tail), 4 reference frames, rectangular + all-zero masks
* lightweight_x2.npz    LightweightSuperResolution x2: eval forward and one train step (output, loss, every gradient,
                        BatchNorm buffers after the step)
* engine_small.npz      EnhancementEngine (recovery 16 ch + SR 16 feat / 1 block, x2): forward dicts (with / without a
                        mask, strength 1 and 0.6, an off-centre clipped window), enhance_video over a 7-frame clip with
                        masks, and the SR-only engine's enhance_video
Weights are a function of the seed (the drop-in modules build the same torch layers in the same order); BatchNorm
running statistics are perturbed from a seeded generator so that eval-mode BatchNorm is not the identity.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from nerve_cl.models import (EnhancementConfig, EnhancementEngine, FrameRecoveryNet,  # noqa: E402  (the reference)
                             LightweightSuperResolution)
from oracle import recovery_oracle as ro  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def relerr(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def save(name, **arrays):
    out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in arrays.items()}
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(f"wrote {name}: {os.path.getsize(path)/1024:.1f} KiB")


def perturb_bn(module, seed):
    """Deterministic non-trivial running statistics for every BatchNorm, in module order."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            with torch.no_grad():
                m.running_mean.copy_(0.05 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(1.0 + 0.2 * torch.rand(m.num_features, generator=g))


def recovery_case():
    torch.manual_seed(40)
    net = FrameRecoveryNet(base_channels=16, temporal_window=2).eval()
    perturb_bn(net, 41)
    sd = net.state_dict()
    g = torch.Generator().manual_seed(42)
    out = {}
    for tag, (h, w) in (("a", (64, 96)), ("b", (70, 90))):
        frame = torch.rand(2, 3, h, w, generator=g)
        refs = torch.rand(2, 4, 3, h, w, generator=g)
        mask = torch.zeros(2, 1, h, w)
        mask[0, :, 10:40, 20:70] = 1
        mask[1, :, 30:60, 5:50] = 1
        with torch.no_grad():
            y = net(frame, refs, mask)
            y0 = net(frame, refs, None)
            o = ro.frame_recovery_forward(sd, frame, refs, mask)
        assert relerr(o, y) < 1e-5, relerr(o, y)
        assert torch.equal(y0, frame)                         # no mask: recovered region is empty
        out.update({f"{tag}/frame": frame, f"{tag}/refs": refs, f"{tag}/mask": mask, f"{tag}/out": y})
    # two reference frames only (a clip border), full-frame mask: the output IS the decoder's picture
    frame, refs = torch.rand(1, 3, 64, 64, generator=g), torch.rand(1, 2, 3, 64, 64, generator=g)
    mask = torch.ones(1, 1, 64, 64)
    with torch.no_grad():
        y = net(frame, refs, mask)
        assert relerr(ro.frame_recovery_forward(sd, frame, refs, mask), y) < 1e-5
    out.update({"c/frame": frame, "c/refs": refs, "c/mask": mask, "c/out": y})
    save("recovery_b16.npz", meta=np.array([40, 41, 16, 2]), **out)


def lightweight_case():
    torch.manual_seed(50)
    net = LightweightSuperResolution(scale_factor=2)
    perturb_bn(net, 51)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(52)
    x = torch.rand(2, 3, 20, 28, generator=g)
    tgt = torch.rand(2, 3, 40, 56, generator=g)
    net.eval()
    with torch.no_grad():
        y_eval = net(x)
        assert relerr(ro.lightweight_forward(sd0, x, 2), y_eval) < 1e-6
    net.train()
    out = net(x)
    loss = torch.nn.functional.mse_loss(out, tgt)
    loss.backward()
    sd_o = {k: v.clone() for k, v in sd0.items()}
    assert relerr(ro.lightweight_forward(sd_o, x, 2, training=True), out) < 1e-6
    arrays = {"x": x, "target": tgt, "out_eval": y_eval, "out_train": out, "loss": loss.detach()}
    for n, p in net.named_parameters():
        arrays["g/" + n] = p.grad
    for k, v in net.state_dict().items():
        if "running" in k or "tracked" in k:
            arrays["bn1/" + k] = v
            assert torch.allclose(sd_o[k].float(), v.float(), rtol=1e-6, atol=1e-7)
    save("lightweight_x2.npz", meta=np.array([50, 51, 2]), **arrays)


def engine_case():
    cfg = dict(frame_recovery_enabled=True, recovery_base_channels=16, recovery_temporal_window=2,
               super_resolution_enabled=True, scale_factor=2, sr_num_features=16, sr_num_residual_blocks=1,
               sr_temporal_window=1)
    torch.manual_seed(60)
    eng = EnhancementEngine(EnhancementConfig(**cfg)).eval()
    perturb_bn(eng, 61)
    g = torch.Generator().manual_seed(62)
    frames = torch.rand(1, 5, 3, 64, 64, generator=g)
    mask = torch.zeros(1, 1, 64, 64)
    mask[:, :, 20:50, 10:40] = 1
    arrays = {"frames": frames, "mask": mask}
    rec_sd = eng.frame_recovery.state_dict()
    sr_sd = eng.super_resolution.state_dict()
    with torch.no_grad():
        for tag, kw in (("plain", {}), ("masked", {"corruption_mask": mask}),
                        ("s06", {"corruption_mask": mask, "enhancement_strength": 0.6}),
                        ("edge", {"center_idx": 0, "corruption_mask": mask, "enhancement_strength": 0.8})):
            res = eng(frames, **kw)
            o = ro.engine_forward(sr_sd, rec_sd, frames, 2, 1, kw.get("center_idx"), kw.get("corruption_mask"),
                                  kw.get("enhancement_strength", 1.0))
            assert set(o) == set(res)
            for k, v in res.items():
                assert relerr(o[k], v) < 2e-5, (tag, k, relerr(o[k], v))
                if k == "enhanced" or (tag == "masked" and k == "recovered"):
                    arrays[f"fwd/{tag}/{k}"] = v
            arrays[f"fwd/{tag}/keys"] = np.array(sorted(res))
        # enhance_video: 7-frame clip, per-frame masks (frames 2 and 5 corrupted)
        video = torch.rand(7, 3, 64, 64, generator=g)
        masks = torch.zeros(7, 1, 64, 64)
        masks[2, :, 8:40, 8:56] = 1
        masks[5, :, 30:60, 20:44] = 1
        arrays["video"], arrays["masks"] = video, masks
        full = eng.enhance_video(video, masks)
        arrays["video_out"] = full[..., ::2, ::2]              # (every second pixel: keeps the fixture small)
        arrays["video_out_nomask"] = eng.enhance_video(video)[..., ::2, ::2]
        table = ro.engine_window_table(7, 2, 1)
        arrays["window_table"] = np.array(table)
        # the oracle's loop reproduces enhance_video
        outs = []
        for t, (s, e, c) in enumerate(table):
            outs.append(ro.engine_forward(sr_sd, rec_sd, video[None, s:e], 2, 1, c, masks[t:t + 1])["enhanced"])
        assert relerr(torch.stack(outs, 1)[0], full) < 2e-5
    # SR-only engine (the launcher's configuration, train_continual.py:125-128)
    torch.manual_seed(63)
    eng2 = EnhancementEngine(EnhancementConfig(frame_recovery_enabled=False, sr_num_features=16, sr_num_residual_blocks=1)).eval()
    perturb_bn(eng2, 64)
    with torch.no_grad():
        arrays["sronly/video_out"] = eng2.enhance_video(video)[..., ::2, ::2]
        arrays["sronly/fwd"] = eng2(frames)["enhanced"]
    save("engine_small.npz", meta=np.array([60, 61, 63, 64]), **arrays)


if __name__ == "__main__":
    recovery_case()
    lightweight_case()
    engine_case()
