"""CPU/ATen restatement of the callers either side of the SR hot path -- TEST INFRASTRUCTURE ONLY (see sr_oracle.py).

* ``frame_recovery_forward``   FrameRecoveryNet.forward, eval mode      nerve_cl/models/frame_recovery.py:386-443
  (SpatialEncoder :93-108, TemporalEncoder :136-167, FusionModule :221-257, Decoder :305-332; layers
  ResidualBlock efficient_layers.py:142-151, TemporalConv3D :283-294, CBAM :225-228)
* ``lightweight_forward``      LightweightSuperResolution.forward        nerve_cl/models/super_resolution.py:467-470
* ``engine_forward``           EnhancementEngine.forward                 nerve_cl/models/enhancement_engine.py:95-184
* ``engine_window_table``      the sliding-window loop of enhance_video  enhancement_engine.py:214-228

Functional code over ``state_dict``s (no nn.Module tree).  Pinned against the live reference by
``tests/golden/make_engine_golden.py`` -> ``tests/golden/{recovery,lightweight,engine}_*.npz`` and re-checked on every
CPU test run by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import sr_oracle

Tensor = torch.Tensor
EPS = 1e-5


def _bn(sd, pre, x, training=False):
    if training and pre + "num_batches_tracked" in sd:
        sd[pre + "num_batches_tracked"] += 1
    return F.batch_norm(x, sd[pre + "running_mean"], sd[pre + "running_var"], sd[pre + "weight"], sd[pre + "bias"],
                        training, 0.1, EPS)


def _dwsep(sd, pre, x, training=False):
    """DepthwiseSeparableConv (efficient_layers.py:62-67)."""
    c = x.shape[1]
    x = F.conv2d(x, sd[pre + "depthwise.weight"], None, 1, 1, 1, c)
    x = F.conv2d(x, sd[pre + "pointwise.weight"])
    return F.relu(_bn(sd, pre + "bn.", x, training))


def _resblock(sd, pre, x):
    """ResidualBlock, use_efficient=True (efficient_layers.py:142-151)."""
    c = x.shape[1]
    y = _dwsep(sd, pre + "conv1.", x)
    y = F.conv2d(y, sd[pre + "conv2.0.weight"], None, 1, 1, 1, c)
    y = F.conv2d(y, sd[pre + "conv2.1.weight"])
    y = _bn(sd, pre + "conv2.2.", y)
    return F.relu(y + x)


def _cbam(sd, pre, x):
    """CBAM: channel then spatial attention (efficient_layers.py:154-228)."""
    g = x.mean((2, 3))
    g = torch.sigmoid(F.linear(F.relu(F.linear(g, sd[pre + "channel_attention.fc.0.weight"])),
                               sd[pre + "channel_attention.fc.2.weight"]))
    x = x * g[:, :, None, None]
    st = torch.cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1)
    return x * torch.sigmoid(F.conv2d(st, sd[pre + "spatial_attention.conv.weight"], None, 1, 3))


def _tconv3d(sd, pre, x):
    """TemporalConv3D (efficient_layers.py:283-294) on (B, C, T, H, W)."""
    x = F.relu(F.batch_norm(F.conv3d(x, sd[pre + "spatial.0.weight"], None, 1, (0, 1, 1)), sd[pre + "spatial.1.running_mean"],
                            sd[pre + "spatial.1.running_var"], sd[pre + "spatial.1.weight"], sd[pre + "spatial.1.bias"],
                            False, 0.1, EPS))
    x = F.relu(F.batch_norm(F.conv3d(x, sd[pre + "temporal.0.weight"], None, 1, (1, 0, 0)), sd[pre + "temporal.1.running_mean"],
                            sd[pre + "temporal.1.running_var"], sd[pre + "temporal.1.weight"], sd[pre + "temporal.1.bias"],
                            False, 0.1, EPS))
    return x


def frame_recovery_forward(sd: Dict[str, Tensor], corrupted: Tensor, refs: Tensor, mask: Optional[Tensor]) -> Tensor:
    """FrameRecoveryNet.forward in eval mode (frame_recovery.py:386-443)."""
    b, c, h, w = corrupted.shape
    if mask is None:
        mask = torch.zeros(b, 1, h, w)
    x = torch.cat([corrupted, mask], 1)
    # SpatialEncoder (:93-108)
    x = F.relu(_bn(sd, "spatial_encoder.stem.1.", F.conv2d(x, sd["spatial_encoder.stem.0.weight"], None, 2, 3)))
    x = F.max_pool2d(x, 3, 2, 1)
    for stage in ("stage1", "stage2", "stage3"):
        pre = f"spatial_encoder.{stage}."
        i = 0
        if pre + "0.0.weight" in sd:                                     # 1x1 stride-2 shortcut + BN (:70-74)
            x = _bn(sd, pre + "0.1.", F.conv2d(x, sd[pre + "0.0.weight"], None, 2))
            i = 1
        while pre + f"{i}.conv1.depthwise.weight" in sd:
            x = _resblock(sd, pre + f"{i}.", x)
            i += 1
    sp = _cbam(sd, "spatial_encoder.attention.", x)
    # TemporalEncoder (:136-167)
    t = refs.permute(0, 2, 1, 3, 4)
    t = F.max_pool3d(_tconv3d(sd, "temporal_encoder.conv1.", t), (1, 2, 2))
    t = F.max_pool3d(_tconv3d(sd, "temporal_encoder.conv2.", t), (1, 2, 2))
    t = _tconv3d(sd, "temporal_encoder.conv3.", t).mean(2)
    # FusionModule (:221-257)
    if sp.shape[2:] != t.shape[2:]:
        t = F.interpolate(t, size=sp.shape[2:], mode="bilinear", align_corners=False)
    aligned = F.conv2d(torch.cat([sp, t], 1), sd["fusion.align.weight"], sd["fusion.align.bias"])
    a = F.conv2d(F.relu(F.conv2d(aligned, sd["fusion.attention.0.weight"], sd["fusion.attention.0.bias"])),
                 sd["fusion.attention.2.weight"], sd["fusion.attention.2.bias"]).softmax(1)
    fused = a[:, 0:1] * sp.mean(1, keepdim=True) + a[:, 1:2] * t.mean(1, keepdim=True)     # all-ones/C 1x1 convs
    y = aligned + fused
    y = _resblock(sd, "fusion.refine.0.", y)
    y = _resblock(sd, "fusion.refine.1.", y)
    y = _cbam(sd, "fusion.refine.2.", y)
    # Decoder (:305-332)
    for k in (1, 2, 3, 4):
        y = F.relu(_bn(sd, f"decoder.up{k}.1.", F.conv_transpose2d(y, sd[f"decoder.up{k}.0.weight"], None, 2, 1)))
    y = torch.tanh(F.conv2d(y, sd["decoder.final.0.weight"], sd["decoder.final.0.bias"], 1, 1))
    if y.shape[2:] != (h, w):
        y = F.interpolate(y, size=(h, w), mode="bilinear", align_corners=False)
    return corrupted * (1 - mask) + y * mask


def lightweight_forward(sd: Dict[str, Tensor], x: Tensor, scale: int, training: bool = False) -> Tensor:
    """LightweightSuperResolution.forward (super_resolution.py:446-470); BN buffers updated in place when training."""
    y = F.relu(F.conv2d(x, sd["net.0.weight"], sd["net.0.bias"], 1, 1))
    for j in range(2, 6):
        y = _dwsep(sd, f"net.{j}.", y, training)
    y = F.pixel_shuffle(F.conv2d(y, sd["net.6.weight"], sd["net.6.bias"], 1, 1), scale)
    return torch.clamp(F.interpolate(x, scale_factor=scale, mode="bicubic", align_corners=False) + y, 0, 1)


def engine_forward(sr_sd: Optional[Dict[str, Tensor]], rec_sd: Optional[Dict[str, Tensor]], frames: Tensor, scale: int,
                   sr_window: int, center_idx: Optional[int] = None, mask: Optional[Tensor] = None,
                   strength: float = 1.0, lightweight: bool = False) -> Dict[str, Tensor]:
    """EnhancementEngine.forward (enhancement_engine.py:95-184) over the two networks' state_dicts (eval mode)."""
    b, t, c, h, w = frames.shape
    ci = t // 2 if center_idx is None else center_idx
    res: Dict[str, Tensor] = {}
    cur = frames[:, ci]
    refs = [i for i in range(t) if i != ci]
    if rec_sd is not None and mask is not None and float(mask.sum()) > 0:
        cur = frame_recovery_forward(rec_sd, cur, frames[:, refs], mask)
        res["recovered"] = cur
    if sr_sd is not None:
        s0, e0 = max(0, ci - sr_window), min(t, ci + sr_window + 1)
        win = frames[:, s0:e0]
        want = 2 * sr_window + 1
        if win.shape[1] < want:
            win = torch.cat([win, win[:, -1:].expand(-1, want - win.shape[1], -1, -1, -1)], 1)
        cur = lightweight_forward(sr_sd, cur, scale) if lightweight else sr_oracle.sr_forward(sr_sd, win, scale, False)
        res["super_resolved"] = cur
    if strength < 1.0 and "super_resolved" in res:
        bic = F.interpolate(frames[:, ci], size=cur.shape[2:], mode="bicubic", align_corners=False)
        cur = strength * cur + (1 - strength) * bic
    res["enhanced"] = cur
    return res


def engine_window_table(num_frames: int, recovery_window: int, sr_window: int) -> List[Tuple[int, int, int]]:
    """(start, end, centre index inside the window) per output frame (enhancement_engine.py:214-228)."""
    size = 2 * max(recovery_window, sr_window) + 1
    out = []
    for t in range(num_frames):
        start, end = max(0, t - size // 2), min(num_frames, t + size // 2 + 1)
        out.append((start, end, t - start))
    return out
