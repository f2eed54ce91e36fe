"""CPU/ATen restatement of the NERVE-CL enhancement hot path -- TEST INFRASTRUCTURE ONLY.

This file is the *checker* for the CUDA path in
``continual-learning-for-dynamic-video-quality-enhancement_b200/``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may
import it.  The product never does (it fails loudly when ``libnervecl.so`` is missing).

What it restates (file:line are relative to the reference checkout):

* ``SuperResolutionNet.forward``                 nerve_cl/models/super_resolution.py:327-391
* ``FeatureExtractor`` / ``DepthwiseSeparableConv``  super_resolution.py:22-54, layers/efficient_layers.py:9-67
* ``LiteFlowNetCorrelation``                      layers/efficient_layers.py:313-343
* ``MotionEstimator.flow_net``                    super_resolution.py:74-82
* ``warp_features``                               super_resolution.py:104-143
* ``TemporalAggregator`` + ``CBAM``               super_resolution.py:146-209, efficient_layers.py:154-228
* ``ResidualDenseBlock``                          super_resolution.py:212-253
* ``PixelShuffleUpsampler`` + bicubic skip + clamp  efficient_layers.py:70-106, super_resolution.py:375-382

The arithmetic itself lives in a third-party dependency that is not vendored in the reference:
PyTorch (``torch>=2.0.0``, unpinned in pyproject.toml:36; 2.11.0+cu128 in this image).  The
restatement is therefore *functional* ATen code over a ``state_dict`` (no ``nn.Module`` tree), so the
same weights can be pushed through the reference module, through this file and through the CUDA path.

Parity pinning: the reference ships no golden vectors (its tests assert shapes only).  This oracle is
pinned against the *live* reference imported from ``/root/reference`` by
``tests/golden/make_golden.py`` (run in the build container; outputs committed under
``tests/golden/*.npz``) and re-checked on every CPU test run by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
RDB_GROWTH = 32          # super_resolution.py:215 (growth_rate default)
RDB_LAYERS = 5           # super_resolution.py:216
CORR_RADIUS = 4          # super_resolution.py:68
BN_EPS = 1e-5            # torch default; efficient_layers.py:59 passes only the channel count
BN_MOMENTUM = 0.1


# ---------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------
def dwsep_block(sd: Dict[str, Tensor], prefix: str, x: Tensor, training: bool) -> Tensor:
    """depthwise 3x3 (no bias) -> pointwise 1x1 (no bias) -> BatchNorm -> ReLU.

    efficient_layers.py:62-67.  In training mode the batch statistics are those of *this call*
    and the running buffers in ``sd`` are updated in place (momentum 0.1, unbiased running var).
    """
    c = x.shape[1]
    y = F.conv2d(x, sd[prefix + "depthwise.weight"], None, 1, 1, 1, groups=c)
    y = F.conv2d(y, sd[prefix + "pointwise.weight"], None)
    y = F.batch_norm(
        y,
        sd[prefix + "bn.running_mean"],
        sd[prefix + "bn.running_var"],
        sd[prefix + "bn.weight"],
        sd[prefix + "bn.bias"],
        training,
        BN_MOMENTUM,
        BN_EPS,
    )
    if training:
        sd[prefix + "bn.num_batches_tracked"] += 1
    return F.relu(y)


def extract_features(sd: Dict[str, Tensor], frame: Tensor, training: bool) -> Tensor:
    """super_resolution.py:51-54: head conv+ReLU, three dw-separable blocks, skip from the head."""
    head = F.relu(F.conv2d(frame, sd["feature_extractor.head.0.weight"],
                           sd["feature_extractor.head.0.bias"], 1, 1))
    y = head
    for i in range(3):
        y = dwsep_block(sd, f"feature_extractor.body.{i}.", y, training)
    return y + head


def correlation(x1: Tensor, x2: Tensor, radius: int = CORR_RADIUS) -> Tensor:
    """81-displacement cost volume, channel index i*9+j with i the vertical shift.

    efficient_layers.py:328-343:
    out[b, i*9+j, y, x] = (1/C) * sum_c x1[b,c,y,x] * x2pad[b,c,y+i,x+j]   (zero padding of 4).
    """
    b, c, h, w = x1.shape
    k = 2 * radius + 1
    x2p = F.pad(x2, [radius] * 4)
    planes = []
    for i in range(k):
        rows = x2p[:, :, i:i + h, :]
        for j in range(k):
            planes.append((x1 * rows[:, :, :, j:j + w]).sum(1, keepdim=True))
    return torch.cat(planes, 1) / c


def estimate_flow(sd: Dict[str, Tensor], src: Tensor, dst: Tensor) -> Tensor:
    """super_resolution.py:100-102: correlation -> 4 convs (ReLU between) -> (dx, dy)."""
    y = correlation(src, dst)
    for idx in (0, 2, 4, 6):
        p = f"motion_estimator.flow_net.{idx}."
        y = F.conv2d(y, sd[p + "weight"], sd[p + "bias"], 1, 1)
        if idx != 6:
            y = F.relu(y)
    return y


def normalised_grid(flow: Tensor) -> Tensor:
    """The exact op sequence of super_resolution.py:121-136 (order matters for bit-exact indices).

    x-channel: ((arange(W) + flow_x) * 2.0) / (W-1) - 1.0 ; same for y with H.  On CUDA ATen turns the
    division by the python scalar into a multiplication by ``1/(W-1)``; on CPU it divides.  This
    function just issues the same torch ops, so it reproduces whichever device it runs on.
    """
    b, _, h, w = flow.shape
    ys, xs = torch.meshgrid(
        torch.arange(h, device=flow.device, dtype=flow.dtype),
        torch.arange(w, device=flow.device, dtype=flow.dtype),
        indexing="ij",
    )
    g = torch.stack([xs, ys], 0).unsqueeze(0).expand(b, -1, -1, -1) + flow
    g[:, 0] = 2.0 * g[:, 0] / (w - 1) - 1.0
    g[:, 1] = 2.0 * g[:, 1] / (h - 1) - 1.0
    return g.permute(0, 2, 3, 1)


def warp(feat: Tensor, flow: Tensor) -> Tensor:
    """super_resolution.py:139-141: bilinear, zero padding, align_corners=True."""
    return F.grid_sample(feat, normalised_grid(flow), mode="bilinear",
                         padding_mode="zeros", align_corners=True)


def warp_corner_indices(flow: Tensor) -> Tensor:
    """Integer (x0, y0) = floor of the un-normalised sample position, as ATen computes it.

    ATen (cuda/GridSampler.cuh:23-31, align_corners=True): ix = ((g + 1) / 2) * (W - 1); x0 = floor(ix).
    Returns int32 (B, H, W, 2).  This is the quantity the CUDA warp kernel must reproduce bit-exactly.
    """
    g = normalised_grid(flow)
    _, h, w, _ = g.shape
    ix = ((g[..., 0] + 1.0) / 2) * (w - 1)
    iy = ((g[..., 1] + 1.0) / 2) * (h - 1)
    return torch.stack([torch.floor(ix), torch.floor(iy)], -1).to(torch.int32)


def channel_attention(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    """efficient_layers.py:176-180: GAP -> FC(no bias) -> ReLU -> FC(no bias) -> sigmoid -> scale."""
    p = "temporal_aggregator.refine.channel_attention.fc."
    s = x.mean((2, 3))
    s = torch.sigmoid(F.linear(F.relu(F.linear(s, sd[p + "0.weight"])), sd[p + "2.weight"]))
    return x * s[:, :, None, None]


def spatial_attention(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    """efficient_layers.py:200-205: [mean_c, max_c] -> 7x7 conv (no bias) -> sigmoid -> scale."""
    stats = torch.cat([x.mean(1, keepdim=True), torch.max(x, 1, keepdim=True)[0]], 1)
    gate = F.conv2d(stats, sd["temporal_aggregator.refine.spatial_attention.conv.weight"], None, 1, 3)
    return x * torch.sigmoid(gate)


def aggregate(sd: Dict[str, Tensor], aligned: List[Tensor]) -> Tensor:
    """super_resolution.py:194-209: frame-major concat -> 3 convs -> softmax over T -> blend -> CBAM."""
    stacked = torch.stack(aligned, 1)
    b, t, c, h, w = stacked.shape
    y = stacked.reshape(b, t * c, h, w)
    for idx in (0, 2, 4):
        p = f"temporal_aggregator.attention.{idx}."
        y = F.conv2d(y, sd[p + "weight"], sd[p + "bias"], 1, 1)
        if idx != 4:
            y = F.relu(y)
    attn = torch.softmax(y, 1)
    blended = (stacked * attn.unsqueeze(2)).sum(1)
    return spatial_attention(sd, channel_attention(sd, blended))


def residual_dense_block(sd: Dict[str, Tensor], k: int, x: Tensor) -> Tensor:
    """super_resolution.py:245-253."""
    feats = [x]
    for i in range(RDB_LAYERS):
        p = f"residual_blocks.{k}.layers.{i}.0."
        feats.append(F.relu(F.conv2d(torch.cat(feats, 1), sd[p + "weight"], sd[p + "bias"], 1, 1)))
    p = f"residual_blocks.{k}.lff."
    return F.conv2d(torch.cat(feats, 1), sd[p + "weight"], sd[p + "bias"]) * 0.2 + x


def count_blocks(sd: Dict[str, Tensor]) -> int:
    n = 0
    while f"residual_blocks.{n}.lff.weight" in sd:
        n += 1
    return n


# ---------------------------------------------------------------------------------------------
# whole network
# ---------------------------------------------------------------------------------------------
def sr_forward(
    sd: Dict[str, Tensor],
    lr_frames: Tensor,
    scale: int,
    training: bool = False,
    want_intermediate: bool = False,
):
    """``SuperResolutionNet.forward`` (super_resolution.py:327-391) over a state_dict.

    ``sd`` must hold parameters *and* BN buffers under the reference's key names; BN buffers are
    updated in place when ``training``.  Autograd works through it (parameters that require grad).
    """
    b, t, c, h, w = lr_frames.shape
    mid = t // 2
    feats = [extract_features(sd, lr_frames[:, i], training) for i in range(t)]
    aligned, flows = [], {}
    for i in range(t):
        if i == mid:
            aligned.append(feats[mid])
            continue
        flow = estimate_flow(sd, feats[i], feats[mid])
        flows[i] = flow
        aligned.append(warp(feats[i], flow))
    agg = aggregate(sd, aligned)
    y = agg
    for k in range(count_blocks(sd)):
        y = residual_dense_block(sd, k, y)
    fused = F.relu(F.conv2d(y, sd["gff.0.weight"], sd["gff.0.bias"], 1, 1)) + feats[mid]
    hr_res = F.pixel_shuffle(
        F.conv2d(fused, sd["upsampler.conv.weight"], sd["upsampler.conv.bias"], 1, 1), scale)
    base = F.interpolate(lr_frames[:, mid], scale_factor=scale, mode="bicubic", align_corners=False)
    out = torch.clamp(base + hr_res, 0, 1)
    if want_intermediate:
        return out, {"features": feats, "aligned": aligned, "aggregated": agg, "flows": flows}
    return out


def clone_state(sd: Dict[str, Tensor], requires_grad: bool = False,
                device: Optional[torch.device] = None) -> Dict[str, Tensor]:
    """Detached copy of a state_dict; floating-point *parameters* optionally become autograd leaves."""
    out = {}
    for k, v in sd.items():
        v = v.detach().clone()
        if device is not None:
            v = v.to(device)
        is_buffer = k.endswith(("running_mean", "running_var", "num_batches_tracked"))
        if requires_grad and v.is_floating_point() and not is_buffer:
            v.requires_grad_(True)
        out[k] = v
    return out


def param_names(sd: Dict[str, Tensor]) -> List[str]:
    return [k for k in sd
            if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]


def train_step_grads(sd: Dict[str, Tensor], lr_frames: Tensor, target: Tensor, scale: int,
                     training: bool = True) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """fwd + ``mse_loss`` + bwd, as experiments/train_baseline.py:85-87 does.  Returns (out, loss, grads)."""
    work = clone_state(sd, requires_grad=True)
    out = sr_forward(work, lr_frames, scale, training)
    loss = F.mse_loss(out, target)
    names = param_names(work)
    grads = torch.autograd.grad(loss, [work[n] for n in names], allow_unused=True)
    for k in sd:  # propagate BN buffer updates back to the caller's dict
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            sd[k].copy_(work[k])
    return out.detach(), loss.detach(), {n: g for n, g in zip(names, grads)}


def random_state_dict(scale: int = 2, features: int = 64, blocks: int = 8, frames: int = 3,
                      seed: int = 0) -> Dict[str, Tensor]:
    """A state_dict with the reference's key names and shapes and fan-in scaled random values,
    for timing the port without importing either the reference or the product package."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(name, cout, cin, k, bias=True):
        sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) / (cin * k * k) ** 0.5
        if bias:
            sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) / (cin * k * k) ** 0.5

    F_ = features
    conv("feature_extractor.head.0", F_, 3, 3)
    for j in range(3):
        pre = f"feature_extractor.body.{j}."
        sd[pre + "depthwise.weight"] = (torch.rand(F_, 1, 3, 3, generator=g) * 2 - 1) / 3.0
        conv(pre + "pointwise", F_, F_, 1, bias=False)
        sd[pre + "bn.weight"], sd[pre + "bn.bias"] = torch.ones(F_), torch.zeros(F_)
        sd[pre + "bn.running_mean"], sd[pre + "bn.running_var"] = torch.zeros(F_), torch.ones(F_)
        sd[pre + "bn.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    for idx, (ci, co) in zip((0, 2, 4, 6), ((81, 128), (128, 64), (64, 32), (32, 2))):
        conv(f"motion_estimator.flow_net.{idx}", co, ci, 3)
    conv("temporal_aggregator.attention.0", F_, F_ * frames, 3)
    conv("temporal_aggregator.attention.2", F_, F_, 3)
    conv("temporal_aggregator.attention.4", frames, F_, 3)
    r = max(F_ // 16, 1)
    sd["temporal_aggregator.refine.channel_attention.fc.0.weight"] = (torch.rand(r, F_, generator=g) * 2 - 1) / F_ ** 0.5
    sd["temporal_aggregator.refine.channel_attention.fc.2.weight"] = (torch.rand(F_, r, generator=g) * 2 - 1) / r ** 0.5
    sd["temporal_aggregator.refine.spatial_attention.conv.weight"] = (torch.rand(1, 2, 7, 7, generator=g) * 2 - 1) / 98 ** 0.5
    for k in range(blocks):
        for i in range(5):
            conv(f"residual_blocks.{k}.layers.{i}.0", 32, F_ + 32 * i, 3)
        conv(f"residual_blocks.{k}.lff", F_, F_ + 160, 1)
    conv("gff.0", F_, F_, 3)
    conv("upsampler.conv", 3 * scale * scale, F_, 3)
    return sd
