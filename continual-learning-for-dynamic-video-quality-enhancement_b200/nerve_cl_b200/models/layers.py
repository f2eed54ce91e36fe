"""Layer classes of the reference's ``nerve_cl/models/layers/efficient_layers.py`` that sit on the
SuperResolutionNet hot path.

Inside ``SuperResolutionNet`` these modules only hold parameters (the engine runs the arithmetic).
Used stand-alone (NCHW fp32 tensors, like the reference) ``LiteFlowNetCorrelation`` and the free
function ``warp_features`` run the same CUDA kernels behind an autograd node, converting NCHW<->NHWC
at the boundary; the remaining holders raise if called directly.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops as _ops

Tensor = torch.Tensor
nv = _ops.nv


def _to_nhwc(x: Tensor, dtype: torch.dtype = torch.float32, pad_to: int = 0) -> Tensor:
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, max(c, pad_to)), device=x.device, dtype=dtype)
    nv.nchw_to_nhwc(x.contiguous().float(), out[..., :c])
    return out[..., :c]


def _to_nchw(x: Tensor) -> Tensor:
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
    nv.nhwc_to_nchw(x, out)
    return out


def _require_cuda(*ts: Tensor) -> None:
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("nerve_cl_b200 layers run on CUDA (sm_100a) only; there is no CPU fallback")


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError(
            f"{type(self).__name__} only holds parameters in nerve_cl_b200; its arithmetic runs inside "
            "SuperResolutionNet's fused engine")


class DepthwiseSeparableConv(_Holder):
    """depthwise 3x3 -> pointwise 1x1 -> BatchNorm -> ReLU (efficient_layers.py:9-67); bias-free convs."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1,
                 padding: int = 1, bias: bool = False):
        super().__init__()
        self.depthwise = nn.Conv2d(in_channels, in_channels, kernel_size, stride, padding,
                                   groups=in_channels, bias=bias)
        self.pointwise = nn.Conv2d(in_channels, out_channels, 1, 1, 0, bias=bias)
        self.bn = nn.BatchNorm2d(out_channels)
        self.act = nn.ReLU(inplace=True)


class PixelShuffleUpsampler(_Holder):
    """conv 3x3 -> PixelShuffle (efficient_layers.py:70-106)."""

    def __init__(self, in_channels: int, scale_factor: int = 2, out_channels: int = 3):
        super().__init__()
        self.scale_factor = scale_factor
        self.conv = nn.Conv2d(in_channels, out_channels * scale_factor ** 2, 3, 1, 1)
        self.pixel_shuffle = nn.PixelShuffle(scale_factor)


class ChannelAttention(_Holder):
    """GAP -> FC -> ReLU -> FC -> sigmoid gate (efficient_layers.py:154-180); FCs have no bias."""

    def __init__(self, channels: int, reduction: int = 16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channels, channels // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channels // reduction, channels, bias=False), nn.Sigmoid())


class SpatialAttention(_Holder):
    """[mean_c, max_c] -> 7x7 conv -> sigmoid gate (efficient_layers.py:183-205)."""

    def __init__(self, kernel_size: int = 7):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)
        self.sigmoid = nn.Sigmoid()


class CBAM(_Holder):
    """Channel then spatial attention (efficient_layers.py:208-228)."""

    def __init__(self, channels: int, reduction: int = 16):
        super().__init__()
        self.channel_attention = ChannelAttention(channels, reduction)
        self.spatial_attention = SpatialAttention()


# ---------------------------------------------------------------------------------------------
# stand-alone differentiable ops
# ---------------------------------------------------------------------------------------------
class _CorrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1: Tensor, x2: Tensor):
        a, b = _to_nhwc(x1), _to_nhwc(x2)
        n, h, w, c = a.shape
        out = torch.empty((n, h, w, 96), device=a.device, dtype=torch.float32)
        nv.corr_fwd(a, b, out)
        ctx.save_for_backward(a, b)
        return _to_nchw(out[..., :81])

    @staticmethod
    def backward(ctx, g: Tensor):
        a, b = ctx.saved_tensors
        gn = _to_nhwc(g, pad_to=96)
        d1, d2 = torch.empty_like(a), torch.empty_like(b)
        nv.corr_bwd(a, b, gn, d1, False, d2, False)
        return _to_nchw(d1), _to_nchw(d2)


class LiteFlowNetCorrelation(nn.Module):
    """81-displacement correlation (efficient_layers.py:297-343), one fused kernel each way."""

    def __init__(self, max_displacement: int = 4):
        super().__init__()
        if max_displacement != 4:
            raise ValueError("the correlation kernel is specialised to max_displacement=4 (reference default)")
        self.max_displacement = max_displacement
        self.pad = max_displacement

    def forward(self, x1: Tensor, x2: Tensor) -> Tensor:
        _require_cuda(x1, x2)
        if x1.shape != x2.shape or x1.shape[1] % 8:
            raise ValueError("correlation needs equal shapes and a channel count divisible by 8")
        return _CorrFn.apply(x1, x2)


class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features: Tensor, flow: Tensor, div_mode: int):
        f = _to_nhwc(features)
        fl = flow.permute(0, 2, 3, 1).contiguous().float()
        out = torch.empty_like(f)
        nv.warp_fwd(f, fl, out, div_mode, None)
        ctx.save_for_backward(f, fl)
        ctx.div_mode = div_mode
        return _to_nchw(out)

    @staticmethod
    def backward(ctx, g: Tensor):
        f, fl = ctx.saved_tensors
        gn = _to_nhwc(g)
        dfeat = torch.zeros_like(f)
        dflow = torch.empty_like(fl)
        nv.warp_bwd(f, fl, gn, dfeat, dflow, ctx.div_mode)
        return _to_nchw(dfeat), dflow.permute(0, 3, 1, 2).contiguous(), None


def warp_features(features: Tensor, flow: Tensor, div_mode: int = 0) -> Tensor:
    """Reference ``warp_features`` (super_resolution.py:104-143): features (B,C,H,W), flow (B,2,H,W)."""
    _require_cuda(features, flow)
    return _WarpFn.apply(features, flow, div_mode)


def warp_indices(flow: Tensor, div_mode: int = 0) -> Tensor:
    """int32 (B,H,W,2) top-left neighbour (x0, y0) the warp kernel samples for ``flow`` (B,2,H,W)."""
    _require_cuda(flow)
    b, _, h, w = flow.shape
    fl = flow.permute(0, 2, 3, 1).contiguous().float()
    dummy = torch.zeros((b, h, w, 8), device=flow.device, dtype=torch.float32)
    out = torch.empty_like(dummy)
    idx = torch.empty((b, h, w, 2), device=flow.device, dtype=torch.int32)
    nv.warp_fwd(dummy, fl, out, div_mode, idx)
    return idx


class ResidualBlock(_Holder):
    """dw-separable residual block (efficient_layers.py:109-151): conv1 = DepthwiseSeparableConv, conv2 = dw 3x3 ->
    pw 1x1 -> BN, ``relu(conv2(conv1(x)) + x)``.  Holder: the arithmetic runs in ``frame_recovery._Runner``."""

    def __init__(self, channels: int, use_efficient: bool = True):
        super().__init__()
        if not use_efficient:
            raise ValueError("nerve_cl_b200.ResidualBlock implements the reference default use_efficient=True only")
        self.conv1 = DepthwiseSeparableConv(channels, channels)
        self.conv2 = nn.Sequential(
            nn.Conv2d(channels, channels, 3, 1, 1, groups=channels, bias=False),
            nn.Conv2d(channels, channels, 1, 1, 0, bias=False),
            nn.BatchNorm2d(channels),
        )
        self.relu = nn.ReLU(inplace=True)


class TemporalConv3D(_Holder):
    """(2+1)D factorised 3-D convolution (efficient_layers.py:231-294): (1,3,3) spatial conv -> BN3d -> ReLU ->
    (k,1,1) temporal conv -> BN3d -> ReLU, with the reference's intermediate width formula."""

    def __init__(self, in_channels: int, out_channels: int, temporal_kernel: int = 3):
        super().__init__()
        if temporal_kernel != 3:
            raise ValueError("nerve_cl_b200.TemporalConv3D implements temporal_kernel=3 (the only value the reference uses)")
        mid = (in_channels * out_channels * 3 * 3 * temporal_kernel) // (in_channels * 3 * 3 + out_channels * temporal_kernel)
        mid = max(mid, out_channels // 2)
        self.mid_channels = mid
        self.spatial = nn.Sequential(
            nn.Conv3d(in_channels, mid, kernel_size=(1, 3, 3), stride=1, padding=(0, 1, 1), bias=False),
            nn.BatchNorm3d(mid), nn.ReLU(inplace=True))
        self.temporal = nn.Sequential(
            nn.Conv3d(mid, out_channels, kernel_size=(temporal_kernel, 1, 1), stride=1, padding=(temporal_kernel // 2, 0, 0),
                      bias=False),
            nn.BatchNorm3d(out_channels), nn.ReLU(inplace=True))
