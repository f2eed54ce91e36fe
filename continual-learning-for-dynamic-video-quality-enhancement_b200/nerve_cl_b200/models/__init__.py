"""Mirror of ``nerve_cl.models`` for the hot path."""
from .super_resolution import SuperResolutionNet, LightweightSuperResolution
from .layers import (DepthwiseSeparableConv, PixelShuffleUpsampler, CBAM, ChannelAttention, SpatialAttention,
                     LiteFlowNetCorrelation, warp_features, warp_indices)

__all__ = ["SuperResolutionNet", "LightweightSuperResolution", "DepthwiseSeparableConv", "PixelShuffleUpsampler",
           "CBAM", "ChannelAttention", "SpatialAttention", "LiteFlowNetCorrelation", "warp_features", "warp_indices"]
