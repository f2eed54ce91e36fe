"""Mirror of ``nerve_cl.models`` for the hot path and its callers."""
from .super_resolution import SuperResolutionNet, LightweightSuperResolution
from .frame_recovery import FrameRecoveryNet
from .enhancement_engine import EnhancementEngine, EnhancementConfig
from .layers import (DepthwiseSeparableConv, PixelShuffleUpsampler, CBAM, ChannelAttention, SpatialAttention,
                     LiteFlowNetCorrelation, ResidualBlock, TemporalConv3D, warp_features, warp_indices)

__all__ = ["SuperResolutionNet", "LightweightSuperResolution", "FrameRecoveryNet", "EnhancementEngine", "EnhancementConfig",
           "DepthwiseSeparableConv", "PixelShuffleUpsampler", "CBAM", "ChannelAttention", "SpatialAttention",
           "LiteFlowNetCorrelation", "ResidualBlock", "TemporalConv3D", "warp_features", "warp_indices"]
