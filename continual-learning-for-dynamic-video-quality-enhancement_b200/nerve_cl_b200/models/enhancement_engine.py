"""Drop-in ``EnhancementEngine`` (reference ``nerve_cl/models/enhancement_engine.py:40-308``): frame recovery ->
super-resolution -> strength blend, on the B200 kernel path.

Same ``EnhancementConfig`` fields, constructor order (so the RNG stream and ``state_dict`` keys ``frame_recovery.*`` /
``super_resolution.*`` / ``enhancement_strength`` match), ``forward`` result dict (``'recovered'`` /
``'super_resolved'`` / ``'enhanced'``), ``enhance_video``, ``get_model_info`` and ``set_enhancement_mode``.

Reference behaviours kept on purpose: the SR network is fed the RAW frames
``frames[:, c-w : c+w+1]``, not the recovered one (``enhancement_engine.py:146-148``); a window clipped by the clip
border is padded by repeating its LAST frame (``:152-158``), so near the leading edge the network's centre slot does
not hold frame ``t``; the strength blend mixes in the bicubic upsample of ``frames[:, center_idx]``.

What changed is the host side of the loop, which capped inference throughput regardless of kernel speed:

* no per-call host synchronisation: ``enhancement_strength.item()`` (``:170``) is cached against the parameter's
  version counter, and ``corruption_mask.sum() > 0`` (``:131``) is not evaluated -- with a mask given, recovery always
  runs; for an all-zero mask its blend returns the input frame exactly, so ``'enhanced'`` is unchanged and the only
  difference is that ``'recovered'`` is present (set ``engine.sync_free = False`` to get the reference's key set, at
  the price of one device->host sync per call);
* the strength blend is one in-place kernel (``nervecl::bicubic_blend``) on the inference path;
* ``enhance_video`` does not run one window per Python iteration: frames whose windows have the same shape (all
  interior frames) are gathered on the device and go through the networks ``batch_size`` windows per call.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops as _ops
from .frame_recovery import FrameRecoveryNet
from .super_resolution import LightweightSuperResolution, SuperResolutionNet

Tensor = torch.Tensor
nv = _ops.nv


@dataclass
class EnhancementConfig:
    """Reference ``EnhancementConfig`` (enhancement_engine.py:18-37), field for field."""
    frame_recovery_enabled: bool = True
    recovery_base_channels: int = 64
    recovery_temporal_window: int = 2
    super_resolution_enabled: bool = True
    scale_factor: int = 2
    sr_num_features: int = 64
    sr_num_residual_blocks: int = 8
    sr_temporal_window: int = 1
    use_lightweight_sr: bool = False
    enhancement_mode: str = "sequential"
    upscale_first: bool = False


class EnhancementEngine(nn.Module):
    def __init__(self, config: Optional[EnhancementConfig] = None):
        super().__init__()
        self.config = config or EnhancementConfig()
        c = self.config
        # same construction order as the reference (:66-92) => same RNG stream => identical initial weights
        self.frame_recovery = (FrameRecoveryNet(base_channels=c.recovery_base_channels, temporal_window=c.recovery_temporal_window)
                               if c.frame_recovery_enabled else None)
        if c.super_resolution_enabled:
            if c.use_lightweight_sr:
                self.super_resolution = LightweightSuperResolution(scale_factor=c.scale_factor)
            else:
                self.super_resolution = SuperResolutionNet(scale_factor=c.scale_factor, num_features=c.sr_num_features,
                                                           num_residual_blocks=c.sr_num_residual_blocks,
                                                           temporal_window=c.sr_temporal_window)
        else:
            self.super_resolution = None
        self.enhancement_strength = nn.Parameter(torch.ones(1))
        self.sync_free = True
        self._strength_cache: Tuple[int, float] = (-1, 1.0)

    # ------------------------------------------------------------------------------------------------------------
    def _strength(self) -> float:
        p = self.enhancement_strength
        if self._strength_cache[0] != p._version or self._strength_cache[0] < 0:
            self._strength_cache = (p._version, float(p.detach().item()))     # one sync per CHANGE of the parameter
        return self._strength_cache[1]

    def _apply(self, fn, *a, **k):                       # .to() / .cuda() replace the parameter: drop the cache
        self._strength_cache = (-1, 1.0)
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._strength_cache = (-1, 1.0)
        return super().load_state_dict(*a, **k)

    def forward(self, frames: Tensor, center_idx: Optional[int] = None, corruption_mask: Optional[Tensor] = None,
                enhancement_strength: Optional[float] = None) -> Dict[str, Tensor]:
        """frames (B,T,C,H,W) -> {'enhanced', ['recovered'], ['super_resolved']} (reference :95-184)."""
        if frames.dim() != 5:
            raise ValueError(f"expected (B, T, C, H, W) frames, got shape {tuple(frames.shape)}")
        B, T, C, H, W = frames.shape
        if center_idx is None:
            center_idx = T // 2
        results: Dict[str, Tensor] = {}
        current = frames[:, center_idx]
        ref_idx = [i for i in range(T) if i != center_idx]

        if self.frame_recovery is not None and corruption_mask is not None:
            if self.sync_free or bool(corruption_mask.sum() > 0):
                if not ref_idx:
                    raise ValueError("frame recovery needs at least one reference frame besides the centre frame")
                recovered = self.frame_recovery(current, frames[:, ref_idx], corruption_mask)
                results["recovered"] = recovered
                current = recovered

        if self.super_resolution is not None:
            w = self.config.sr_temporal_window
            s0, e0 = max(0, center_idx - w), min(T, center_idx + w + 1)
            sr_frames = frames[:, s0:e0]
            want = 2 * w + 1
            if sr_frames.shape[1] < want:                                    # pad by repeating the LAST frame (:152-158)
                sr_frames = torch.cat([sr_frames, sr_frames[:, -1:].expand(-1, want - sr_frames.shape[1], -1, -1, -1)], 1)
            if isinstance(self.super_resolution, LightweightSuperResolution):
                sr = self.super_resolution(current)
            else:
                sr = self.super_resolution(sr_frames)
            results["super_resolved"] = sr
            current = sr

        strength = enhancement_strength if enhancement_strength is not None else self._strength()
        if strength < 1.0 and "super_resolved" in results:
            lr = frames[:, center_idx].detach().float()
            if lr.stride(-1) != 1:
                lr = lr.contiguous()
            s = self.config.scale_factor
            if current.requires_grad:                                        # training: keep the SR node in the graph
                bic = torch.zeros_like(current)
                nv.bicubic_blend(bic, lr, s, 0.0)
                current = strength * current + (1.0 - strength) * bic
            else:
                current = current.clone()
                nv.bicubic_blend(current, lr, s, float(strength))
        results["enhanced"] = current
        return results

    # ------------------------------------------------------------------------------------------------------------
    def window_table(self, num_frames: int) -> List[Tuple[int, int, int]]:
        """(start, end, centre-in-window) of every output frame's window, as the reference loop (:214-228) cuts it."""
        half = (2 * max(self.config.recovery_temporal_window, self.config.sr_temporal_window) + 1) // 2
        rows = []
        for t in range(num_frames):
            start, end = max(0, t - half), min(num_frames, t + half + 1)
            rows.append((start, end, t - start))
        return rows

    def _video_plan(self, T: int, per_call: int, device) -> List[Tuple[int, Tensor, Tensor]]:
        """[(centre index, source-frame table (n, L), output frame ids (n,))] per network call for a T-frame clip, built once
        per (T, batch size, device) and kept on the device: the loop itself issues no host->device copies."""
        key = (T, per_call, str(device), self.config.recovery_temporal_window, self.config.sr_temporal_window)
        cache = self.__dict__.setdefault("_video_plans", {})
        if key not in cache:
            table = self.window_table(T)
            groups: Dict[Tuple[int, int], List[int]] = {}
            for t, (start, end, c) in enumerate(table):
                groups.setdefault((end - start, c), []).append(t)
            plan = []
            for (length, c), ts in groups.items():
                for i in range(0, len(ts), per_call):
                    chunk = ts[i:i + per_call]
                    idx = torch.tensor([[table[t][0] + j for j in range(length)] for t in chunk], device=device)
                    plan.append((c, idx, torch.tensor(chunk, device=device)))
            if len(cache) > 8:
                cache.clear()
            cache[key] = plan
        return cache[key]

    @torch.no_grad()
    def enhance_video(self, video: Tensor, corruption_masks: Optional[Tensor] = None, batch_size: int = 4) -> Tensor:
        """video (T,C,H,W) or (B,T,C,H,W) [+ masks (T,1,H,W)] -> enhanced video (reference :186-248).

        ``batch_size`` = windows per network call.  Frames whose windows have the same (length, centre) -- all
        interior frames -- are batched; every window still sees exactly the frames the reference loop gives it."""
        squeeze = video.dim() == 4
        if squeeze:
            video = video.unsqueeze(0)
        if video.dim() != 5:
            raise RuntimeError("enhance_video: video must be (T,C,H,W) or (B,T,C,H,W)")
        B, T, C, H, W = video.shape
        per_call = max(1, batch_size)
        out: Optional[Tensor] = None
        if self.super_resolution is not None and (self.frame_recovery is None or corruption_masks is None):
            return self._enhance_video_sr_only(video, per_call, squeeze)
        for c, idx, frames_t in self._video_plan(T, per_call, video.device):
            n, length = idx.shape
            win = video[:, idx]                                           # (B, n, L, C, H, W): one device gather
            win = win.transpose(0, 1).reshape(n * B, length, C, H, W)
            mask = None
            if corruption_masks is not None:
                m = corruption_masks[frames_t.to(corruption_masks.device)]                          # (n, 1, H, W)
                mask = m.unsqueeze(1).expand(-1, B, -1, -1, -1).reshape(n * B, *m.shape[1:])
            enh = self.forward(win, center_idx=c, corruption_mask=mask)["enhanced"]
            if out is None:
                out = torch.empty((B, T) + tuple(enh.shape[1:]), device=enh.device, dtype=enh.dtype)
            out[:, frames_t] = enh.view(n, B, *enh.shape[1:]).transpose(0, 1)
        return out.squeeze(0) if squeeze else out

    def _enhance_video_sr_only(self, video: Tensor, per_call: int, squeeze: bool) -> Tensor:
        """``enhance_video`` when no recovery runs (no recovery network, or no masks): every frame's result is the SR
        network on its SR window, and all those windows have the SAME shape -- the clipped ones are padded to
        ``2 * sr_temporal_window + 1`` frames (:152-158) -- so the windows of ALL frames, edge frames included, go through
        the network ``per_call`` at a time.  (Grouping by (window length, centre) as the general path must for the
        recovery network gives every edge frame a network call of its own: 4 extra batch-1 calls per clip at T = 5.)
        The window -> source-frame table is ``inference.sr_window_indices`` (pinned to the reference loop by
        ``tests/golden/enhance_windows.npz``)."""
        from ..inference import sr_window_indices
        B, T, C, H, W = video.shape
        light = isinstance(self.super_resolution, LightweightSuperResolution)
        key = ("sr_only", T, str(video.device), self.config.recovery_temporal_window, self.config.sr_temporal_window)
        cache = self.__dict__.setdefault("_video_plans", {})
        if key not in cache:
            if len(cache) > 8:
                cache.clear()
            cache[key] = sr_window_indices(T, self.config.sr_temporal_window,
                                           self.config.recovery_temporal_window).to(video.device)
        idx = cache[key]
        strength = self._strength()
        s = self.config.scale_factor
        out: Optional[Tensor] = None
        for t0 in range(0, T, per_call):
            n = min(per_call, T - t0)
            centre = video[:, t0:t0 + n].transpose(0, 1).reshape(n * B, C, H, W)      # the frames themselves (blend, lightweight)
            if light:
                sr = self.super_resolution(centre)
            else:
                win = video[:, idx[t0:t0 + n]]                                         # (B, n, L, C, H, W): one device gather
                sr = self.super_resolution(win.transpose(0, 1).reshape(n * B, idx.shape[1], C, H, W))
            if strength < 1.0:
                sr = sr.clone()
                _ops.nv.bicubic_blend(sr, centre.detach().float().contiguous(), s, float(strength))
            if out is None:
                out = torch.empty((B, T) + tuple(sr.shape[1:]), device=sr.device, dtype=sr.dtype)
            out[:, t0:t0 + n] = sr.view(n, B, *sr.shape[1:]).transpose(0, 1)
        return out.squeeze(0) if squeeze else out

    # ------------------------------------------------------------------------------------------------------------
    def get_model_info(self) -> Dict[str, Any]:
        """Reference :250-272."""
        info: Dict[str, Any] = {
            "config": {"frame_recovery_enabled": self.config.frame_recovery_enabled,
                       "super_resolution_enabled": self.config.super_resolution_enabled,
                       "scale_factor": self.config.scale_factor, "use_lightweight_sr": self.config.use_lightweight_sr},
            "parameters": {"total": sum(p.numel() for p in self.parameters()),
                           "trainable": sum(p.numel() for p in self.parameters() if p.requires_grad)},
        }
        if self.frame_recovery is not None:
            info["parameters"]["frame_recovery"] = self.frame_recovery.get_num_parameters()
        if self.super_resolution is not None:
            sr = self.super_resolution
            info["parameters"]["super_resolution"] = (sr.get_num_parameters() if hasattr(sr, "get_num_parameters")
                                                      else sum(p.numel() for p in sr.parameters()))
        return info

    def set_enhancement_mode(self, mode: str) -> None:
        """Reference :274-294 (only flips the config flags, like the reference: the sub-modules stay as built)."""
        c = self.config
        if mode == "full":
            c.frame_recovery_enabled, c.super_resolution_enabled = True, True
        elif mode == "recovery_only":
            c.frame_recovery_enabled, c.super_resolution_enabled = True, False
        elif mode == "sr_only":
            c.frame_recovery_enabled, c.super_resolution_enabled = False, True
        elif mode == "lightweight":
            c.frame_recovery_enabled, c.super_resolution_enabled, c.use_lightweight_sr = False, True, True
