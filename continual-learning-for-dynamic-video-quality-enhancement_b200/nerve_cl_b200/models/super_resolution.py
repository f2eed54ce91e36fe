"""Drop-in ``SuperResolutionNet`` backed by the sm_100a kernel library.

Same constructor, public attributes, methods, sub-module tree and ``state_dict`` keys as the reference
``nerve_cl/models/super_resolution.py:256-431`` (so checkpoints, ``EWC`` parameter names and
``train_baseline.py`` / ``train_continual.py`` keep working), but ``forward`` does not execute the
sub-modules: they only *hold* the parameters, created by the same torch constructors in the same order
as the reference, so ``torch.manual_seed(s); SuperResolutionNet(...)`` yields bit-identical initial
weights.  The arithmetic runs in ``engine.Plan`` through ``torch.ops.nervecl`` as one autograd node.

There is no CPU path: calling ``forward`` on CPU tensors raises ``RuntimeError``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import engine as _engine
from .. import ops as _ops
from .layers import CBAM, DepthwiseSeparableConv, LiteFlowNetCorrelation, PixelShuffleUpsampler

Tensor = torch.Tensor


def _conv_relu(cin: int, cout: int) -> List[nn.Module]:
    return [nn.Conv2d(cin, cout, 3, 1, 1), nn.ReLU(inplace=True)]


class FeatureExtractor(nn.Module):
    """Parameter holder for reference ``FeatureExtractor`` (super_resolution.py:22-54)."""

    def __init__(self, in_channels: int = 3, num_features: int = 64):
        super().__init__()
        self.head = nn.Sequential(*_conv_relu(in_channels, num_features))
        self.body = nn.Sequential(*[DepthwiseSeparableConv(num_features, num_features) for _ in range(3)])


class MotionEstimator(nn.Module):
    """Parameter holder for reference ``MotionEstimator`` (super_resolution.py:57-101)."""

    def __init__(self, in_channels: int = 64):
        super().__init__()
        self.correlation = LiteFlowNetCorrelation(max_displacement=4)
        widths = [(2 * 4 + 1) ** 2, 128, 64, 32]
        mods: List[nn.Module] = []
        for a, b in zip(widths[:-1], widths[1:]):
            mods += _conv_relu(a, b)
        mods.append(nn.Conv2d(widths[-1], 2, 3, 1, 1))
        self.flow_net = nn.Sequential(*mods)


class TemporalAggregator(nn.Module):
    """Parameter holder for reference ``TemporalAggregator`` (super_resolution.py:146-209)."""

    def __init__(self, num_features: int = 64, num_frames: int = 3):
        super().__init__()
        self.num_frames = num_frames
        self.attention = nn.Sequential(
            *_conv_relu(num_features * num_frames, num_features),
            *_conv_relu(num_features, num_features),
            nn.Conv2d(num_features, num_frames, 3, 1, 1),
            nn.Softmax(dim=1),
        )
        self.refine = CBAM(num_features)


class ResidualDenseBlock(nn.Module):
    """Parameter holder for reference ``ResidualDenseBlock`` (super_resolution.py:212-253)."""

    def __init__(self, num_features: int = 64, growth_rate: int = 32, num_layers: int = 5):
        super().__init__()
        if growth_rate != _engine.GROWTH or num_layers != _engine.RDB_LAYERS:
            raise ValueError("the kernel engine is specialised to growth_rate=32, num_layers=5 (reference defaults)")
        self.layers = nn.ModuleList(
            nn.Sequential(*_conv_relu(num_features + i * growth_rate, growth_rate)) for i in range(num_layers))
        self.lff = nn.Conv2d(num_features + num_layers * growth_rate, num_features, 1)


class _SRFunction(torch.autograd.Function):
    """forward + backward of the whole network as one autograd node."""

    @staticmethod
    def forward(ctx, module: "SuperResolutionNet", lr_frames: Tensor, *params: Tensor):
        names = module._param_names
        P = {n: p.detach() for n, p in zip(names, params)}
        BUF = {n: b for n, b in module.named_buffers()}
        # (needs_input_grad reflects requires_grad only; the grad mode of the CALLER is recorded by the module,
        #  because autograd runs this method with grad mode off)
        need_bwd = module._grad_mode and any(ctx.needs_input_grad[2:])
        plan = module._plan_for(lr_frames)
        B, T, C, H, W = lr_frames.shape
        s = module.scale_factor
        out = torch.empty((B, C, H * s, W * s), device=lr_frames.device, dtype=torch.float32)
        acts = plan.forward(lr_frames.detach(), P, BUF, module.training, need_bwd, out)
        module._last_acts = acts if module._keep_intermediate else None
        if need_bwd:
            ctx.plan, ctx.acts, ctx.module, ctx.P = plan, acts, module, P
        elif not module._keep_intermediate:
            plan.release(acts)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dout: Optional[Tensor]):
        module, plan, acts, P = ctx.module, ctx.plan, ctx.acts, ctx.P
        names = module._param_names
        if acts is None:
            raise RuntimeError("nerve_cl_b200: backward called twice on the same SuperResolutionNet forward")
        if dout is None:
            return (None, None) + (None,) * len(names)
        # a fresh flat fp32 gradient buffer per backward: autograd may adopt ("steal") the returned
        # views as param.grad, so it must never be recycled by the engine.
        flat = torch.empty(module._flat_numel, device=dout.device, dtype=torch.float32)
        _ops.nv.fill_zero(flat)
        G = {n: flat[o:o + k].view(shape) for n, (o, k, shape) in module._flat_layout.items()}
        hook = module._grad_sync.on_ready if module._grad_sync is not None else None
        if module._grad_sync is not None:
            module._grad_sync.begin(flat, module._flat_layout)
        plan.backward(acts, dout.contiguous().float(), P, G, hook)
        if module._grad_sync is not None:
            module._grad_sync.finish()
        module._last_flat_grad = flat
        ctx.acts = None
        plan.release(acts)
        return (None, None) + tuple(G[n] for n in names)


class SuperResolutionNet(nn.Module):
    """Lightweight temporal super-resolution network (reference super_resolution.py:256).

    Args (unchanged from the reference):
        in_channels: image channels (the kernels are specialised to 3)
        scale_factor: 2, 3 or 4
        num_features: feature channels F (multiple of 8; 8 <= F <= 256, power-of-two multiple of 8)
        num_residual_blocks: number of residual dense blocks
        temporal_window: reference frames on each side; ``num_frames = 2*temporal_window + 1``

    Extra, non-reference knobs (attributes, not constructor arguments):
        ``compute_dtype``: ``None`` (default: bf16 under ``torch.autocast('cuda', torch.bfloat16)``,
        else fp32), ``torch.float32`` or ``torch.bfloat16``.
        ``conv_engine``: ``ops.CONV_AUTO`` / ``CONV_SIMT`` / ``CONV_TC``.
        ``warp_div_mode``: 0 replays ATen-CUDA's ``x * (1/(W-1))`` (default), 1 ATen-CPU's ``x / (W-1)``.
    """

    def __init__(self, in_channels: int = 3, scale_factor: int = 2, num_features: int = 64,
                 num_residual_blocks: int = 8, temporal_window: int = 1):
        super().__init__()
        self.scale_factor = scale_factor
        self.temporal_window = temporal_window
        self.num_frames = 2 * temporal_window + 1
        self.in_channels = in_channels
        self.num_features = num_features
        self.num_residual_blocks = num_residual_blocks

        # same construction order as the reference => same RNG stream => identical initial weights
        self.feature_extractor = FeatureExtractor(in_channels, num_features)
        self.motion_estimator = MotionEstimator(num_features)
        self.temporal_aggregator = TemporalAggregator(num_features, self.num_frames)
        self.residual_blocks = nn.Sequential(*[ResidualDenseBlock(num_features) for _ in range(num_residual_blocks)])
        self.gff = nn.Sequential(*_conv_relu(num_features, num_features))
        self.upsampler = PixelShuffleUpsampler(in_channels=num_features, scale_factor=scale_factor,
                                               out_channels=in_channels)
        self.bicubic_upsample = nn.Upsample(scale_factor=scale_factor, mode="bicubic", align_corners=False)

        self.compute_dtype: Optional[torch.dtype] = None
        self.conv_engine = _ops.CONV_AUTO
        self.warp_div_mode = 0
        self._plans: Dict[Tuple, _engine.Plan] = {}
        self._keep_intermediate = False
        self._grad_mode = True
        self._last_acts = None
        self._last_flat_grad: Optional[Tensor] = None
        self._grad_sync = None
        self._param_names: List[str] = [n for n, _ in self.named_parameters()]
        off = 0
        self._flat_layout: Dict[str, Tuple[int, int, torch.Size]] = {}
        for n, p in self.named_parameters():
            self._flat_layout[n] = (off, p.numel(), p.shape)
            off += (p.numel() + 3) // 4 * 4          # keep every tensor 16-byte aligned in the flat buffer
        self._flat_numel = off

    # ------------------------------------------------------------------------------------
    def _dtype_now(self) -> torch.dtype:
        if self.compute_dtype is not None:
            return self.compute_dtype
        if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            return torch.bfloat16
        return torch.float32

    def _plan_for(self, lr_frames: Tensor) -> _engine.Plan:
        B, T, C, H, W = lr_frames.shape
        adt = self._dtype_now()
        key = (B, T, H, W, adt, lr_frames.device)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 4:                      # bound workspace growth across shape changes
                self._plans.pop(next(iter(self._plans)))
            R = self.temporal_aggregator.refine.channel_attention.fc[0].out_features
            plan = _engine.Plan(self.num_features, self.num_residual_blocks, self.scale_factor, R, B, T, H, W, adt,
                                lr_frames.device)
            self._plans[key] = plan
        plan.engine = self.conv_engine
        plan.div_mode = self.warp_div_mode
        return plan

    def _validate(self, lr_frames: Tensor) -> None:
        if lr_frames.dim() != 5:
            raise ValueError(f"expected (B, T, C, H, W) frames, got shape {tuple(lr_frames.shape)}")
        if not lr_frames.is_cuda:
            raise RuntimeError("nerve_cl_b200.SuperResolutionNet runs on CUDA (sm_100a) only; there is no CPU "
                               "fallback. Move the module and its inputs to a B200.")
        p0 = next(self.parameters())
        if p0.device != lr_frames.device:
            raise RuntimeError(f"module parameters are on {p0.device} but the frames are on {lr_frames.device}")
        B, T, C, H, W = lr_frames.shape
        if C != 3 or self.in_channels != 3:
            raise ValueError("the kernel engine is specialised to in_channels=3")
        if T != self.num_frames:
            # the reference fails here too: attention.0 expects num_frames*F input channels (:167)
            raise RuntimeError(f"expected {self.num_frames} frames (temporal_window={self.temporal_window}), got {T}")
        if H < 2 or W < 2:
            raise ValueError("H and W must be >= 2 (the reference divides by (W-1), super_resolution.py:129-133)")
        F = self.num_features
        if F % 8 or F > 256 or ((F // 8) & (F // 8 - 1)):
            raise ValueError("num_features must be 8, 16, 32, 64, 128 or 256 for the kernel engine")
        if lr_frames.requires_grad:
            raise NotImplementedError("gradients w.r.t. the input frames are not produced by the kernel engine")

    def forward(self, lr_frames: Tensor, return_intermediate: bool = False):
        """Upscale the centre frame of a (B, T, C, H, W) window -> (B, C, H*s, W*s) in [0, 1]."""
        self._validate(lr_frames)
        if lr_frames.dtype != torch.float32:
            lr_frames = lr_frames.float()
        self._keep_intermediate = return_intermediate
        self._grad_mode = torch.is_grad_enabled()
        try:
            out = _SRFunction.apply(self, lr_frames, *self.parameters())
            if not return_intermediate:
                return out
            return out, self._export_intermediate(lr_frames)
        finally:
            self._keep_intermediate = False
            self._last_acts = None

    def _export_intermediate(self, lr_frames: Tensor) -> Dict[str, object]:
        """NCHW fp32 copies of the tensors the reference returns at super_resolution.py:384-389."""
        A = self._last_acts
        B, T, C, H, W = lr_frames.shape
        F = self.num_features
        nv = _ops.nv

        def nchw(view: Tensor) -> Tensor:
            dst = torch.empty((view.shape[0], view.shape[3], H, W), device=view.device, dtype=torch.float32)
            nv.nhwc_to_nchw(view, dst)
            return dst

        feat = A.feat.view(T, B, H, W, F)
        feats = [nchw(feat[t]) for t in range(T)]
        aligned = [nchw(A.cat[..., t * F:(t + 1) * F]) for t in range(T)]
        agg_view = A.rdb[0][..., :F] if self.num_residual_blocks > 0 else A.trunk
        return {"features": feats, "aligned": aligned, "aggregated": nchw(agg_view),
                "flows": {t: A.flow[t].permute(0, 3, 1, 2).contiguous() for t in A.flow}}

    def forward_single(self, lr_frame: Tensor) -> Tensor:
        """Upscale one frame by replicating it ``num_frames`` times (super_resolution.py:393-405)."""
        return self.forward(lr_frame.unsqueeze(1).expand(-1, self.num_frames, -1, -1, -1))

    def get_num_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def get_flops(self, input_size: Tuple[int, int] = (128, 128)) -> int:
        """The reference's rough estimate (super_resolution.py:411-431), reproduced for API parity.

        It hard-codes F=64 / 8 blocks and ignores flow_net, aggregator and dense growth; use
        ``conv_macs_per_pixel`` for the exact multiply-accumulate count."""
        h, w = input_size
        f, c, s = 64, 3, self.scale_factor
        return h * w * (c * f * 9 + f * 81 * (self.num_frames - 1) + f * f * 9 * 8 + f * (c * s * s) * 9)

    def conv_macs_per_pixel(self) -> int:
        """Exact forward conv MACs per LR pixel per sample (closed form of SURVEY.md section 8a)."""
        F, T, s, NB = self.num_features, self.num_frames, self.scale_factor, self.num_residual_blocks
        fe = 27 * F + 3 * (9 * F + F * F)
        flow = 9 * (81 * 128 + 128 * 64 + 64 * 32 + 32 * 2)
        agg = 9 * (T * F * F + F * F + F * T) + 98
        rdb = sum(9 * (F + 32 * i) * 32 for i in range(5)) + (F + 160) * F
        return T * fe + (T - 1) * flow + agg + NB * rdb + 9 * F * F + 9 * F * 3 * s * s

    # ------------------------------------------------------------------------------------
    def last_flat_grad(self) -> Optional[Tensor]:
        """The flat fp32 gradient buffer of the most recent backward (parameter order, 16-byte aligned
        slots) if every ``param.grad`` still aliases it -- lets optimisers / EWC run one fused kernel."""
        flat = self._last_flat_grad
        if flat is None:
            return None
        for n, p in self.named_parameters():
            o, k, _ = self._flat_layout[n]
            if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + 4 * o:
                return None
        return flat

    def set_gradient_sync(self, sync) -> None:
        """Install a ``distributed.GradSync`` (bucketed all-reduce overlapped with backward)."""
        self._grad_sync = sync


class _LightFunction(torch.autograd.Function):
    """forward + backward of LightweightSuperResolution as one autograd node (same scheme as ``_SRFunction``)."""

    @staticmethod
    def forward(ctx, module: "LightweightSuperResolution", x: Tensor, *params: Tensor):
        names = module._param_names
        P = {n: p.detach() for n, p in zip(names, params)}
        need_bwd = module._grad_mode and any(ctx.needs_input_grad[2:])
        out, saved = module._run_forward(x.detach(), P, need_bwd)
        if need_bwd:
            ctx.module, ctx.P, ctx.saved = module, P, saved
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dout: Optional[Tensor]):
        module, P, saved = ctx.module, ctx.P, ctx.saved
        names = module._param_names
        if dout is None:
            return (None, None) + (None,) * len(names)
        G = {n: torch.zeros_like(P[n]) for n in names}
        module._run_backward(saved, dout.contiguous().float(), P, G)
        ctx.saved = None
        return (None, None) + tuple(G[n] for n in names)


class LightweightSuperResolution(nn.Module):
    """Ultra-light single-frame SR (reference super_resolution.py:434-470): conv 3->32 + ReLU, four depthwise-separable
    blocks (dw 3x3 -> pw 1x1 -> BatchNorm -> ReLU), conv 32 -> 3*s^2, PixelShuffle, + bicubic(x), clamp(0, 1).

    Same constructor / ``net`` Sequential / ``state_dict`` keys as the reference; the arithmetic runs on the kernels the
    SuperResolutionNet extractor and output stage use (9 868 parameters: every layer is bandwidth-bound): the 3-channel
    head as a 1x1 over 3x3-unfolded pixels, TMA-free generic depthwise / BatchNorm kernels at 32 channels, pointwise
    convs through ``conv2d_fwd`` (BatchNorm folded into them on the pure-inference path), PixelShuffle + bicubic skip +
    clamp in ``upfinish``.  Training (batch-statistics BatchNorm, running-stat updates, full backward) is supported."""

    def __init__(self, scale_factor: int = 2):
        super().__init__()
        self.scale_factor = scale_factor
        self.net = nn.Sequential(
            nn.Conv2d(3, 32, 3, 1, 1), nn.ReLU(inplace=True),
            *[DepthwiseSeparableConv(32, 32) for _ in range(4)],
            nn.Conv2d(32, 3 * scale_factor ** 2, 3, 1, 1), nn.PixelShuffle(scale_factor))
        self.bicubic = nn.Upsample(scale_factor=scale_factor, mode="bicubic", align_corners=False)
        self.compute_dtype: Optional[torch.dtype] = None
        self.conv_engine = _ops.CONV_AUTO
        self._grad_mode = True
        self._param_names: List[str] = [n for n, _ in self.named_parameters()]

    def get_num_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def _dtype_now(self) -> torch.dtype:
        if self.compute_dtype is not None:
            return self.compute_dtype
        if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            return torch.bfloat16
        return torch.float32

    def forward(self, x: Tensor) -> Tensor:
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected (B, 3, H, W) frames, got shape {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("nerve_cl_b200.LightweightSuperResolution runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.requires_grad:
            raise NotImplementedError("gradients w.r.t. the input frames are not produced by the kernel engine")
        self._grad_mode = torch.is_grad_enabled()
        return _LightFunction.apply(self, x.float(), *self.parameters())

    # ---- engine -----------------------------------------------------------------------------------------------
    C, BN_EPS, BN_MOMENTUM = 32, 1e-5, 0.1

    def _pack(self, w: Tensor, adt, flip: bool, rows: int, cols: int) -> Tensor:
        k = w.shape[-1]
        dst = torch.empty((k * k, rows, (cols + 7) // 8 * 8), device=w.device, dtype=adt)
        _ops.nv.pack_conv_weight(w.contiguous().float(), dst, flip)
        return dst

    def _run_forward(self, x: Tensor, P: Dict[str, Tensor], need_bwd: bool):
        nv, eng = _ops.nv, self.conv_engine
        B, _, H, W = x.shape
        C, s, adt, dev, f32 = self.C, self.scale_factor, self._dtype_now(), x.device, torch.float32
        training = self.training
        fold = (not training) and (not need_bwd)
        BUF = dict(self.named_buffers())

        def act(c, dtype=adt):
            return torch.empty((B, H, W, c), device=dev, dtype=dtype)

        S: Dict[str, object] = {"x": x, "training": training, "adt": adt}
        x_in = act(32)
        nv.pack_frames_unfold3(x.unsqueeze(1), x_in)                       # (B,1,3,H,W): one "frame" per sample
        head = act(C)
        w0 = P["net.0.weight"]
        nv.conv2d_fwd(x_in, self._pack(w0.view(C, 27, 1, 1), adt, False, C, 32), P["net.0.bias"], None, None, None, head, C,
                      1, False, 0, 0, 1.0, eng)
        S["x_in"], S["head"] = x_in, head
        cur = head
        S["dwo"], S["pwo"], S["act"], S["stat"] = [], [], [], []
        for j in range(4):
            pre = f"net.{2 + j}."
            dwo = act(C)
            nv.dwconv3x3_fwd(cur, P[pre + "depthwise.weight"], dwo, False, False)
            y = act(C)
            if fold:
                sc = P[pre + "bn.weight"] * torch.rsqrt(BUF[pre + "bn.running_var"] + self.BN_EPS)
                bias = (P[pre + "bn.bias"] - BUF[pre + "bn.running_mean"] * sc).contiguous()
                wf = self._pack(P[pre + "pointwise.weight"] * sc.view(-1, 1, 1, 1), adt, False, C, C)
                nv.conv2d_fwd(dwo, wf, bias, None, None, None, y, C, 1, False, 0, 0, 1.0, eng)
            else:
                pwo = act(C)
                nv.conv2d_fwd(dwo, self._pack(P[pre + "pointwise.weight"], adt, False, C, C), None, None, None, None, pwo, C,
                              0, False, 0, 0, 1.0, eng)
                stat = torch.empty((1, C, 2), device=dev, dtype=f32)
                sums = None
                if training:
                    sums = torch.zeros((1, C, 2), device=dev, dtype=torch.float64)
                    nv.bn_stats(pwo, 1, sums)
                nv.bn_finalize(sums, stat, BUF[pre + "bn.running_mean"], BUF[pre + "bn.running_var"],
                               BUF[pre + "bn.num_batches_tracked"], B * H * W, 1, self.BN_MOMENTUM, self.BN_EPS, training)
                nv.bn_relu_fwd(pwo, stat, P[pre + "bn.weight"], P[pre + "bn.bias"], None, y, 1)
                S["pwo"].append(pwo); S["stat"].append(stat)
            S["dwo"].append(dwo); S["act"].append(y)
            cur = y
        ncs = 3 * s * s
        up = act(ncs, f32)
        nv.conv2d_fwd(cur, self._pack(P["net.6.weight"], adt, False, ncs, C), P["net.6.bias"], None, None, None, up, ncs, 0,
                      False, 0, 0, 1.0, eng)
        out = torch.empty((B, 3, H * s, W * s), device=dev, dtype=f32)
        nv.upfinish_fwd(up, x, out, s)
        S["up"] = up
        return out, (S if need_bwd else None)

    def _run_backward(self, S, dout: Tensor, P: Dict[str, Tensor], G: Dict[str, Tensor]) -> None:
        nv, eng = _ops.nv, self.conv_engine
        x, adt, training = S["x"], S["adt"], S["training"]
        B, _, H, W = x.shape
        C, s, dev, f32 = self.C, self.scale_factor, x.device, torch.float32
        ncs = 3 * s * s

        def act(c, dtype=adt):
            return torch.empty((B, H, W, c), device=dev, dtype=dtype)

        dup = act(ncs, f32)
        nv.upfinish_bwd(S["up"], x, dout, dup, s)
        dup_a = torch.zeros((B, H, W, (ncs + 15) // 16 * 16), device=dev, dtype=adt)
        nv.axpy(dup, dup_a[..., :ncs], 1.0, False)
        last = S["act"][3]
        nv.conv2d_wgrad(last, dup_a[..., :ncs], G["net.6.weight"], G["net.6.bias"], 1.0, eng)
        dy = act(C)
        wb = self._pack(P["net.6.weight"], adt, True, C, dup_a.shape[-1])
        nv.conv2d_fwd(dup_a, wb, None, None, None, None, dy, C, 0, False, 0, 0, 1.0, eng)
        for j in reversed(range(4)):
            pre = f"net.{2 + j}."
            gamma, beta = P[pre + "bn.weight"], P[pre + "bn.bias"]
            bsums = torch.zeros((1, C, 2), device=dev, dtype=torch.float64)
            nv.bn_relu_bwd_reduce(S["pwo"][j], dy, S["stat"][j], gamma, beta, 1, bsums)
            s0 = act(C)
            nv.bn_relu_bwd_apply(S["pwo"][j], dy, S["stat"][j], gamma, beta, bsums, s0, G[pre + "bn.weight"],
                                 G[pre + "bn.bias"], 1, training)
            nv.conv2d_wgrad(S["dwo"][j], s0, G[pre + "pointwise.weight"], None, 1.0, eng)
            s1 = act(C)
            nv.conv2d_fwd(s0, self._pack(P[pre + "pointwise.weight"], adt, True, C, C), None, None, None, None, s1, C, 0,
                          False, 0, 0, 1.0, eng)
            x_in = S["act"][j - 1] if j > 0 else S["head"]
            nv.dwconv3x3_wgrad(x_in, s1, G[pre + "depthwise.weight"])
            dy = act(C)
            nv.dwconv3x3_fwd(s1, P[pre + "depthwise.weight"], dy, True, False)
        t0 = act(C)
        nv.relu_bwd(dy, S["head"], None, t0)
        gw = G["net.0.weight"]
        nv.conv2d_wgrad(S["x_in"][..., :27], t0, gw.view(C, 27, 1, 1), G["net.0.bias"], 1.0, eng)
