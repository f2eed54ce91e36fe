"""Execution engine for the SuperResolutionNet hot path.

Host-side orchestration (Python, as in the reference) of the ``torch.ops.nervecl`` kernels: one
forward pass is ~60 launches, one backward ~150, versus 1 036 / 2 091 ATen ops in the reference
(SURVEY.md section 2.1).  The engine owns every intermediate buffer:

* activations are NHWC in the compute dtype (fp32 parity path or bf16), flow / logits / attention /
  BN statistics / upsampler output are fp32;
* the T frames go through the shared feature extractor as one batch of T*B images whose BatchNorm
  statistics are reduced per frame group (the reference calls the extractor T times,
  super_resolution.py:346-349, so each call has its own batch statistics and running-stat update);
* ``torch.stack``/``view`` of the aligned features (super_resolution.py:194-197) and the ``torch.cat``
  chains of the residual dense blocks (:246-252) do not exist: producers write channel slices of one
  pitched buffer (T*F channels for the aggregator, F+5*32 for each dense block);
* backward mirrors that: one (F+5*32)-channel gradient buffer per dense block into which each layer's
  data gradient is accumulated by the convolution epilogue, with the ReLU mask of the slice that just
  became final applied in the same epilogue.

The whole forward+backward is exposed to autograd as ONE ``torch.autograd.Function`` (see
``models/super_resolution.py``) whose inputs are the frames and the 131 parameter tensors.
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from ._lib import CONV_AUTO, CONV_SIMT

Tensor = torch.Tensor


class _OpsProxy:
    """``torch.ops.nervecl`` with optional per-op CUDA-event timing (set ``.timer`` to a KernelTimer)."""

    _SPANNED_ELSEWHERE = {"conv2d_fwd", "conv2d_wgrad", "conv3x3_wgrad_grouped"}      # timed with FLOP counts by Plan._span

    def __init__(self):
        self.timer = None
        self._raw = ops.nv

    def __getattr__(self, name):
        fn = getattr(self._raw, name)
        if name in self._SPANNED_ELSEWHERE:
            return fn

        def call(*a, **k):
            t = self.timer
            if t is None:
                return fn(*a, **k)
            with t.span(name):
                return fn(*a, **k)
        return call


nv = _OpsProxy()

GROWTH = 32      # ResidualDenseBlock growth rate  (super_resolution.py:215)
RDB_LAYERS = 5   # (super_resolution.py:216)
CORR_CH = 81     # (2*4+1)^2 displacements          (super_resolution.py:73)
HEAD_UNFOLD = 32 # unfolded 3x3 RGB neighbourhood: 27 live channels padded to two 16-channel MMA k-steps
CORR_PAD = 96    # correlation volume is stored with 96 channels (zeros beyond 81): 16-byte aligned rows
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


class KernelTimer:
    """CUDA-event timing of individual launches on the launching stream (bench.py roofline numbers).

    ``with timer.span(kind, flops, bytes)`` brackets one op; ``summary()`` synchronises and returns
    kind -> {launches, ms, flops, bytes}."""

    def __init__(self):
        self.records: List[Tuple[str, float, float, torch.cuda.Event, torch.cuda.Event]] = []

    class _Span:
        def __init__(self, timer, kind, flops, nbytes):
            self.t, self.kind, self.flops, self.nbytes = timer, kind, flops, nbytes

        def __enter__(self):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

        def __exit__(self, *exc):
            self.e1.record()
            self.t.records.append((self.kind, self.flops, self.nbytes, self.e0, self.e1))

    def span(self, kind: str, flops: float = 0.0, nbytes: float = 0.0):
        return KernelTimer._Span(self, kind, flops, nbytes)

    def summary(self) -> Dict[str, Dict[str, float]]:
        torch.cuda.synchronize()
        out: Dict[str, Dict[str, float]] = {}
        for kind, flops, nbytes, e0, e1 in self.records:
            d = out.setdefault(kind, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


class _NoSpan:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NOSPAN = _NoSpan()


class ConvSpec:
    """One dense convolution of the network: where its parameters live and how they are packed."""

    def __init__(self, name: str, cin: int, cout: int, k: int, bias: bool, cin_pad: Optional[int] = None):
        self.name, self.cin, self.cout, self.k, self.has_bias = name, cin, cout, k, bias
        self.cin_pad = cin_pad or cin          # input buffer channels (corr volume is padded)


def _align(v: int, a: int) -> int:
    return (v + a - 1) // a * a


class Activations:
    """All buffers of one forward pass (kept until its backward has run)."""

    def __init__(self, plan: "Plan"):
        p, dev, adt = plan, plan.device, plan.adt
        B, T, H, W, F, s = p.B, p.T, p.H, p.W, p.F, p.scale
        f32 = torch.float32

        def act(n, c, dtype=adt):
            return torch.empty((n, H, W, c), device=dev, dtype=dtype)

        self.x_in = act(T * B, HEAD_UNFOLD)    # RGB 3x3 neighbourhoods unfolded into 27 (+5 zero) channels: the head conv is 1x1
        self.head = act(T * B, F)
        self.dwo = [act(T * B, F) for _ in range(3)]
        self.pwo = [act(T * B, F) for _ in range(3)]
        self.fact = [act(T * B, F) for _ in range(2)]
        self.feat = act(T * B, F)
        self.bn_sums = [torch.zeros((T, F, 2), device=dev, dtype=torch.float64) for _ in range(3)]
        self.bn_stat = [torch.empty((T, F, 2), device=dev, dtype=f32) for _ in range(3)]
        self.cat = act(B, T * F)
        self.corr = {t: act(B, CORR_PAD) for t in p.others}
        self.fn1 = {t: act(B, 128) for t in p.others}
        self.fn2 = {t: act(B, 64) for t in p.others}
        self.fn3 = {t: act(B, 32) for t in p.others}
        self.flow = {t: act(B, 2, f32) for t in p.others}
        self.a1, self.a2 = act(B, F), act(B, F)
        self.logits, self.attn = act(B, T, f32), act(B, T, f32)
        self.blend = act(B, F)
        self.pool = torch.zeros((B, F), device=dev, dtype=f32)
        self.hidden = torch.empty((B, p.R), device=dev, dtype=f32)
        self.gate = torch.empty((B, F), device=dev, dtype=f32)
        self.stats = act(B, 2, f32)
        self.sgate = torch.empty((B, H, W), device=dev, dtype=f32)
        # dense-block buffers: F + 5*32 channels at a pixel pitch rounded up to 64 channels (128 B), so that
        # every 64-channel TMA box row of the conv kernels is one aligned 128-byte line
        ct = F + RDB_LAYERS * GROWTH
        self.rdb = [act(B, _align(ct, 64))[..., :ct] for _ in range(p.NB)]
        self.trunk = act(B, F)
        self.fused = act(B, F)
        self.up = act(B, 3 * s * s, f32)
        self.lr_centre: Optional[Tensor] = None   # view of the caller's frames (kept for backward)
        self.training = False
        # packed ReLU signs of the dense-block layers (int16 per pixel and 16 channels), written by the forward convs and
        # read back as the masks of the fused block backward; allocated on first use
        self.rdb_bits: Optional[List[List[Tensor]]] = None
        self.has_bits = False
        self.sbits: Dict[object, Tensor] = {}     # the same for flow_net / attention activations, by (name, frame)
        self.use_bits = False


class Plan:
    """Shape-specialised buffer plan + op sequence for one (B,T,H,W,dtype) configuration."""

    def __init__(self, F: int, NB: int, scale: int, R: int, B: int, T: int, H: int, W: int, adt: torch.dtype,
                 device):
        self.B, self.T, self.H, self.W = B, T, H, W
        self.F, self.NB, self.scale, self.R = F, NB, scale, R
        self.adt, self.device = adt, device
        self.mid = T // 2
        self.others = [t for t in range(T) if t != self.mid]
        self.engine = CONV_AUTO
        self.div_mode = 0
        self.fold_bn = True               # fold running BatchNorm into the pointwise convs on the pure-inference path
        self.timer: Optional[KernelTimer] = None
        self._free: List[Activations] = []
        self._bwd_ws = None
        F = self.F
        # dense convolutions in registration order of the reference module
        cs: Dict[str, ConvSpec] = {}

        def add(name, cin, cout, k, bias=True, cin_pad=None):
            cs[name] = ConvSpec(name, cin, cout, k, bias, cin_pad)

        # the 3-channel 3x3 head runs as a 1x1 conv over the unfolded frames (nervecl_pack_frames_unfold3)
        add("feature_extractor.head.0", 27, F, 1, cin_pad=HEAD_UNFOLD)
        for j in range(3):
            add(f"feature_extractor.body.{j}.pointwise", F, F, 1, bias=False)
        add("motion_estimator.flow_net.0", CORR_CH, 128, 3, cin_pad=CORR_PAD)
        add("motion_estimator.flow_net.2", 128, 64, 3)
        add("motion_estimator.flow_net.4", 64, 32, 3)
        add("motion_estimator.flow_net.6", 32, 2, 3)
        add("temporal_aggregator.attention.0", T * F, F, 3)
        add("temporal_aggregator.attention.2", F, F, 3)
        add("temporal_aggregator.attention.4", F, T, 3)
        for k in range(self.NB):
            for i in range(RDB_LAYERS):
                add(f"residual_blocks.{k}.layers.{i}.0", F + i * GROWTH, GROWTH, 3)
            add(f"residual_blocks.{k}.lff", F + RDB_LAYERS * GROWTH, F, 1)
        add("gff.0", F, F, 3)
        add("upsampler.conv", F, 3 * self.scale ** 2, 3)
        self.convs = cs
        # packed weights: forward operator [K*K, Cout, Cin_pad8]; data-gradient operator [K*K, Cin_pad, Cout_pad8]
        self.wf: Dict[str, Tensor] = {}
        self.wb: Dict[str, Tensor] = {}
        for name, c in cs.items():
            kk = c.k * c.k
            self.wf[name] = torch.empty((kk, c.cout, _align(c.cin_pad, 8)), device=device, dtype=adt)
            if name != "feature_extractor.head.0":   # the input frames need no gradient
                # (>= 16 columns: the output gradients of the 2/3/12-channel convs are 16-channel zero-padded buffers)
                self.wb[name] = torch.empty((kk, c.cin_pad, _align(c.cout, 16)), device=device, dtype=adt)

        # Fused dense-block data gradient (bf16 tcgen05 path): the gradient of buffer slice s is ONE conv over the
        # already-final gradients of all later layers (adjacent channels of the gradient buffer) plus the LFF
        # 1x1 branch as a centre-tap-only second input.  wslice[s][k]: packed [9, rows, cols] per block k with
        # cols = [later layers' 32-ch groups, padded to 64 | 64 LFF channels]; s = 0 is the block input x.
        self.fused_rdb_bwd = False
        self.wslice: List[Tensor] = []
        if adt == torch.bfloat16 and W >= 64 and H >= 3 and self.NB > 0:
            self.fused_rdb_bwd = True
            for s in range(RDB_LAYERS):                       # s = 0: x slice (F channels); s >= 1: layer s-1's output
                rows = F if s == 0 else GROWTH
                cmain = GROWTH * (RDB_LAYERS - s) if s > 0 else GROWTH * RDB_LAYERS
                cols = _align(cmain, 64) + F
                self.wslice.append(torch.zeros((9, self.NB * rows, cols), device=device, dtype=adt))
            # The last slice's gradient is a 1x1 conv (only the LFF reaches it).  With sign-bit masks it runs as a 3x3 conv
            # whose eight outer taps are zero, because only the CTA-pair row kernel (3x3) reads the packed masks.
            self.wlast = torch.zeros((self.NB, 9, GROWTH, _align(F, 16)), device=device, dtype=adt)
        # ReLU masks of the dense-block layers as packed sign bits (nervecl_conv_params.sign_bits): 4 bytes per pixel and
        # layer instead of 64 for every mask read of the block backward.  Switched off for the plan the first time the
        # library answers "unsupported" (shapes too small for the CTA-pair kernel).
        self.sign_bits = self.fused_rdb_bwd
        self._bits_off = set()            # activations whose conv shape the sign-bit path does not take

    # ---- activation-set pool ---------------------------------------------------------------
    def acquire(self) -> Activations:
        return self._free.pop() if self._free else Activations(self)

    def release(self, acts: Activations) -> None:
        acts.lr_centre = None
        if len(self._free) < 2:
            self._free.append(acts)

    # ---- weights ---------------------------------------------------------------------------
    def pack_weights(self, P: Dict[str, Tensor], need_bwd: bool, fold_bn: Optional[Dict[str, Tensor]] = None) -> None:
        """``fold_bn``: pointwise conv name -> per-output-channel BatchNorm scale to fold into its packed weights
        (pure inference, see ``forward``)."""
        ws, ds, fl = [], [], []
        for name in self.convs:
            w = P[name + ".weight"]
            if name == "feature_extractor.head.0":
                w = w.view(w.shape[0], -1, 1, 1)           # OIHW flattening == unfolded channel order c*9 + tap
            if fold_bn is not None and name in fold_bn:
                w = w * fold_bn[name].view(-1, 1, 1, 1)
            ws.append(w); ds.append(self.wf[name]); fl.append(False)
            if need_bwd and name in self.wb:
                ws.append(w); ds.append(self.wb[name]); fl.append(True)
        nv.pack_conv_weights_batched(ws, ds, fl)

    def pack_rdb_slice_weights(self, P: Dict[str, Tensor]) -> None:
        """Assemble (in fp32, batched over the blocks) and pack the slice-gradient operators described in
        ``__init__``:  W_s[c, (i,o), tap] = W_i[o, c_lo+c, 8-tap] for the later layers i, and for the LFF columns
        the centre tap 0.2 * W_lff[d, c_lo+c].  The later layers' column groups are in DESCENDING layer order, the
        order of their gradients in the gradient buffer (``_rdb_backward_fused``)."""
        F, NB = self.F, self.NB
        Wl = [torch.stack([P[f"residual_blocks.{k}.layers.{i}.0.weight"] for k in range(NB)]) for i in range(RDB_LAYERS)]
        Wf = torch.stack([P[f"residual_blocks.{k}.lff.weight"] for k in range(NB)])[..., 0, 0]      # [NB, F, CT]
        for s in range(RDB_LAYERS):
            rows = F if s == 0 else GROWTH
            c_lo = 0 if s == 0 else F + (s - 1) * GROWTH
            first = 0 if s == 0 else s                      # first later layer
            cmain = GROWTH * (RDB_LAYERS - first)
            cols = _align(cmain, 64) + F
            comb = torch.zeros((NB, rows, cols, 3, 3), device=self.device, dtype=torch.float32)
            for j, i in enumerate(range(RDB_LAYERS - 1, first - 1, -1)):
                blk = Wl[i][:, :, c_lo:c_lo + rows]          # [NB, o, c, 3, 3]
                comb[:, :, j * GROWTH:(j + 1) * GROWTH] = blk.permute(0, 2, 1, 3, 4).flip(3, 4)
            comb[:, :, _align(cmain, 64):, 1, 1] = 0.2 * Wf[:, :, c_lo:c_lo + rows].permute(0, 2, 1)    # [NB, c, d]
            nv.pack_conv_weight(comb.view(NB * rows, cols, 3, 3), self.wslice[s], False)
        if self.sign_bits:
            c4 = F + (RDB_LAYERS - 1) * GROWTH
            self.wlast[:, 4].copy_(torch.stack([self.wb[f"residual_blocks.{k}.lff"][0, c4:c4 + GROWTH] for k in range(NB)]))

    # ---- helpers ---------------------------------------------------------------------------
    def _span(self, kind: str, x: Tensor, cin: int, cout: int, k: int):
        """Timing span for one dense-conv launch; flops = 2 * pixels * Cin * Cout * k^2 (algorithmic)."""
        if self.timer is None:
            return _NOSPAN
        npix = x.shape[0] * x.shape[1] * x.shape[2]
        return self.timer.span(f"{kind}|{cin}>{cout}k{k}", 2.0 * npix * cin * cout * k * k,
                               float(npix) * (cin + cout) * x.element_size())

    def _span_flops(self, kind: str, x: Tensor, flops_per_px: float, channels_moved: int = 0):
        """``channels_moved``: activation channels read + written per pixel (algorithmic bytes = that x element size)."""
        if self.timer is None:
            return _NOSPAN
        npix = x.shape[0] * x.shape[1] * x.shape[2]
        return self.timer.span(kind, flops_per_px * npix, float(npix) * channels_moved * x.element_size())

    def conv(self, name: str, x: Tensor, out: Tensor, P, *, relu=False, res=None, res_channels=0, alpha=1.0,
             bias=True, sign_bits=None) -> None:
        """``sign_bits``: int16 tensor that receives the packed signs of the (ReLU'd) output."""
        c = self.convs[name]
        b = P[name + ".bias"] if (c.has_bias and bias) else None
        with self._span("conv_fwd", x, x.shape[-1], c.cout, c.k):
            nv.conv2d_fwd(x, self.wf[name], b, res, None, None, out, c.cout, relu, False,
                          res_channels if res is not None else 0, 0, alpha, self.engine, None, False, None,
                          sign_bits, 1 if sign_bits is not None else 0)

    def conv_signed(self, A: Activations, key, name: str, x: Tensor, out: Tensor, P, **kw) -> None:
        """Forward conv + ReLU that also leaves the packed signs of its output in ``A.sbits[key]`` (for the data
        gradient that will be gated by this activation) when the library takes the shape."""
        if A.use_bits and key not in self._bits_off:
            b = A.sbits.get(key)
            if b is None:
                b = torch.empty((self.convs[name].cout // 16,) + tuple(out.shape[:3]), device=self.device, dtype=torch.int16)
            try:
                self.conv(name, x, out, P, sign_bits=b, **kw)
                A.sbits[key] = b
                return
            except RuntimeError as e:
                if "not supported" not in str(e):
                    raise
                self._bits_off.add(key)
                A.sbits.pop(key, None)
        self.conv(name, x, out, P, **kw)

    def dgrad(self, name: str, dy: Tensor, out: Tensor, *, cout=None, accumulate=False, res=None, res_channels=0,
              alpha=1.0, mask=None, mask_sub=None, mask_c0=0, bits=None) -> None:
        """``bits``: packed signs of ``mask`` (written by ``conv_signed``); used instead of it where supported."""
        c = self.convs[name]
        with self._span("conv_dgrad", dy, dy.shape[-1], cout or c.cin_pad, c.k):
            if bits is not None and mask is not None and not accumulate and res is None and mask_sub is None:
                try:
                    nv.conv2d_fwd(dy, self.wb[name], None, None, None, None, out, cout or c.cin_pad, False, False, 0, 0,
                                  alpha, self.engine, None, False, None, bits, 2)
                    return
                except RuntimeError as e:
                    if "not supported" not in str(e):
                        raise
            nv.conv2d_fwd(dy, self.wb[name], None, res, mask, mask_sub, out, cout or c.cin_pad, False, accumulate,
                          res_channels if res is not None else 0, mask_c0, alpha, self.engine)

    def wgrad(self, name: str, x: Tensor, dy: Tensor, G: Dict[str, Tensor], scale: float = 1.0,
              dy_padded: Optional[Tensor] = None) -> None:
        """``dy_padded``: the zero-padded (>= 16 channel) buffer ``dy`` is a prefix view of, for the small-Cout
        convs -- lets the row-resident tcgen05 weight gradient take them (it masks the unused columns)."""
        c = self.convs[name]
        tc3 = (dy_padded is not None and c.k == 3 and self.adt == torch.bfloat16 and self.engine != CONV_SIMT
               and self.W >= 64 and x.shape[-1] >= 16)
        if tc3 and c.cout <= 3 and x.shape[-1] % 8 == 0 and x.shape[-1] <= 64:
            # 2-3 output channels: unfold dy (9 shifted copies per channel) and take ONE 1x1 weight-gradient GEMM
            # dW1[o*9+tap, c] = sum_q U[q, o*9+tap] x[q, c]; its centre-tap columns' sums are the bias gradient
            ws = self._workspace()
            cin = x.shape[-1]
            nv.unfold3_grad(dy, ws["unf"])
            tw = ws["wg_tmp"][:32 * cin].view(32, cin, 1, 1)
            tb = ws["wg_tmp"][32 * 64:32 * 64 + 32]
            nv.fill_zero(ws["wg_tmp"])
            with self._span("conv_wgrad", x, cin, dy.shape[-1], c.k):
                nv.conv2d_wgrad(x, ws["unf"], tw, tb, scale, self.engine)
            o = c.cout
            G[name + ".weight"].view(o, cin, 9).add_(tw.view(32, cin)[:o * 9].view(o, 9, cin).transpose(1, 2))
            if c.has_bias:
                G[name + ".bias"].add_(tb[:o * 9].view(o, 9)[:, 4])
            return
        if tc3:
            with self._span("conv_wgrad", x, x.shape[-1], dy.shape[-1], c.k):
                nv.conv3x3_wgrad_grouped(x, dy_padded, [G[name + ".weight"]],
                                         [G[name + ".bias"]] if c.has_bias else [], [0], scale)
            return
        with self._span("conv_wgrad", x, x.shape[-1], dy.shape[-1], c.k):
            nv.conv2d_wgrad(x, dy, G[name + ".weight"], G[name + ".bias"] if c.has_bias else None, scale,
                            self.engine)

    # =======================================================================================
    # forward
    # =======================================================================================
    def forward(self, lr_frames: Tensor, P: Dict[str, Tensor], BUF: Dict[str, Tensor], training: bool,
                need_bwd: bool, out: Tensor) -> Activations:
        """P: parameter name -> fp32 tensor; BUF: BN buffers by name.  Writes the HR frame into ``out``."""
        B, T, H, W, F, s = self.B, self.T, self.H, self.W, self.F, self.scale
        A = self.acquire()
        A.training = training
        A.use_bits = bool(need_bwd and self.sign_bits and self.engine != CONV_SIMT and self.adt == torch.bfloat16)
        if not A.use_bits:
            A.sbits.clear()                  # (a pooled activation set may hold another forward's signs)
        # Pure inference (eval mode, no backward) on the tcgen05 path: BatchNorm's running statistics are folded into
        # the pointwise conv -- W' = W * gamma / sqrt(var + eps) per output channel, b' = beta - mean * that -- so
        # conv + BN + ReLU (+ the extractor skip) is ONE conv launch with a bias / ReLU / residual epilogue and the
        # pre-BN tensor is never written (three full passes over the 64-channel extractor maps less per layer).
        fold = ((not training) and (not need_bwd) and self.adt == torch.bfloat16 and self.engine != CONV_SIMT
                and self.fold_bn)
        fold_scale, fold_bias = {}, {}
        if fold:
            for j in range(3):
                pre = f"feature_extractor.body.{j}."
                sc = P[pre + "bn.weight"] * torch.rsqrt(BUF[pre + "bn.running_var"] + BN_EPS)
                fold_scale[pre + "pointwise"] = sc
                fold_bias[j] = (P[pre + "bn.bias"] - BUF[pre + "bn.running_mean"] * sc).contiguous()
        self.pack_weights(P, need_bwd, fold_scale if fold else None)

        # ---- feature extractor over all T*B frames (super_resolution.py:346-349) ----
        nv.pack_frames_unfold3(lr_frames, A.x_in)
        self.conv("feature_extractor.head.0", A.x_in, A.head, P, relu=True)
        x = A.head
        for j in range(3):
            pre = f"feature_extractor.body.{j}."
            nv.dwconv3x3_fwd(x, P[pre + "depthwise.weight"], A.dwo[j], False, False)
            if fold:
                y = A.fact[j] if j < 2 else A.feat
                with self._span("conv_fwd", A.dwo[j], F, F, 1):
                    nv.conv2d_fwd(A.dwo[j], self.wf[pre + "pointwise"], fold_bias[j], A.head if j == 2 else None, None,
                                  None, y, F, True, False, F if j == 2 else 0, 0, 1.0, self.engine)
                x = y
                continue
            self.conv(pre + "pointwise", A.dwo[j], A.pwo[j], P)
            if training:
                nv.fill_zero(A.bn_sums[j])
                nv.bn_stats(A.pwo[j], T, A.bn_sums[j])
            nv.bn_finalize(A.bn_sums[j] if training else None, A.bn_stat[j], BUF[pre + "bn.running_mean"],
                           BUF[pre + "bn.running_var"], BUF[pre + "bn.num_batches_tracked"], B * H * W, T,
                           BN_MOMENTUM, BN_EPS, training)
            y = A.fact[j] if j < 2 else A.feat
            nv.bn_relu_fwd(A.pwo[j], A.bn_stat[j], P[pre + "bn.weight"], P[pre + "bn.bias"],
                           A.head if j == 2 else None, y, T)
            x = y
        feat = A.feat.view(T, B, H, W, F)
        centre = feat[self.mid]

        # ---- motion estimation + compensation (super_resolution.py:355-363) ----
        nv.axpy(centre, A.cat[..., self.mid * F:(self.mid + 1) * F], 1.0, False)
        for t in self.others:
            nv.corr_fwd(feat[t], centre, A.corr[t])
            self.conv_signed(A, ("fn1", t), "motion_estimator.flow_net.0", A.corr[t], A.fn1[t], P, relu=True)
            self.conv_signed(A, ("fn2", t), "motion_estimator.flow_net.2", A.fn1[t], A.fn2[t], P, relu=True)
            self.conv_signed(A, ("fn3", t), "motion_estimator.flow_net.4", A.fn2[t], A.fn3[t], P, relu=True)
            self.conv("motion_estimator.flow_net.6", A.fn3[t], A.flow[t], P)
            nv.warp_fwd(feat[t], A.flow[t], A.cat[..., t * F:(t + 1) * F], self.div_mode, None)

        # ---- temporal aggregation (super_resolution.py:194-209) ----
        self.conv_signed(A, "a1", "temporal_aggregator.attention.0", A.cat, A.a1, P, relu=True)
        self.conv_signed(A, "a2", "temporal_aggregator.attention.2", A.a1, A.a2, P, relu=True)
        self.conv("temporal_aggregator.attention.4", A.a2, A.logits, P)
        nv.tfuse_fwd(A.cat, A.logits, A.attn, A.blend)
        pre = "temporal_aggregator.refine."
        nv.fill_zero(A.pool)
        nv.chan_sum(A.blend, 1.0 / (H * W), A.pool)
        nv.ca_gate_fwd(A.pool, P[pre + "channel_attention.fc.0.weight"], P[pre + "channel_attention.fc.2.weight"],
                       A.hidden, A.gate)
        nv.cbam_stats_fwd(A.blend, A.gate, A.stats)
        trunk_in = A.rdb[0][..., :F] if self.NB > 0 else A.trunk
        nv.cbam_apply_fwd(A.blend, A.gate, A.stats, P[pre + "spatial_attention.conv.weight"], A.sgate, trunk_in)

        # ---- residual dense blocks (super_resolution.py:245-253), concat-free ----
        use_bits = A.use_bits
        if use_bits and A.rdb_bits is None:
            A.rdb_bits = [[torch.empty((GROWTH // 16, B, H, W), device=self.device, dtype=torch.int16)
                           for _ in range(RDB_LAYERS)] for _ in range(self.NB)]
        for k in range(self.NB):
            buf = A.rdb[k]
            for i in range(RDB_LAYERS):
                c0 = F + i * GROWTH
                name = f"residual_blocks.{k}.layers.{i}.0"
                if use_bits:
                    try:
                        self.conv(name, buf[..., :c0], buf[..., c0:c0 + GROWTH], P, relu=True, sign_bits=A.rdb_bits[k][i])
                        continue
                    except RuntimeError as e:
                        if "not supported" not in str(e) or (k, i) != (0, 0):
                            raise
                        self.sign_bits = use_bits = False          # too small for the CTA-pair kernel: bf16 masks
                self.conv(name, buf[..., :c0], buf[..., c0:c0 + GROWTH], P, relu=True)
            nxt = A.rdb[k + 1][..., :F] if k + 1 < self.NB else A.trunk
            self.conv(f"residual_blocks.{k}.lff", buf, nxt, P, alpha=0.2, res=buf[..., :F], res_channels=F)

        A.has_bits = bool(use_bits)

        # ---- global fusion, upsampler, bicubic skip, clamp (super_resolution.py:372-382) ----
        self.conv("gff.0", A.trunk, A.fused, P, relu=True, res=centre, res_channels=F)
        self.conv("upsampler.conv", A.fused, A.up, P)
        A.lr_centre = lr_frames[:, self.mid]
        nv.upfinish_fwd(A.up, A.lr_centre, out, s)
        return A

    # =======================================================================================
    # backward
    # =======================================================================================
    def _workspace(self):
        if self._bwd_ws is None:
            B, T, H, W, F, s, dev, adt = self.B, self.T, self.H, self.W, self.F, self.scale, self.device, self.adt
            f32 = torch.float32

            def act(n, c, dtype=adt):
                return torch.empty((n, H, W, c), device=dev, dtype=dtype)

            ws = {}
            ws["dup"] = act(B, 3 * s * s, f32)
            # fp32 gradients of the three fp32-output convs, re-cast to the activation dtype in buffers whose
            # channel count is padded to a multiple of 8 (pad channels stay zero: never written after this)
            ws["dup_a"] = torch.zeros((B, H, W, _align(3 * s * s, 16)), device=dev, dtype=adt)
            ws["dfused"] = act(B, F)
            ws["dgff"] = act(B, F)
            ws["dtrunk"] = act(B, F)
            ct = F + RDB_LAYERS * GROWTH
            ws["g"] = [act(B, _align(ct, 64))[..., :ct] for _ in range(2)]
            ws["dz"] = torch.empty((B, H, W), device=dev, dtype=f32)
            ws["dstats"] = act(B, 2, f32)
            ws["dblend"] = act(B, F)
            ws["dgate"] = torch.empty((B, F), device=dev, dtype=f32)
            ws["dpool"] = torch.empty((B, F), device=dev, dtype=f32)
            ws["dcat"] = act(B, T * F)
            ws["dlogits"] = act(B, T, f32)
            ws["dlogits_a"] = torch.zeros((B, H, W, _align(T, 16)), device=dev, dtype=adt)
            ws["da2"], ws["da1"] = act(B, F), act(B, F)
            ws["dfeat"] = act(T * B, F)
            ws["dfeat32"] = act(B, F, f32)
            ws["dflow"] = act(B, 2, f32)
            ws["dflow_a"] = torch.zeros((B, H, W, 16), device=dev, dtype=adt)
            ws["unf"] = torch.zeros((B, H, W, 32), device=dev, dtype=adt)           # unfolded 2-3 channel gradients
            ws["wg_tmp"] = torch.zeros(32 * 64 + 32, device=dev, dtype=f32)         # their 1x1 weight / bias gradients
            ws["dfn3"], ws["dfn2"], ws["dfn1"] = act(B, 32), act(B, 64), act(B, 128)
            ws["dcorr"] = act(B, CORR_PAD)
            ws["dcorr_t"] = act(B, CORR_PAD)          # the centre frame's view of the correlation gradient (corr_bwd workspace)
            ws["t"] = [act(T * B, F) for _ in range(3)]
            ws["bsums"] = torch.zeros((T, F, 2), device=dev, dtype=torch.float64)
            self._bwd_ws = ws
        return self._bwd_ws

    def _rdb_backward_layers(self, k: int, buf: Tensor, g: Tensor, dblock: Tensor, G: Dict[str, Tensor]) -> None:
        """Layer-by-layer dense-block backward (fp32 parity path and small shapes): each layer's data gradient is
        accumulated into the shared gradient buffer by the conv epilogue."""
        F = self.F
        CT = F + RDB_LAYERS * GROWTH
        name = f"residual_blocks.{k}.lff"
        # g[:, :CT] = 0.2 * lff^T(dblock) (+ dblock on the first F channels: the block skip);
        # slice 4 (channels >= F+4G) has no other consumer, so its ReLU mask is applied here.
        self.dgrad(name, dblock, g, cout=CT, alpha=0.2, res=dblock, res_channels=F, mask=buf,
                   mask_c0=F + (RDB_LAYERS - 1) * GROWTH)
        for i in reversed(range(RDB_LAYERS)):
            c0 = F + i * GROWTH
            name = f"residual_blocks.{k}.layers.{i}.0"
            dy = g[..., c0:c0 + GROWTH]
            self.wgrad(name, buf[..., :c0], dy, G)
            if i >= 1:
                self.dgrad(name, dy, g[..., :c0], cout=c0, accumulate=True, mask=buf, mask_c0=c0 - GROWTH)
            else:
                self.dgrad(name, dy, g[..., :c0], cout=c0, accumulate=True)

    def _masked_conv(self, x, w, mask, bits, out, rows, alpha, x2, colsum) -> None:
        """Mask-gated data-gradient conv of the fused block backward: the mask from the packed sign bits where the
        library takes them, from the forward activation otherwise."""
        if bits is not None:
            try:
                nv.conv2d_fwd(x, w, None, None, None, None, out, rows, False, False, 0, 0, alpha, self.engine, x2,
                              x2 is not None, colsum, bits, 2)
                return
            except RuntimeError as e:
                if "not supported" not in str(e):
                    raise
        nv.conv2d_fwd(x, w, None, None, mask, None, out, rows, False, False, 0, 0, alpha, self.engine, x2,
                      x2 is not None, colsum)

    def _rdb_backward_fused(self, k: int, buf: Tensor, g: Tensor, dblock: Tensor, G: Dict[str, Tensor],
                            bits: Optional[List[Tensor]] = None) -> None:
        """Dense-block backward as 1 + 5 convolutions that each WRITE one slice of the gradient buffer once
        (no read-modify-write accumulation) followed by one grouped weight-gradient GEMM.

        Gradient-buffer layout (differs from the forward buffer's): ``[x | layer 4 | layer 3 | ... | layer 0]``.  The
        gradient of a slice reads the gradients of ALL LATER layers; in descending order those are the channels
        ``[F, F + 32 n)`` for every slice, so each 64-channel TMA box row is one aligned 128-byte line (in forward
        order the later layers of slices 1 and 3 start 64 bytes into a line: 0.05-0.07 ms more per launch,
        ``scripts/bench_conv.py align``)."""
        F = self.F
        CT = F + RDB_LAYERS * GROWTH

        def gpos(i):                                          # first channel of layer i's output gradient
            return F + (RDB_LAYERS - 1 - i) * GROWTH

        lff = f"residual_blocks.{k}.lff"
        # last slice: only the LFF reaches it.  dy_4 = relu'(o_4) * 0.2 * W_lff[:, slice]^T dblock
        c4 = F + (RDB_LAYERS - 1) * GROWTH
        # (every slice gradient g is the previous layer's output gradient: its per-channel sum, taken in the conv
        #  epilogue, is that layer's bias gradient -- no reduction pass over the gradient buffer)
        names = [f"residual_blocks.{k}.layers.{i}.0" for i in range(RDB_LAYERS)]
        g4 = gpos(RDB_LAYERS - 1)
        with self._span("conv_dgrad", dblock, F, GROWTH, 1):
            if bits is not None:
                self._masked_conv(dblock, self.wlast[k], buf[..., c4:CT], bits[RDB_LAYERS - 1], g[..., g4:g4 + GROWTH],
                                  GROWTH, 0.2, None, G[names[RDB_LAYERS - 1] + ".bias"])
            else:
                nv.conv2d_fwd(dblock, self.wb[lff][:, c4:CT, :], None, None, buf[..., c4:CT], None,
                              g[..., g4:g4 + GROWTH], GROWTH, False, False, 0, 0, 0.2, self.engine, None, False,
                              G[names[RDB_LAYERS - 1] + ".bias"])
        for s in range(RDB_LAYERS - 1, -1, -1):              # slices F+(s-1)G .. (s >= 1), then the x slice (s = 0)
            rows = F if s == 0 else GROWTH
            c_lo = 0 if s == 0 else F + (s - 1) * GROWTH      # the slice in the FORWARD buffer (its ReLU mask)
            first = 0 if s == 0 else s                        # first later layer
            n_later = GROWTH * (RDB_LAYERS - first)           # their gradients: channels [F, F + n_later)
            o_lo = 0 if s == 0 else gpos(s - 1)               # where this slice's gradient goes (== F + n_later)
            w = self.wslice[s][:, k * rows:(k + 1) * rows, :]
            mask = buf[..., c_lo:c_lo + rows] if s > 0 else None
            # bytes: later layers' gradients + the block gradient (1x1 branch; + once more as the x slice's residual) in,
            # the slice's ReLU mask in (s > 0), the slice out
            moved = n_later + F + (F if s == 0 else rows) + rows
            with self._span_flops(f"conv_dgrad|slice{s}", buf, 2.0 * rows * (9 * n_later + F), moved):
                if s > 0:
                    self._masked_conv(g[..., F:F + n_later], w, mask, bits[s - 1] if bits is not None else None,
                                      g[..., o_lo:o_lo + rows], rows, 1.0, dblock, G[names[s - 1] + ".bias"])
                else:
                    # the x slice also receives the block's own skip connection (+ dblock) as the epilogue residual
                    nv.conv2d_fwd(g[..., F:F + n_later], w, None, dblock, None, None, g[..., :rows], rows, False, False,
                                  F, 0, 1.0, self.engine, dblock, True, None)
        cx = F + (RDB_LAYERS - 1) * GROWTH
        with self._span_flops("conv_wgrad|rdb_grouped", buf, sum(2.0 * 9 * (F + i * GROWTH) * GROWTH
                                                                for i in range(RDB_LAYERS)), cx + RDB_LAYERS * GROWTH):
            nv.conv3x3_wgrad_grouped(buf[..., :cx], g[..., F:CT], [G[n + ".weight"] for n in names], [],
                                     [gpos(i) - F for i in range(RDB_LAYERS)], 1.0)

    def backward(self, A: Activations, dout: Tensor, P: Dict[str, Tensor], G: Dict[str, Tensor],
                 on_grads_ready=None) -> None:
        """Accumulate parameter gradients into ``G`` (name -> fp32 tensor, same shapes as ``P``).

        ``G`` must be zero-initialised by the caller.  ``on_grads_ready(prefix)`` is invoked as soon as
        every gradient under a parameter-name prefix is final (used to launch bucketed all-reduces
        while the rest of backward is still running).
        """
        B, T, H, W, F, s = self.B, self.T, self.H, self.W, self.F, self.scale
        ws = self._workspace()
        ready = on_grads_ready or (lambda prefix: None)
        feat = A.feat.view(T, B, H, W, F)
        centre = feat[self.mid]
        ncs = 3 * s * s
        training = A.training

        # ---- output stage ----
        nv.upfinish_bwd(A.up, A.lr_centre, dout, ws["dup"], s)
        dup = ws["dup_a"][..., :ncs]
        nv.axpy(ws["dup"], dup, 1.0, False)
        self.wgrad("upsampler.conv", A.fused, dup, G, dy_padded=ws["dup_a"])
        ready("upsampler.")
        self.dgrad("upsampler.conv", ws["dup_a"], ws["dfused"])      # padded view: zero channels x zero weights
        nv.relu_bwd(ws["dfused"], A.fused, centre, ws["dgff"])
        self.wgrad("gff.0", A.trunk, ws["dgff"], G)
        ready("gff.")
        dblock = ws["dtrunk"]
        self.dgrad("gff.0", ws["dgff"], dblock)

        # ---- residual dense blocks, last to first ----
        CT = F + RDB_LAYERS * GROWTH
        fused = self.fused_rdb_bwd and self.engine != CONV_SIMT
        if fused:
            self.pack_rdb_slice_weights(P)
        for k in reversed(range(self.NB)):
            buf, g = A.rdb[k], ws["g"][k & 1]
            name = f"residual_blocks.{k}.lff"
            self.wgrad(name, buf, dblock, G, scale=0.2)
            if fused:
                self._rdb_backward_fused(k, buf, g, dblock, G, A.rdb_bits[k] if A.has_bits else None)
            else:
                self._rdb_backward_layers(k, buf, g, dblock, G)
            ready(f"residual_blocks.{k}.")
            dblock = g[..., :F]
        dagg = dblock

        # ---- CBAM ----
        pre = "temporal_aggregator.refine."
        w7 = P[pre + "spatial_attention.conv.weight"]
        nv.cbam_bwd_dz(A.blend, A.gate, A.sgate, dagg, ws["dz"])
        nv.cbam_bwd_spatial(ws["dz"], A.stats, w7, ws["dstats"], G[pre + "spatial_attention.conv.weight"])
        nv.fill_zero(ws["dgate"])
        nv.cbam_bwd_dx(A.blend, A.gate, A.sgate, A.stats, ws["dstats"], dagg, ws["dblend"], ws["dgate"])
        nv.ca_gate_bwd(A.pool, P[pre + "channel_attention.fc.0.weight"], P[pre + "channel_attention.fc.2.weight"],
                       A.hidden, A.gate, ws["dgate"], ws["dpool"], G[pre + "channel_attention.fc.0.weight"],
                       G[pre + "channel_attention.fc.2.weight"])
        nv.ewc_axpby(ws["dpool"], None, 1.0 / (H * W), 0.0)     # d(mean over H*W)

        # ---- temporal fusion + attention convs ----
        nv.tfuse_bwd(A.cat, A.attn, ws["dblend"], ws["dpool"], ws["dcat"], ws["dlogits"])
        dlog = ws["dlogits_a"][..., :T]
        nv.axpy(ws["dlogits"], dlog, 1.0, False)
        self.wgrad("temporal_aggregator.attention.4", A.a2, dlog, G, dy_padded=ws["dlogits_a"])
        self.dgrad("temporal_aggregator.attention.4", ws["dlogits_a"], ws["da2"], mask=A.a2, bits=A.sbits.get("a2"))
        self.wgrad("temporal_aggregator.attention.2", A.a1, ws["da2"], G)
        self.dgrad("temporal_aggregator.attention.2", ws["da2"], ws["da1"], mask=A.a1, bits=A.sbits.get("a1"))
        self.wgrad("temporal_aggregator.attention.0", A.cat, ws["da1"], G)
        self.dgrad("temporal_aggregator.attention.0", ws["da1"], ws["dcat"], accumulate=True)
        ready("temporal_aggregator.")

        # ---- alignment: warp, flow_net, correlation ----
        dfeat = ws["dfeat"].view(T, B, H, W, F)
        dcat = ws["dcat"]
        nv.axpy(dcat[..., self.mid * F:(self.mid + 1) * F], dfeat[self.mid], 1.0, False)
        nv.axpy(ws["dfused"], dfeat[self.mid], 1.0, True)        # gff skip: fused = relu(..) + centre
        for t in self.others:
            if self.adt == torch.bfloat16:
                # scatter straight into the bf16 gradient (packed 8 x bf16 reductions): no fp32 staging round trip
                nv.fill_zero(dfeat[t])
                nv.warp_bwd_lp(feat[t], A.flow[t], dcat[..., t * F:(t + 1) * F], dfeat[t], ws["dflow"], self.div_mode)
            else:
                nv.fill_zero(ws["dfeat32"])
                nv.warp_bwd(feat[t], A.flow[t], dcat[..., t * F:(t + 1) * F], ws["dfeat32"], ws["dflow"], self.div_mode)
                nv.axpy(ws["dfeat32"], dfeat[t], 1.0, False)
            dflow = ws["dflow_a"][..., :2]
            nv.axpy(ws["dflow"], dflow, 1.0, False)
            self.wgrad("motion_estimator.flow_net.6", A.fn3[t], dflow, G, dy_padded=ws["dflow_a"])
            self.dgrad("motion_estimator.flow_net.6", ws["dflow_a"], ws["dfn3"], mask=A.fn3[t], bits=A.sbits.get(("fn3", t)))
            self.wgrad("motion_estimator.flow_net.4", A.fn2[t], ws["dfn3"], G)
            self.dgrad("motion_estimator.flow_net.4", ws["dfn3"], ws["dfn2"], mask=A.fn2[t], bits=A.sbits.get(("fn2", t)))
            self.wgrad("motion_estimator.flow_net.2", A.fn1[t], ws["dfn2"], G)
            self.dgrad("motion_estimator.flow_net.2", ws["dfn2"], ws["dfn1"], mask=A.fn1[t], bits=A.sbits.get(("fn1", t)))
            self.wgrad("motion_estimator.flow_net.0", A.corr[t][..., :CORR_CH], ws["dfn1"], G)
            self.dgrad("motion_estimator.flow_net.0", ws["dfn1"], ws["dcorr"])
            nv.corr_bwd(feat[t], centre, ws["dcorr"], dfeat[t], True, dfeat[self.mid], True, ws["dcorr_t"])
        ready("motion_estimator.")

        # ---- feature extractor (all T*B frames at once, BN per frame group) ----
        # dy: gradient w.r.t. the ReLU output of body layer j.  s0/s1/s2 are scratch: by the time s2
        # (the previous dy) is overwritten by the depthwise data gradient it has been consumed.
        dy = ws["dfeat"]
        s0, s1, s2 = ws["t"]
        for j in reversed(range(3)):
            pre = f"feature_extractor.body.{j}."
            gamma, beta = P[pre + "bn.weight"], P[pre + "bn.bias"]
            nv.fill_zero(ws["bsums"])
            nv.bn_relu_bwd_reduce(A.pwo[j], dy, A.bn_stat[j], gamma, beta, T, ws["bsums"])
            nv.bn_relu_bwd_apply(A.pwo[j], dy, A.bn_stat[j], gamma, beta, ws["bsums"], s0, G[pre + "bn.weight"],
                                 G[pre + "bn.bias"], T, training)
            self.wgrad(pre + "pointwise", A.dwo[j], s0, G)
            self.dgrad(pre + "pointwise", s0, s1)
            x_in = A.fact[j - 1] if j > 0 else A.head
            nv.dwconv3x3_wgrad(x_in, s1, G[pre + "depthwise.weight"])
            if j > 0:
                nv.dwconv3x3_fwd(s1, P[pre + "depthwise.weight"], s2, True, False)
                dy = s2
            elif self.adt == torch.bfloat16 and self.engine != CONV_SIMT:
                # extractor skip (feat = body(head) + head): d(head) = d(feat) + depthwise data gradient, gated by the
                # head conv's ReLU -- one pass (s0 is free again: the pointwise gradients above consumed it)
                nv.dwconv3x3_fwd_masked(s1, P[pre + "depthwise.weight"], ws["dfeat"], A.head, s0, True)
            else:
                nv.dwconv3x3_fwd(s1, P[pre + "depthwise.weight"], ws["dfeat"], True, True)
                nv.relu_bwd(ws["dfeat"], A.head, None, s0)
        gw = G["feature_extractor.head.0.weight"]
        with self._span("conv_wgrad", A.x_in, 27, F, 1):
            nv.conv2d_wgrad(A.x_in[..., :27], ws["t"][0], gw.view(gw.shape[0], -1, 1, 1),
                            G["feature_extractor.head.0.bias"], 1.0, self.engine)
        ready("feature_extractor.")
