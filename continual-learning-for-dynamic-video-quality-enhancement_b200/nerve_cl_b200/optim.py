"""Fused AdamW over the module's flat parameter / gradient buffers.

``torch.optim.AdamW`` semantics (decoupled weight decay, bias correction; the reference's optimiser at
experiments/train_baseline.py:62) as ONE ``nervecl::adamw_step`` launch per step instead of a
multi-tensor apply over 131 tensors.  Works on any module exposing ``_flat_layout`` / ``_flat_numel`` /
``last_flat_grad()`` (``SuperResolutionNet``); parameters are re-homed once into a flat fp32 buffer
(``param.data`` become views, values unchanged).  If the gradients are not the engine's flat buffer --
``loss = mse + ewc.penalty()``: autograd sums the two incoming gradients of every leaf out of place, so each
``param.grad`` is its own tensor -- they are gathered into a flat buffer by ``nervecl::flat_gather`` (one launch
per 32 tensors) and the step is still ONE ``adamw_step`` launch.  Only when some ``param.grad`` is ``None``
(torch.optim.AdamW leaves such a parameter completely untouched) does it step per tensor -- same kernel, never ATen.

``state_dict()`` / ``load_state_dict()`` use ``torch.optim.AdamW``'s format (``state`` / ``param_groups`` keyed by
parameter index), so optimiser checkpoints are exchangeable with the reference's ``train_baseline.py:62``.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops as _ops

nv = _ops.nv


def flatten_parameters(module) -> torch.Tensor:
    """Move every parameter of ``module`` into one flat fp32 buffer laid out by ``module._flat_layout``."""
    flat = getattr(module, "_flat_params", None)
    p0 = next(module.parameters())
    if flat is not None and flat.device == p0.device:
        ok = all(p.data_ptr() == flat.data_ptr() + 4 * module._flat_layout[n][0]
                 for n, p in module.named_parameters())
        if ok:
            return flat
    flat = torch.zeros(module._flat_numel, device=p0.device, dtype=torch.float32)
    with torch.no_grad():
        for n, p in module.named_parameters():
            off, k, shape = module._flat_layout[n]
            view = flat[off:off + k].view(shape)
            view.copy_(p.data)
            p.data = view
    module._flat_params = flat
    return flat


class FlatAdamW:
    def __init__(self, module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        self.module = module
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.flat = flatten_parameters(module)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = 0
        # (the remaining keys are torch.optim.AdamW's defaults, so a saved group loads into the torch optimiser as is)
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay, "amsgrad": False,
                              "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                              "fused": None, "decoupled_weight_decay": True, "params": list(module.parameters())}]
        self._gather: Optional[torch.Tensor] = None
        # step count on the device as well (incremented by a captured kernel): lets graphs.GraphedTrainStep replay
        # the optimiser step with the right bias corrections
        self.step_dev = torch.zeros(1, device=self.flat.device, dtype=torch.int32)
        self.device_step = False

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.module.parameters():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _check_aliasing(self) -> None:
        """``module.to()`` / ``.float()`` re-allocate ``param.data`` and silently break the flat aliasing: re-flatten."""
        n0, p0 = next(iter(self.module.named_parameters()))
        if p0.data_ptr() != self.flat.data_ptr() + 4 * self.module._flat_layout[n0][0] or p0.device != self.flat.device:
            old = self.flat
            self.module._flat_params = None
            self.flat = flatten_parameters(self.module)
            if self.flat.device != old.device:
                self.exp_avg, self.exp_avg_sq = self.exp_avg.to(self.flat.device), self.exp_avg_sq.to(self.flat.device)
            self._gather = None

    @torch.no_grad()
    def step(self) -> None:
        self._check_aliasing()
        self.step_count += 1
        if self.device_step:
            self.step_dev.add_(1)
        sd = self.step_dev if self.device_step else None
        grp = self.param_groups[0]
        lr, (b1, b2), eps, wd = grp["lr"], grp["betas"], grp["eps"], grp["weight_decay"]
        g: Optional[torch.Tensor] = self.module.last_flat_grad()
        if g is None:
            named = list(self.module.named_parameters())
            if all(p.grad is not None for _, p in named):
                if getattr(self, "_gather", None) is None:
                    self._gather = torch.zeros_like(self.flat)
                nv.flat_gather([p.grad.detach().contiguous() for _, p in named],
                               [self.module._flat_layout[n][0] for n, _ in named], self._gather)
                g = self._gather
        if g is not None:
            nv.adamw_step(self.flat, g, self.exp_avg, self.exp_avg_sq, lr, b1, b2, eps, wd, self.step_count, 1.0, sd)
            return
        for n, p in self.module.named_parameters():
            if p.grad is None:
                continue
            off, k, _ = self.module._flat_layout[n]
            nv.adamw_step(self.flat[off:off + k], p.grad.contiguous().view(-1), self.exp_avg[off:off + k],
                          self.exp_avg_sq[off:off + k], lr, b1, b2, eps, wd, self.step_count, 1.0, sd)

    # ---- torch.optim.AdamW-compatible checkpoints -------------------------------------------------------------
    def state_dict(self):
        """``{'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}`` like ``torch.optim.AdamW``."""
        state = {}
        if self.step_count > 0:
            for i, (n, _) in enumerate(self.module.named_parameters()):
                off, k, shape = self.module._flat_layout[n]
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[off:off + k].view(shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + k].view(shape).clone()}
        grp = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        grp["params"] = list(range(len(list(self.module.parameters()))))
        return {"state": state, "param_groups": [grp]}

    def load_state_dict(self, sd) -> None:
        """Accepts ``torch.optim.AdamW.state_dict()`` (and this class's own, which has the same format)."""
        grp = sd["param_groups"][0]
        for key in ("lr", "betas", "eps", "weight_decay"):
            if key in grp:
                self.param_groups[0][key] = tuple(grp[key]) if key == "betas" else grp[key]
        self.lr, self.betas = self.param_groups[0]["lr"], self.param_groups[0]["betas"]
        self.eps, self.weight_decay = self.param_groups[0]["eps"], self.param_groups[0]["weight_decay"]
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count = 0
        for i, (n, _) in enumerate(self.module.named_parameters()):
            st = sd["state"].get(i)
            if st is None:
                continue
            off, k, shape = self.module._flat_layout[n]
            self.exp_avg[off:off + k].view(shape).copy_(st["exp_avg"])
            self.exp_avg_sq[off:off + k].view(shape).copy_(st["exp_avg_sq"])
            self.step_count = max(self.step_count, int(st["step"]))
        self.step_dev.fill_(self.step_count)
