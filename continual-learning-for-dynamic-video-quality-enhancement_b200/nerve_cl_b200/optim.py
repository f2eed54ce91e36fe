"""Fused AdamW over the module's flat parameter / gradient buffers.

``torch.optim.AdamW`` semantics (decoupled weight decay, bias correction; the reference's optimiser at
experiments/train_baseline.py:62) as ONE ``nervecl::adamw_step`` launch per step instead of a
multi-tensor apply over 131 tensors.  Works on any module exposing ``_flat_layout`` / ``_flat_numel`` /
``last_flat_grad()`` (``SuperResolutionNet``); parameters are re-homed once into a flat fp32 buffer
(``param.data`` become views, values unchanged).  If the gradients are not the engine's flat buffer
(e.g. they were accumulated over several backward passes) it falls back to one launch per tensor --
still the same kernel, never ATen.
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
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay,
                              "params": list(module.parameters())}]

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.module.parameters():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self) -> None:
        self.step_count += 1
        lr = self.param_groups[0]["lr"]
        b1, b2 = self.betas
        g: Optional[torch.Tensor] = self.module.last_flat_grad()
        if g is not None:
            nv.adamw_step(self.flat, g, self.exp_avg, self.exp_avg_sq, lr, b1, b2, self.eps, self.weight_decay,
                          self.step_count, 1.0)
            return
        for n, p in self.module.named_parameters():
            if p.grad is None:
                continue
            off, k, _ = self.module._flat_layout[n]
            nv.adamw_step(self.flat[off:off + k], p.grad.contiguous().view(-1), self.exp_avg[off:off + k],
                          self.exp_avg_sq[off:off + k], lr, b1, b2, self.eps, self.weight_decay, self.step_count, 1.0)

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg.cpu(), "exp_avg_sq": self.exp_avg_sq.cpu(),
                "lr": self.param_groups[0]["lr"], "betas": self.betas, "eps": self.eps,
                "weight_decay": self.weight_decay}

    def load_state_dict(self, sd) -> None:
        self.step_count = sd["step"]
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0]["lr"] = sd["lr"]
