"""CUDA-graph replay of a whole training step (forward + loss + backward + fused optimiser step).

At the benchmark shape a step is ~100 ms of GPU work behind ~230 kernel launches, so launch overhead is invisible.
At the reference launcher's own configuration (32 features / 4 dense blocks / 64x64 crops / batch 16,
``experiments/train_baseline.py``) the same ~230 launches carry well under a millisecond of work each and the step is
bound by Python + launch latency.  Shapes are static per configuration, every op of the path enqueues on the current
stream without host synchronisation, and the library keeps no mutable host state, so the whole step can be captured
once and replayed:

    step = GraphedTrainStep(model, opt)            # model: SuperResolutionNet, opt: FlatAdamW
    for lr, hr in loader:
        loss = step(lr, hr)                        # copies into the static inputs, replays the graph

What is baked into the graph: the learning rate and the other optimiser hyper-parameters (re-capture after changing
them, or call ``step.recapture()``), the batch shape, train/eval mode.  What is NOT: the step count of the bias
corrections (kept in a device counter that the graph increments), BatchNorm's running statistics and
``num_batches_tracked`` (updated by captured kernels).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .optim import FlatAdamW

Tensor = torch.Tensor


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: FlatAdamW,
                 loss_fn: Callable[[Tensor, Tensor], Tensor] = torch.nn.functional.mse_loss, warmup: int = 3):
        self.model, self.opt, self.loss_fn, self.warmup = model, optimizer, loss_fn, max(int(warmup), 1)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_lr: Optional[Tensor] = None
        self.static_hr: Optional[Tensor] = None
        self.loss: Optional[Tensor] = None

    def _eager(self, lr: Tensor, hr: Tensor) -> Tensor:
        self.opt.zero_grad()
        loss = self.loss_fn(self.model(lr), hr)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def _capture(self, lr: Tensor, hr: Tensor) -> None:
        if getattr(self.model, "_grad_sync", None) is not None:
            raise RuntimeError("GraphedTrainStep: capture a single-process step (no gradient sync installed)")
        self.opt.device_step = True
        self.opt.step_dev.fill_(self.opt.step_count)
        self.static_lr, self.static_hr = lr.clone(), hr.clone()
        # warm-up on a side stream (allocates the shape plan, the activation pool and the optimiser's buffers); these
        # are REAL optimiser steps on the first batch
        side = torch.cuda.Stream(device=lr.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._eager(self.static_lr, self.static_hr)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager(self.static_lr, self.static_hr)
        # the captured step was recorded, not run: the host-side counter already moved on
        self.opt.step_count -= 1

    def recapture(self) -> None:
        self.graph = None

    def __call__(self, lr: Tensor, hr: Tensor) -> Tensor:
        """One optimiser step on (lr, hr); returns the (static, device-resident) loss tensor of that step."""
        if self.graph is None or self.static_lr.shape != lr.shape or self.static_hr.shape != hr.shape:
            self._capture(lr, hr)
        self.static_lr.copy_(lr, non_blocking=True)
        self.static_hr.copy_(hr, non_blocking=True)
        self.graph.replay()
        self.opt.step_count += 1
        return self.loss
