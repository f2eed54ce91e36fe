"""Drop-in ``SynapticIntelligence`` (reference ``nerve_cl/continual/ewc.py:306-379``) on the flat-buffer kernels.

Same constructor and methods as the reference -- ``SynapticIntelligence(model, si_lambda=1.0, damping=0.1)``,
``update_importance()`` after every optimiser step, ``register_task()`` at a task boundary, ``penalty()`` -- and the
same public dicts ``W`` / ``p_old`` / ``omega`` (name -> tensor), whose values are views into three flat fp32
buffers so that each method is ONE kernel launch per 32 parameter tensors instead of 3-5 ATen launches per tensor:

* ``update_importance``:  W += -grad * (theta - p_old);  p_old = theta         -> ``nervecl::si_update``
* ``register_task``:      omega += W / ((theta - p_old)^2 + damping); W = 0; p_old = theta   -> ``nervecl::si_register``
* ``penalty``:            si_lambda * sum omega (theta - p_old)^2  (+ autograd)  -> ``nervecl::ewc_penalty_fwd/bwd``

Reference behaviour kept on purpose: ``update_importance`` moves ``p_old`` to the current parameters every step, so by
the time ``register_task`` runs ``theta - p_old`` is the LAST step's movement (zero if nothing stepped in between) and
the normaliser is essentially ``damping``; a parameter whose ``.grad`` is ``None`` keeps both its ``W`` and ``p_old``.
CUDA only: a model on CPU raises ``RuntimeError``.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from .. import ops as _ops
from .ewc import _FlatState, _PenaltyFn

Tensor = torch.Tensor
nv = _ops.nv


class SynapticIntelligence:
    def __init__(self, model: nn.Module, si_lambda: float = 1.0, damping: float = 0.1):
        self.model = model
        self.si_lambda = si_lambda
        self.damping = damping
        self.W: Dict[str, Tensor] = {}
        self.p_old: Dict[str, Tensor] = {}
        self.omega: Dict[str, Tensor] = {}
        self._init_tracking()

    def _params(self):
        return [(n, p) for n, p in self.model.named_parameters() if n in self.W]

    def _init_tracking(self) -> None:
        """Reference ewc.py:333-339: W = 0, p_old = theta, omega = 0 for every trainable parameter."""
        named = [(n, p) for n, p in self.model.named_parameters() if p.requires_grad]
        if not named:
            return
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("nerve_cl_b200.SynapticIntelligence runs on CUDA (sm_100a) only; there is no CPU "
                               "fallback. Move the model to a B200 first.")
        layout = [(n, p.numel(), p.shape) for n, p in named]
        self._W, self._omega = _FlatState(layout, dev), _FlatState(layout, dev)
        self._p_old = _FlatState(layout, dev, zero=False)
        with torch.no_grad():
            for n, p in named:
                self._p_old.views[n].copy_(p.data)
        self.W, self.p_old, self.omega = self._W.views, self._p_old.views, self._omega.views

    def update_importance(self) -> None:
        """Call after each optimiser step (reference ewc.py:342-352)."""
        named = self._params()
        theta = [p.detach() for _, p in named]
        grads = [None if p.grad is None else p.grad.detach().contiguous() for _, p in named]
        nv.si_update(theta, grads, self._W.flat, self._p_old.flat)

    def register_task(self) -> None:
        """Reference ewc.py:354-366."""
        theta = [p.detach() for _, p in self._params()]
        nv.si_register(theta, self._W.flat, self._p_old.flat, self._omega.flat, float(self.damping))

    def penalty(self):
        """si_lambda * sum omega (theta - p_old)^2 (reference ewc.py:368-379)."""
        named = self._params()
        if not named:
            return self.si_lambda * 0.0
        return _PenaltyFn.apply(self.model, float(self.si_lambda), [(self._omega, self._p_old)], *[p for _, p in named])
