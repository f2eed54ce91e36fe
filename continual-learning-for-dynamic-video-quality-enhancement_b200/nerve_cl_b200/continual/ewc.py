"""Drop-in ``EWC`` / ``OnlineEWC`` (reference ``nerve_cl/continual/ewc.py:19-303``) on fused kernels.

Public surface is the reference's: constructor ``EWC(model, ewc_lambda=5000.0, mode='online',
decay=0.999)``, dicts ``fisher_dict`` / ``optpar_dict`` / ``task_fisher`` / ``task_optpar``,
``num_tasks``, ``compute_fisher``, ``register_task``, ``penalty``, ``get_importance_stats``,
``state_dict`` / ``load_state_dict`` (same 8 keys).  The per-name dict values are *views into one
flat fp32 buffer*, so the name-keyed API and the single-kernel arithmetic coexist:

* Fisher accumulation  F += g^2            -> ``nervecl::ewc_fisher_accum``   (ewc.py:139-141)
* normalisation / online consolidation     -> ``nervecl::ewc_axpby``          (ewc.py:146-147,186-190)
* penalty  lambda/2 * sum F (theta-theta*)^2  and its gradient  lambda F (theta-theta*)
                                           -> ``nervecl::ewc_penalty_fwd/bwd`` (ewc.py:226-232)

instead of ~5 ATen launches per parameter tensor (656 fwd + 788 bwd ops for the 131 tensors of the
SR model).  CUDA only: a model on CPU raises ``RuntimeError``.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops as _ops

Tensor = torch.Tensor
nv = _ops.nv


class _FlatState:
    """A flat fp32 buffer plus name -> view dict, laid out in ``named_parameters`` order."""

    def __init__(self, layout: List[Tuple[str, int, torch.Size]], device, zero: bool = True):
        self.layout = layout
        total = sum(n for _, n, _ in layout)
        self.flat = (torch.zeros if zero else torch.empty)(max(total, 1), device=device, dtype=torch.float32)
        self.views: Dict[str, Tensor] = {}
        off = 0
        for name, n, shape in layout:
            self.views[name] = self.flat[off:off + n].view(shape)
            off += n

    @classmethod
    def from_dict(cls, layout, tensors: Dict[str, Tensor], device) -> "_FlatState":
        st = cls(layout, device)
        for name, _, _ in layout:
            if name in tensors:
                st.views[name].copy_(tensors[name])
        return st


class _PenaltyFn(torch.autograd.Function):
    """penalty = coef * sum_states sum F (theta - star)^2 as one autograd node over all parameters."""

    @staticmethod
    def forward(ctx, model, coef: float, states: List[Tuple[_FlatState, _FlatState]], *params: Tensor):
        theta = [p.detach() for p in params]
        out = torch.zeros((), device=theta[0].device, dtype=torch.float32)
        for fisher, star in states:
            nv.ewc_penalty_fwd(theta, fisher.flat, star.flat, coef, out)
        ctx.model, ctx.coef, ctx.states, ctx.theta = model, coef, states, theta
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        theta = ctx.theta
        layout = ctx.states[0][0].layout
        g = _FlatState(layout, theta[0].device)             # fresh zeroed flat gradient
        grads = [g.views[name] for name, _, _ in layout]
        gs = gout.detach().reshape(1).float().contiguous()
        for fisher, star in ctx.states:
            nv.ewc_penalty_bwd(theta, grads, fisher.flat, star.flat, 2.0 * ctx.coef, gs)
        return (None, None, None) + tuple(grads)


class EWC:
    """Elastic Weight Consolidation; see the module docstring and reference ewc.py:19-65."""

    def __init__(self, model: nn.Module, ewc_lambda: float = 5000.0, mode: str = "online", decay: float = 0.999):
        self.model = model
        self.ewc_lambda = ewc_lambda
        self.mode = mode
        self.decay = decay
        self.fisher_dict: Dict[str, Tensor] = {}
        self.optpar_dict: Dict[str, Tensor] = {}
        self.task_fisher: Dict[int, Dict[str, Tensor]] = {}
        self.task_optpar: Dict[int, Dict[str, Tensor]] = {}
        self.num_tasks = 0
        self.process_group = None        # set to a torch.distributed group to sum Fisher over ranks
        # flat backing stores of the dicts above (online: one pair; separate: one pair per task)
        self._online: Optional[Tuple[_FlatState, _FlatState]] = None
        self._tasks: Dict[int, Tuple[_FlatState, _FlatState]] = {}

    # ------------------------------------------------------------------------------------
    def _get_params(self) -> Iterator[tuple]:
        for name, param in self.model.named_parameters():
            if param.requires_grad:
                yield name, param

    def _layout(self) -> List[Tuple[str, int, torch.Size]]:
        return [(n, p.numel(), p.shape) for n, p in self._get_params()]

    def _device(self) -> torch.device:
        dev = next(self.model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("nerve_cl_b200.EWC runs on CUDA (sm_100a) only; there is no CPU fallback. "
                               "Move the model to a B200 first.")
        return dev

    # ------------------------------------------------------------------------------------
    def compute_fisher(self, dataloader, num_samples: Optional[int] = None, empirical: bool = True
                       ) -> Dict[str, Tensor]:
        """Diagonal Fisher estimate exactly as reference ewc.py:73-149 defines it:
        sum over batches of (gradient of the batch-mean loss)^2, divided by the number of samples.
        Leaves the model in ``eval()`` mode with ``.grad`` populated, like the reference."""
        dev = self._device()
        layout = self._layout()
        fisher = _FlatState(layout, dev)
        params = [p for _, p in self._get_params()]
        numels = [p.numel() for p in params]
        self.model.eval()
        used = 0
        # Data-parallel runs: the Fisher is sum over LOCAL batches of (local batch-mean gradient)^2, summed over ranks
        # below.  A gradient all-reduce inside backward would square the rank-AVERAGED
        # gradient instead (and hang when ranks see different batch counts), so it is switched off for this pass.
        synced = [(m, m._grad_sync) for m in self.model.modules() if getattr(m, "_grad_sync", None) is not None]
        for m, _ in synced:
            m._grad_sync = None
        try:
            for batch in dataloader:
                if num_samples is not None and used >= num_samples:
                    break
                if isinstance(batch, (tuple, list)):
                    inputs = batch[0]
                    targets = batch[1] if len(batch) > 1 else None
                else:
                    inputs, targets = batch, None
                inputs = inputs.to(dev)
                self.model.zero_grad()
                if empirical and targets is not None:
                    targets = targets.to(inputs.device)
                    loss = nn.functional.mse_loss(self.model(inputs), targets)
                    loss.backward()
                else:
                    outputs = self.model(inputs)
                    log_prob = -0.5 * (outputs ** 2).sum() if outputs.dim() > 1 else outputs.sum()
                    log_prob.backward()
                grads = [None if p.grad is None else p.grad.detach().contiguous() for p in params]
                nv.ewc_fisher_accum(fisher.flat, grads, numels, 1.0)
                used += inputs.size(0)
        finally:
            for m, sync in synced:
                m._grad_sync = sync
        if self.process_group is not None:
            import torch.distributed as dist
            cnt = torch.tensor([float(used)], device=dev)
            dist.all_reduce(fisher.flat, group=self.process_group)
            dist.all_reduce(cnt, group=self.process_group)
            used = int(cnt.item())
        nv.ewc_axpby(fisher.flat, None, 1.0 / max(used, 1), 0.0)
        self._last_fisher = fisher
        return fisher.views

    def register_task(self, task_id: int, dataloader, num_samples: Optional[int] = None) -> None:
        """Reference ewc.py:151-193."""
        self.compute_fisher(dataloader, num_samples)
        fisher: _FlatState = self._last_fisher
        star = _FlatState(fisher.layout, fisher.flat.device, zero=False)
        with torch.no_grad():
            for name, param in self._get_params():
                star.views[name].copy_(param.data)
        if self.mode == "separate":
            self._tasks[task_id] = (fisher, star)
            self.task_fisher[task_id] = fisher.views
            self.task_optpar[task_id] = star.views
        elif self.mode == "online":
            if self._online is None or len(self.fisher_dict) == 0:
                self._online = (fisher, star)
            else:
                running = self._online[0]
                nv.ewc_axpby(running.flat, fisher.flat, float(self.decay), float(1 - self.decay))
                self._online = (running, star)
            self.fisher_dict = self._online[0].views
            self.optpar_dict = self._online[1].views
        self.num_tasks += 1

    def _adopt_loaded_state(self) -> None:
        """Rebuild the flat stores after ``load_state_dict`` (whose dicts hold CPU tensors)."""
        dev = self._device()
        layout = self._layout()
        if self.fisher_dict:
            pair = (_FlatState.from_dict(layout, self.fisher_dict, dev),
                    _FlatState.from_dict(layout, self.optpar_dict, dev))
            self._online = pair
            self.fisher_dict, self.optpar_dict = pair[0].views, pair[1].views
        self._tasks = {}
        for t in self.task_fisher:
            pair = (_FlatState.from_dict(layout, self.task_fisher[t], dev),
                    _FlatState.from_dict(layout, self.task_optpar[t], dev))
            self._tasks[t] = pair
            self.task_fisher[t], self.task_optpar[t] = pair[0].views, pair[1].views

    def penalty(self, model: Optional[nn.Module] = None):
        """lambda/2 * sum_i F_i (theta_i - theta*_i)^2 (reference ewc.py:195-232).

        Returns the Python float ``0.0`` before any task is registered, like the reference."""
        if model is None:
            model = self.model
        if self.mode == "separate":
            states = [self._tasks[t] for t in self.task_fisher if t in self._tasks]
        else:
            states = [self._online] if (self._online is not None and len(self.fisher_dict) > 0) else []
        if not states:
            return self.ewc_lambda / 2 * 0.0
        names = [n for n, _, _ in states[0][0].layout]
        by_name = dict(model.named_parameters())          # parameters are matched by NAME (ewc.py:226-227)
        params = [by_name[n] for n in names]
        if not params[0].is_cuda:
            raise RuntimeError("nerve_cl_b200.EWC.penalty: model parameters must be on CUDA")
        return _PenaltyFn.apply(model, self.ewc_lambda / 2.0, states, *params)

    def get_importance_stats(self) -> Dict[str, Dict[str, float]]:
        """Reference ewc.py:234-257."""
        if self.mode == "online":
            fisher = self.fisher_dict
        else:
            fisher = {}
            for tf in self.task_fisher.values():
                for name, f in tf.items():
                    fisher[name] = f.clone() if name not in fisher else fisher[name] + f
        return {name: {"mean": f.mean().item(), "max": f.max().item(), "std": f.std().item(),
                       "nonzero": (f > 0).float().mean().item()} for name, f in fisher.items()}

    def state_dict(self) -> Dict:
        """Reference ewc.py:259-275 (CPU copies, same keys)."""
        cpu = lambda d: {k: v.detach().cpu().clone() for k, v in d.items()}  # noqa: E731
        return {
            "ewc_lambda": self.ewc_lambda, "mode": self.mode, "decay": self.decay, "num_tasks": self.num_tasks,
            "fisher_dict": cpu(self.fisher_dict), "optpar_dict": cpu(self.optpar_dict),
            "task_fisher": {t: cpu(f) for t, f in self.task_fisher.items()},
            "task_optpar": {t: cpu(o) for t, o in self.task_optpar.items()},
        }

    def load_state_dict(self, state: Dict) -> None:
        """Reference ewc.py:277-287; the tensors are re-homed into flat device buffers."""
        self.ewc_lambda = state["ewc_lambda"]
        self.mode = state["mode"]
        self.decay = state["decay"]
        self.num_tasks = state["num_tasks"]
        self.fisher_dict = dict(state["fisher_dict"])
        self.optpar_dict = dict(state["optpar_dict"])
        self.task_fisher = {t: dict(f) for t, f in state["task_fisher"].items()}
        self.task_optpar = {t: dict(o) for t, o in state["task_optpar"].items()}
        self._online = None
        self._adopt_loaded_state()


class OnlineEWC(EWC):
    """Reference ewc.py:290-303."""

    def __init__(self, model: nn.Module, ewc_lambda: float = 5000.0, decay: float = 0.999):
        super().__init__(model, ewc_lambda, mode="online", decay=decay)
