"""CPU restatement of the EWC arithmetic -- TEST INFRASTRUCTURE ONLY (see sr_oracle.py header).

Follows nerve_cl/continual/ewc.py of the reference:

* ``fisher_from_batches``   compute_fisher, ewc.py:73-149  (sum over batches of (grad of the
  batch-mean loss)^2, divided by the number of *samples* -- not a per-sample Fisher)
* ``consolidate``           register_task online branch, ewc.py:181-191
* ``penalty`` / ``penalty_grad``  ewc.py:195-232 and its autograd derivative lambda*F*(theta-theta*)

Everything is plain numpy over flat fp32 vectors (the layout the CUDA kernels use), accumulating the
penalty in float64 so the checker is strictly more accurate than either implementation under test.
Pinned against the live reference by tests/golden/make_golden.py -> tests/golden/ewc_linear.npz.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np


def fisher_from_batches(batch_grads: Iterable[np.ndarray], batch_sizes: Sequence[int]) -> np.ndarray:
    """ewc.py:139-147.  ``batch_grads[i]`` is the flat gradient of batch i's mean loss."""
    acc = None
    for g in batch_grads:
        g = np.asarray(g, dtype=np.float32)
        sq = g * g
        acc = sq if acc is None else (acc + sq).astype(np.float32)
    n = max(int(sum(batch_sizes)), 1)
    return (acc / np.float32(n)).astype(np.float32)


def consolidate(running: np.ndarray, new: np.ndarray, decay: float) -> np.ndarray:
    """ewc.py:186-190: F <- decay*F + (1-decay)*F_new, in fp32 like the reference."""
    d = np.float32(decay)
    return (d * running.astype(np.float32) + (np.float32(1.0) - d) * new.astype(np.float32)).astype(np.float32)


def penalty(theta: np.ndarray, fisher: np.ndarray, star: np.ndarray, ewc_lambda: float) -> float:
    """ewc.py:226-232: lambda/2 * sum F*(theta-theta*)^2 (float64 accumulation)."""
    d = theta.astype(np.float64) - star.astype(np.float64)
    return float(ewc_lambda) / 2.0 * float(np.sum(fisher.astype(np.float64) * d * d))


def penalty_grad(theta: np.ndarray, fisher: np.ndarray, star: np.ndarray, ewc_lambda: float) -> np.ndarray:
    """d penalty / d theta = lambda * F * (theta - theta*)."""
    return (np.float32(ewc_lambda) * fisher.astype(np.float32)
            * (theta.astype(np.float32) - star.astype(np.float32))).astype(np.float32)


def penalty_separate(theta: np.ndarray, fishers: Sequence[np.ndarray], stars: Sequence[np.ndarray],
                     ewc_lambda: float) -> float:
    """ewc.py:213-223 ('separate' mode): lambda/2 * sum over tasks of sum F_t (theta - theta*_t)^2."""
    return sum(penalty(theta, f, s, ewc_lambda) for f, s in zip(fishers, stars))


def penalty_separate_grad(theta: np.ndarray, fishers: Sequence[np.ndarray], stars: Sequence[np.ndarray],
                          ewc_lambda: float) -> np.ndarray:
    out = np.zeros_like(theta, dtype=np.float32)
    for f, s in zip(fishers, stars):
        out = (out + penalty_grad(theta, f, s, ewc_lambda)).astype(np.float32)
    return out


# ---- Synaptic Intelligence (ewc.py:306-379) over flat fp32 vectors ----
def si_update(W: np.ndarray, p_old: np.ndarray, theta: np.ndarray, grad: np.ndarray):
    """ewc.py:342-352: W += -grad * (theta - p_old); p_old = theta.  Returns (W, p_old)."""
    delta = theta.astype(np.float32) - p_old.astype(np.float32)
    W = (W.astype(np.float32) + (-grad.astype(np.float32)) * delta).astype(np.float32)
    return W, theta.astype(np.float32).copy()


def si_register(W: np.ndarray, p_old: np.ndarray, omega: np.ndarray, theta: np.ndarray, damping: float):
    """ewc.py:354-366: omega += W / ((theta - p_old)^2 + damping); W = 0; p_old = theta.  Returns (W, p_old, omega)."""
    delta = theta.astype(np.float32) - p_old.astype(np.float32)
    denom = (delta * delta + np.float32(damping)).astype(np.float32)
    omega = (omega.astype(np.float32) + W.astype(np.float32) / denom).astype(np.float32)
    return np.zeros_like(W, dtype=np.float32), theta.astype(np.float32).copy(), omega


def si_penalty(theta: np.ndarray, omega: np.ndarray, p_old: np.ndarray, si_lambda: float) -> float:
    """ewc.py:368-379: si_lambda * sum omega (theta - p_old)^2 (float64 accumulation)."""
    d = theta.astype(np.float64) - p_old.astype(np.float64)
    return float(si_lambda) * float(np.sum(omega.astype(np.float64) * d * d))


def si_penalty_grad(theta: np.ndarray, omega: np.ndarray, p_old: np.ndarray, si_lambda: float) -> np.ndarray:
    return (np.float32(2.0 * si_lambda) * omega.astype(np.float32)
            * (theta.astype(np.float32) - p_old.astype(np.float32))).astype(np.float32)
