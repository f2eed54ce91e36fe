"""Mirror of ``nerve_cl.continual`` for the hot path (EWC only)."""
from .ewc import EWC, OnlineEWC

__all__ = ["EWC", "OnlineEWC"]
