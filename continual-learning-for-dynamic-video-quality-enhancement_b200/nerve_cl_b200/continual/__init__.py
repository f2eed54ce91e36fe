"""Mirror of ``nerve_cl.continual`` for the hot path (EWC and Synaptic Intelligence)."""
from .ewc import EWC, OnlineEWC
from .si import SynapticIntelligence

__all__ = ["EWC", "OnlineEWC", "SynapticIntelligence"]
