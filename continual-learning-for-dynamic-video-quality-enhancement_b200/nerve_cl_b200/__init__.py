"""nerve_cl_b200 -- B200-native (sm_100a) implementation of NERVE-CL's enhancement hot path.

Mirrors the reference package layout for the path it replaces::

    from nerve_cl_b200.models import SuperResolutionNet          # nerve_cl.models.SuperResolutionNet
    from nerve_cl_b200.continual import EWC, OnlineEWC            # nerve_cl.continual.EWC

Everything executes in ``libnervecl.so`` (hand-written CUDA, C ABI in ``include/nervecl.h``) through
``torch.ops.nervecl``.  Importing the package loads the library and fails loudly if it is missing.
"""
from . import _lib

_lib.load()          # no library => RuntimeError here, never a silent fallback

from . import ops  # noqa: E402,F401
from .models import SuperResolutionNet, LightweightSuperResolution  # noqa: E402,F401
from .continual import EWC, OnlineEWC, SynapticIntelligence  # noqa: E402,F401

__version__ = "0.1.0"
