"""Source root of the B200-native NERVE-CL hot path.

The directory name is not an importable identifier; put this directory on ``sys.path`` (tests,
``bench.py`` and ``__graft_entry__.py`` do) and ``import nerve_cl_b200``.  ``csrc/`` holds the CUDA
kernels and the C ABI, ``nerve_cl_b200/`` the host-side mirror of the reference interface.
"""
