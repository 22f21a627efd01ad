"""sculptmate_b200 -- B200-native (sm_100a) implementation of SculptMate's
TripoSR ``TSR.extract_mesh`` hot path: triplane query + NeRFMLP on tcgen05 tensor
cores and a deterministic CUDA marching-cubes pipeline, behind the reference's own
Python call signatures.  See DESIGN.md / INTEGRATION.md.

Layout
  csrc/        CUDA kernels + the C ABI (include/sculptmate_b200.h)
  _capi.py     ctypes binding of the C ABI
  runtime.py   tensor-level wrappers (torch = device memory + streams only)
  tsr/         host-side mirror of the reference interface for this path
  dist.py      x-slab sharding across the GPUs of one node
"""
from . import _capi  # noqa: F401

__version__ = "0.1.0"
__all__ = ["_capi", "runtime", "tsr", "dist"]
