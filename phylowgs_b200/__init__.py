"""phylowgs_b200 — B200-native (sm_100a CUDA, C ABI, ctypes) replacement for the PhyloWGS
Metropolis-Hastings / SSM-likelihood hot path.  See DESIGN.md and include/pwgs_b200.h."""
from ._lib import SEM_MH, SEM_PY, PwgsError, device_count, fp64_peak  # noqa: F401
from .engine import Engine  # noqa: F401

__all__ = ["Engine", "SEM_MH", "SEM_PY", "PwgsError", "device_count", "fp64_peak"]
