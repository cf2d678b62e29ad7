"""red-diffeq_b200 -- B200 (sm_100a) implementation of RED-DiffEq's FD wave-solver hot path.

Public surface (same names as the reference package exports for this path):
    FWIForward, v_normalize, v_denormalize, s_normalize_none
Import it as ``red_diffeq_b200`` (the shim red_diffeq_b200.py at the repo root maps the importable
name onto this directory, whose name has a hyphen).
"""
from .core.inversion import InversionEngine
from .regularization.diffusion import REDDiffEq
from .solvers.pde import FWIForward
from .solvers.sharding import ShardedFWIForward
from .utils.data_trans import add_noise_to_seismic, missing_trace, s_normalize_none, v_denormalize, v_normalize

__version__ = "0.1.0"
__all__ = ["FWIForward", "ShardedFWIForward", "InversionEngine", "REDDiffEq", "v_normalize", "v_denormalize", "s_normalize_none",
           "add_noise_to_seismic", "missing_trace"]
