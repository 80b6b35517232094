"""B200-native lattice-Boltzmann engine for BADChIMP's hot loop (host-side Python mirror).

geometry : voxel array -> the reference's per-rank tables (vtklb.py / LBnodes.h / LBbndmpi.h numbering)
cases    : host-side setup of the three target mains (std_case, std_one_phase, twophase)
checkpoint : the reference's .lblbf / .lbsca / .lbvec restart files
capi     : ctypes binding of the C-ABI in include/chimp_b200.h (libchimp_b200.so, CUDA only)
"""
from . import geometry  # noqa: F401
from . import cases  # noqa: F401
from . import capi  # noqa: F401
from . import checkpoint  # noqa: F401

__all__ = ["geometry", "cases", "capi", "checkpoint"]
