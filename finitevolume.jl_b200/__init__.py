"""finitevolume.jl_b200 -- B200-native (sm_100a) drop-in for the assemble -> solve hot path of
madsjulia/FiniteVolume.jl: same function names and argument meaning as the reference's Julia
API, every heavy step executed by hand-written CUDA kernels in libfvb200.so (C ABI in
include/fvb200.h).  There is no CPU fallback: importing works anywhere, calling requires the
built library and a CUDA device.

The directory name contains a dot, so load it with `__graft_entry__.load_package()` (which
registers it as module `fvb200`) rather than a plain import statement.
"""
from . import _lib, jld
from ._lib import FVBError, LIB_PATH
from .api import (ConvergenceHistory, DEFAULT_MAXITER, DeviceArray, MultiSystem, SQRT_EPS, SparseMatrixCSC, System, assembleA, assembleb,
                  freenodes2nodes, getfreenodes, solvediffusion)
from .grid import grid_sizes, nodehycos2neighborhycos, regulargrid
from .transient import (adaptivebackwardeulerstep, adjointintegrate, adjointintegrate_generic, backwardeulerintegrate,
                        backwardeulerintegrate_generic, gradientintegrate_generic, fixedbackwardeulerstep, getadjointfunctions,
                        getcontinuoussolution, gradientintegrate, integrate_g, integratedfdplambda)

__all__ = [
    "FVBError", "LIB_PATH", "DeviceArray", "ConvergenceHistory", "DEFAULT_MAXITER", "SQRT_EPS", "SparseMatrixCSC", "System", "MultiSystem",
    "assembleA", "assembleb", "freenodes2nodes", "getfreenodes", "solvediffusion", "grid_sizes",
    "nodehycos2neighborhycos", "regulargrid", "adaptivebackwardeulerstep", "adjointintegrate",
    "backwardeulerintegrate", "backwardeulerintegrate_generic", "adjointintegrate_generic", "gradientintegrate_generic", "fixedbackwardeulerstep", "getadjointfunctions",
    "getcontinuoussolution", "gradientintegrate", "integrate_g", "integratedfdplambda", "jld",
]
