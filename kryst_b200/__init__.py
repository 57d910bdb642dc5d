"""kryst_b200 — B200-native Krylov hot path behind kryst's trait API.

Hand-written sm_100a CUDA kernels (kryst_b200/csrc) behind a C ABI (include/kryst_b200.h); this
package is the host-side mirror of the reference's operator / preconditioner / solver interface.
There is no CPU fallback: importing fails when libkryst_b200.so has not been built.
"""
from . import _ffi
from .api import (AdditiveSchwarz, BiCgStabSolver, BlockJacobiIlu0, CgNormType, Context, DeviceCsr, FactorError, FgmresSolver, GmresSolver, Ilu0,
                  IndefiniteMatrix, IndefinitePreconditioner, Jacobi, KError, PcgSolver, Preconditioning, SolveError,
                  PC, SolveStats, Unsupported, ZeroPivot, default_context, get_history, partition_range)
from . import stencils
from . import mmio
from .context import KspContext, SolverKind

_ffi.lib()   # fail loudly at import time if the CUDA library is missing

__all__ = ["AdditiveSchwarz", "PC", "get_history", "BiCgStabSolver", "BlockJacobiIlu0", "CgNormType", "Context", "DeviceCsr", "FactorError", "FgmresSolver", "GmresSolver", "Ilu0",
           "IndefiniteMatrix", "IndefinitePreconditioner", "Jacobi", "KError", "PcgSolver", "Preconditioning", "SolveError",
           "SolveStats", "Unsupported", "ZeroPivot", "default_context", "partition_range", "stencils", "mmio", "KspContext", "SolverKind"]
