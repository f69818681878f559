"""pysolvers_b200 -- B200-native drop-in for the solve phase of PySolvers.

``import pysolvers_b200 as PySolvers`` gives the reference's package surface
(PySolvers/__init__.py:1-3): ``PySolvers.Linear``, ``PySolvers.Nonlinear``,
``PySolvers.CommonSolverArgs``.  All arithmetic on the solve path runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/pysolv_b200.h.
"""
from .core import CommonSolverArgs, SolveStatus  # noqa: F401
from . import Linear  # noqa: F401
from . import Nonlinear  # noqa: F401
