"""Import-path compatibility: ``PySolvers.IterativeSolver``."""
from .core import CommonSolverArgs, IterativeSolver, SolveStatus, NamedObject  # noqa: F401
