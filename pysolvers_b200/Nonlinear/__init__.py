"""``PySolvers.Nonlinear`` surface (PySolvers/Nonlinear/__init__.py:1-2)."""
from .newton import (NewtonSolver, SimpleBacktrack, TrivialLinesearch, LineSearch,  # noqa: F401
                     PreconditionerFreeze, FuncAdapter1D)
