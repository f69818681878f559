"""Import-path compatibility: ``PySolvers.SolveStatus``."""
from .core import SolveStatus  # noqa: F401
