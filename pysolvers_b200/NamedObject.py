"""Import-path compatibility: ``PySolvers.NamedObject``."""
from .core import NamedObject  # noqa: F401
