"""Default direct solver passthrough (PySolvers/Linear/DefaultDirectSolver.py:21-74).

Out of scope as a kernel target: a direct
factorisation is not on the Krylov hot path.  Kept for API completeness --
NewtonSolver's default linear solver is DefaultDirect -- as a thin call into
scipy / numpy exactly like the reference."""
import numpy as np
import numpy.linalg as npla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from ..core import SolveStatus
from .base import LinearSolver, LinearSolverType


class DefaultDirect(LinearSolverType):
    def __init__(self, name='Default direct'):
        super().__init__(name=name)

    def makeSolver(self, name=None):
        return DefaultDirectSolver(name=self.name() if name is None else name)


class DefaultDirectSolver(LinearSolver):
    def __init__(self, name='Default direct'):
        super().__init__(name=name)

    def solve(self, A, b):
        n, nc = A.shape
        assert n == nc
        assert n == len(b)
        try:
            if sp.isspmatrix(A):
                x = spla.spsolve(A, b)
            elif isinstance(A, np.ndarray):
                x = npla.solve(A, b)
            else:
                return SolveStatus(False, None, None, None,
                                   'Input to solver [%s] not numpy or scipy' % self.name())
            return SolveStatus(True, x, None, None, '%s solve succeeded' % self.name())
        except Exception as ex:
            return SolveStatus(False, None, None, None,
                               '{} solve failed: {}'.format(self.name(), ex))
