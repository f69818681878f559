"""Solver base classes and ``mvmult`` (host side).

API mirror of PySolvers/Linear/LinearSolver.py:8-42 and
PySolvers/Linear/IterativeLinearSolver.py:31-106.  ``mvmult`` keeps the
reference signature (scipy matrix or 2-D ndarray times numpy vector -> numpy
vector) but the product is computed by the CSR SpMV kernel on the device.
"""
from abc import ABC, abstractmethod

import numpy as np
import torch

from ..core import CommonSolverArgs, IterativeSolver, NamedObject
from ..device import DeviceCSR, to_device, to_host
from .precond import IdentityPreconditionerType


class LinearSolverType(ABC, NamedObject):
    """Factory interface: ``makeSolver(name=None)`` returns a solver object."""

    def __init__(self, name=''):
        NamedObject.__init__(self, name=name)

    @abstractmethod
    def makeSolver(self, name=None):
        ...


class LinearSolver(ABC, NamedObject):
    """``solve(A, b) -> SolveStatus`` plus the matrix-freeze flag."""

    def __init__(self, name=''):
        NamedObject.__init__(self, name=name)
        self._matrixFrozen = False

    @abstractmethod
    def solve(self, A, b):
        ...

    def freezeMatrix(self):
        self._matrixFrozen = True

    def unfreezeMatrix(self):
        self._matrixFrozen = False

    def matrixFrozen(self):
        return self._matrixFrozen


class IterativeLinearSolverType(LinearSolverType):
    """Factory base for iterative solvers: carries the shared control object
    and the preconditioner factory.  The defaults are def-time singletons, as
    in the reference."""

    def __init__(self, control=CommonSolverArgs(),
                 precond=IdentityPreconditionerType(), name=''):
        super().__init__(name)
        self._control = control
        self._precondType = precond

    def precond(self):
        return self._precondType

    def control(self):
        return self._control


class IterativeLinearSolver(LinearSolver, IterativeSolver):
    def __init__(self, control, precond=IdentityPreconditionerType(), name=''):
        LinearSolver.__init__(self, name=name)
        IterativeSolver.__init__(self, control=control, name=name)
        self._precondType = precond
        self._precFrozen = False

    def precondType(self):
        return self._precondType

    def setTolerance(self, tau):
        IterativeSolver.setTolerance(self, tau)

    def freezePrec(self):
        self._precFrozen = True

    def unfreezePrec(self):
        self._precFrozen = False

    def precFrozen(self):
        return self._precFrozen

    # -- device plumbing shared by the Krylov solvers ------------------------
    @staticmethod
    def _check_system(A, b):
        n, nc = A.shape
        assert n == nc
        assert n == len(b)
        return n

    @staticmethod
    def _to_host(t, like):
        out = to_host(t)
        return out if out.dtype == like.dtype else out.astype(like.dtype)


def mvmult(A, x):
    """y = A x through the device SpMV; returns a numpy vector."""
    dA = A if isinstance(A, DeviceCSR) else DeviceCSR(A)
    y = dA.matvec(to_device(x))
    return to_host(y)
