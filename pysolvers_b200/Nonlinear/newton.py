"""Newton's method with backtracking line search: the host-side OUTER loop of
configuration 5 (PySolvers/Nonlinear/Newton.py:12-101, LineSearch.py:4-81,
PreconditionerFreeze.py:3-24, FuncAdapter1D.py).  It stays in Python, as in
the reference; all it does on the hot path is call ``solver.solve(J, -F)`` on
one of the device solvers."""
from abc import ABC, abstractmethod

import numpy as np

from ..core import CommonSolverArgs, IterativeSolver, Tab
from ..Linear.base import IterativeLinearSolver
from ..Linear.direct import DefaultDirect


class LineSearch(ABC):
    def __init__(self, maxsteps=15, low=0.1, alpha=0.0001, report=True):
        self._maxsteps, self._low, self._alpha = maxsteps, low, alpha
        self._report = report
        self._norm = None

    @abstractmethod
    def search(self, x0, resid, newtStep, func):
        ...

    def maxsteps(self):
        return self._maxsteps

    def alpha(self):
        return self._alpha

    def low(self):
        return self._low

    def setNorm(self, norm):
        self._norm = norm

    def norm(self, x):
        if self._norm is None:
            raise RuntimeError('Norm not set in line search')
        return self._norm(x)

    def report(self, k, t, ratio):
        if self._report:
            print('%sk=%4d t=%12.5g ||F_k||/||F_0||=%12.5g' % (Tab(), k, t, ratio))


class TrivialLinesearch(LineSearch):
    """Accepts the full step (testing only)."""

    def __init__(self, report=True):
        super().__init__(report=report)

    def search(self, x0, normF0, newtStep, func):
        x1 = x0 + newtStep
        F1 = func.evalF(x1)
        return (True, x1, F1, self.norm(F1))


class SimpleBacktrack(LineSearch):
    """Dennis & Schnabel backtracking (LineSearch.py:57-81): accept when
    ||F(x + t p)|| <= (1 - alpha t) ||F(x)||, else shrink t by
    max(0.5 / ratio, low)."""

    def __init__(self, maxsteps=10, low=0.1, alpha=0.0001, report=True):
        super().__init__(maxsteps=maxsteps, low=low, alpha=alpha, report=report)

    def search(self, x0, normF0, newtStep, func):
        t = 1.0
        x_k = F_k = normF_k = None
        for k in range(self.maxsteps()):
            x_k = x0 + t * newtStep
            F_k = func.evalF(x_k)
            normF_k = self.norm(F_k)
            ratio = normF_k / normF0
            self.report(k, t, ratio)
            if normF_k <= (1.0 - self.alpha() * t) * normF0:
                return (True, x_k, F_k, normF_k)
            t = t * max(0.5 / ratio, self.low())
        return (False, x_k, F_k, normF_k)


class PreconditionerFreeze:
    """Freezes the linear solver's preconditioner for the duration of a Newton
    solve.  The reference's un-freeze hook is mis-spelt ``__def__``
    (PreconditionerFreeze.py:23) and never runs, so the freeze PERSISTS across
    Newton solves; reproduced."""

    def __init__(self, solver, freezePrec):
        self.solver = solver
        self.freezePrec = freezePrec
        self.freeze()

    def _applies(self):
        return self.freezePrec and isinstance(self.solver, IterativeLinearSolver)

    def freeze(self):
        if self._applies():
            self.solver.freezePrec()

    def unfreeze(self):
        if self._applies():
            self.solver.unfreezePrec()

    def __def__(self):
        self.unfreeze()


class FuncAdapter1D(ABC):
    """Adapts a scalar function to the evalF / evalJ protocol: subclasses give
    ``_evalF(y)`` and ``_evalJ(y)`` (FuncAdapter1D.py:4-24)."""

    @abstractmethod
    def _evalF(self, x):
        ...

    @abstractmethod
    def _evalJ(self, x):
        ...

    def evalF(self, x):
        return np.array([self._evalF(x[0])])

    def evalJ(self, x):
        return self._evalJ(x[0]) * np.eye(1)


class NewtonSolver(IterativeSolver):
    """Inexact Newton: tau_lin = max(tolFudge ||F||/||F0||, minLinTol) unless
    fixLinTol (Newton.py:62-73); converged when ||F|| <= ||F0|| tau + tau."""

    def __init__(self, control=CommonSolverArgs(), solver=DefaultDirect(),
                 linesearch=SimpleBacktrack(), fixLinTol=False, tolFudge=0.1,
                 minLinTol=1.0e-10, freezePrec=True, name='Newton'):
        super().__init__(control, name=name)
        self.solver = solver.makeSolver()
        self.linesearch = linesearch
        self.fixLinTol, self.tolFudge, self.minLinTol = fixLinTol, tolFudge, minLinTol
        self.freezePrec = freezePrec

    def solve(self, func, xInit):
        tab = Tab()
        xCur = xInit.clone() if hasattr(xInit, 'clone') else xInit.copy()
        FCur = func.evalF(xCur)
        print('freeze prec for solver=', self.freezePrec)
        PreconditionerFreeze(self.solver, self.freezePrec)
        self.linesearch.setNorm(self.norm)
        r0 = self.norm(FCur)
        normFCur = r0
        for i in range(self.maxiter()):
            self.reportIter(i, normFCur, r0)
            if normFCur <= r0 * self.tau() + self.tau():
                return self.handleConvergence(i, xCur, normFCur, r0)
            J = func.evalJ(xCur)
            if isinstance(self.solver, IterativeLinearSolver):
                tau_lin = self.minLinTol if self.fixLinTol else max(
                    self.tolFudge * normFCur / r0, self.minLinTol)
                self.solver.setTolerance(tau_lin)
            tab.indent()
            status = self.solver.solve(J, -FCur)
            tab.unindent()
            if not status.success():
                return self.handleBreakdown(
                    i, 'solve for Newton step failed with msg={}'.format(status.msg()))
            tab.indent()
            ok, xCur, FCur, normFCur = self.linesearch.search(xCur, normFCur, status.soln(), func)
            tab.unindent()
            if not ok:
                return self.handleBreakdown(i, msg='Line search failed')
        return self.handleMaxiter(self.maxiter(), xCur, normFCur, r0)
