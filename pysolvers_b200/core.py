"""Control, reporting and result protocol shared by all solvers (host side).

Mirrors the public surface of the reference's core layer so user scripts run
unchanged:

* ``CommonSolverArgs``  <- PySolvers/IterativeSolver.py:25-57
* ``IterativeSolver``   <- PySolvers/IterativeSolver.py:62-155
* ``SolveStatus``       <- PySolvers/SolveStatus.py:8-56
* ``NamedObject``       <- PySolvers/NamedObject.py:2-11

The ``iters`` / ``resid`` / ``msg`` conventions of the three exits (converged,
breakdown, maxiter) are reproduced exactly; the
device loop reports a 0-based iteration index k and the helpers below turn it
into the reference's numbers.
"""
import numpy.linalg as npla


def _is_cuda_tensor(x):
    return type(x).__module__.startswith('torch') and getattr(x, 'is_cuda', False)


def _device_norm2(x):
    import ctypes as C
    import torch
    from . import _native as nat
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    nat.check(nat.lib().psb_dot(x.numel(), C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()),
                                C.c_void_p(out.data_ptr()),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'psb_dot')
    return float(out.item()) ** 0.5


class Tab:
    """Indentation prefix for nested solver output (the reference uses the
    un-vendored PyTab package for this)."""
    _level = 0

    def indent(self):
        Tab._level += 1

    def unindent(self):
        if Tab._level > 0:
            Tab._level -= 1

    def __str__(self):
        return '  ' * Tab._level

    def __format__(self, spec):
        return format(str(self), spec)


class NamedObject:
    def __init__(self, name=''):
        self._name = name

    def name(self):
        return self._name


class SolveStatus:
    """Outcome of a linear or nonlinear solve: success flag, solution,
    residual norm, iteration count and an optional message."""
    __slots__ = ('_success', '_soln', '_resid', '_iters', '_msg')

    def __init__(self, success, soln, resid, iters, msg=None):
        self._success, self._soln, self._resid = success, soln, resid
        self._iters, self._msg = iters, msg

    def success(self):
        return self._success

    def soln(self):
        return self._soln

    def resid(self):
        return self._resid

    def iters(self):
        return self._iters

    def msg(self):
        return self._msg

    def __str__(self):
        return 'SolverState(success={}, resid={}, iters={})'.format(
            self._success, self._resid, self._iters)


class CommonSolverArgs:
    """Knobs common to the iterative solvers: maxiter (100), failOnMaxiter,
    relative tolerance tau (1e-8), the norm callable, and print controls.

    The GPU path evaluates the Euclidean norm on the device; any ``norm`` other
    than ``numpy.linalg.norm`` makes the device solvers raise
    NotImplementedError (there is no CPU fallback)."""

    def __init__(self, maxiter=100, failOnMaxiter=True, tau=1.0e-8,
                 norm=npla.norm, showIters=True, showFinal=True, interval=1):
        self.maxiter = maxiter
        self.failOnMaxiter = failOnMaxiter
        self.tau = tau
        self.norm = norm
        self.showIters = showIters
        self.showFinal = showFinal
        self.interval = interval


class IterativeSolver(NamedObject):
    """Bookkeeping base: accessors for the control object, per-iteration and
    final reporting, and the three ``handle*`` exits that build SolveStatus."""

    def __init__(self, control, name=''):
        NamedObject.__init__(self, name)
        self._control = control

    # -- control accessors -------------------------------------------------
    def maxiter(self):
        return self._control.maxiter

    def failOnMaxiter(self):
        return self._control.failOnMaxiter

    def tau(self):
        return self._control.tau

    def setTolerance(self, tau):
        self._control.tau = tau

    def norm(self, x):
        if self._control.norm is npla.norm and _is_cuda_tensor(x):
            return _device_norm2(x)             # same 2-norm, evaluated where the vector lives
        return self._control.norm(x)

    def _require_euclidean_norm(self):
        if self._control.norm is not npla.norm:
            raise NotImplementedError(
                'the device solvers support only the default 2-norm '
                '(numpy.linalg.norm); got %r' % (self._control.norm,))

    # -- reporting ---------------------------------------------------------
    def reportIter(self, iter, normR, normR0):
        c = self._control
        if c.showIters and iter % c.interval == 0:
            print('%s%s iter=%7d ||r||=%12.5g ||r||/r0=%12.5g' % (
                Tab(), self.name(), iter, normR, normR / normR0))

    def _final(self, verdict, iters, normR, normB):
        if self._control.showFinal:
            print('%s%s solve %s: iters=%7d, ||r||/r0=%12.5g' % (
                Tab(), self.name(), verdict, iters, normR))

    def reportSuccess(self, iter, normR, normB):
        rel = normR / normB if normR != 0 else normR
        self._final('succeeded', iter, rel, normB)

    def reportFailure(self, iter, normR, normB):
        self._final('FAILED', iter, normR / normB, normB)

    def reportBreakdown(self, msg=''):
        if self._control.showFinal:
            print('%s%s solve broke down: %s' % (Tab(), self.name(), msg))

    # -- exits ---------------------------------------------------------------
    def handleConvergence(self, iter, x, normR, normB):
        self.reportSuccess(iter + 1, normR, normB)
        return SolveStatus(True, x, normR, iter + 1)

    def handleBreakdown(self, iter, msg):
        self.reportBreakdown(msg=msg)
        return SolveStatus(False, None, None, iter, msg)

    def handleMaxiter(self, iter, x, normR, normB):
        if self.failOnMaxiter():
            self.reportFailure(iter, normR, normB)
            return SolveStatus(False, x, normR, iter, 'failure to converge')
        self.reportSuccess(iter + 1, normR, normB)
        return SolveStatus(True, x, normR, iter)
