"""Preconditioner protocol (host side) -- the drop-in boundary of SURVEY.md 8b.

API mirror of PySolvers/Linear/Preconditioner.py:3-68 and
PySolvers/Linear/PreconditionerType.py:4-19: a factory ``form(A)`` returns an
object with ``applyLeft(vec)`` / ``applyRight(vec)``.

Device contract added by this implementation: the Krylov solvers keep the
whole loop in HBM, so instead of calling ``applyRight`` per iteration they ask
the preconditioner for ``right_device_handle()`` -- a ``psb_prec_t`` (or None
for "applyRight is the identity").  ``applyLeft/applyRight`` on numpy vectors
still work (upload, apply on the device, download).  A user-defined
preconditioner without a device handle makes the solvers raise
NotImplementedError: there is no CPU fallback on the solve path.
"""
from abc import ABC, abstractmethod


class Preconditioner(ABC):
    """Two-sided preconditioner interface."""

    def __init__(self):
        pass

    @abstractmethod
    def applyLeft(self, vec):
        ...

    @abstractmethod
    def applyRight(self, vec):
        ...

    def right_device_handle(self):
        raise NotImplementedError(
            '%s has no device implementation; the GPU solvers cannot call a '
            'host applyRight() per iteration' % type(self).__name__)


class GenericPreconditioner(Preconditioner):
    """One operator ``apply`` used from either side."""

    @abstractmethod
    def apply(self, vec):
        ...

    def applyLeft(self, vec):
        return self.apply(vec)

    def applyRight(self, vec):
        return self.apply(vec)


class LeftPreconditioner(Preconditioner):
    """Acts from the left only; from the right it is the identity.  PCG and
    GMRES call only ``applyRight``, so a left preconditioner is a silent no-op
    there -- reproduced."""

    def applyRight(self, vec):
        return vec

    def right_device_handle(self):
        return None


class RightPreconditioner(Preconditioner):
    def applyLeft(self, vec):
        return vec


class IdentityPreconditioner:
    def applyLeft(self, vec):
        return vec

    def applyRight(self, vec):
        return vec

    def right_device_handle(self):
        return None


class PreconditionerType(ABC):
    @abstractmethod
    def form(self, A):
        ...


class IdentityPreconditionerType(PreconditionerType):
    def form(self, A):
        return IdentityPreconditioner()


def right_handle_of(prec):
    """psb_prec_t (ctypes void pointer) for ``prec.applyRight`` or None when it
    is the identity; raises for host-only preconditioners."""
    fn = getattr(prec, 'right_device_handle', None)
    if fn is None:
        raise NotImplementedError(
            'preconditioner %r does not provide right_device_handle(); the '
            'GPU solve path has no CPU fallback' % (prec,))
    return fn()
