"""Incomplete-Cholesky and ILUT preconditioners (host setup, device apply).

Same factories and classes as PySolvers/Linear/ICPreconditioner.py:20-63 and
PySolvers/Linear/ILUTPreconditioner.py:10-78.  As north_star prescribes, the
SETUP is the reference's own: SuperLU's ILUTP through
``scipy.sparse.linalg.spilu`` on the host, with the reference's arguments.  The
factors are then level-analysed and uploaded once (psb_trsv_create), and every
APPLICATION is two sync-free sparse triangular solves on the device.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .. import _native as nat
from ..device import DeviceCSR, DevicePrec, DeviceSplitLU, DeviceTrsv, current_stream_ptr
from .precond import (LeftPreconditioner, Preconditioner, PreconditionerType,
                      RightPreconditioner)


def _host_matrix(A):
    if isinstance(A, DeviceCSR):
        return A.to_scipy()          # setup runs on the host (SuperLU), as in the reference
    if not sp.issparse(A):
        raise TypeError('IC / ILUT need a scipy sparse matrix (the reference '
                        'calls A.tocsc()); got %r' % type(A))
    return A


class RightIC(PreconditionerType):
    def __init__(self, drop_tol=0.001, fill_factor=15):
        self.drop_tol = drop_tol
        self.fill_factor = fill_factor

    def form(self, A):
        return ICRightPreconditioner(A, drop_tol=self.drop_tol,
                                     fill_factor=self.fill_factor)


class ICRightPreconditioner(RightPreconditioner):
    """M^-1 = L^-T L^-1 with L the incomplete Cholesky factor obtained from
    SuperLU's ILU with pivoting and column permutation switched off
    (ICPreconditioner.py:45-56): L^T = diag(U)^(-1/2) U."""

    def __init__(self, A, drop_tol=0.001, fill_factor=15):
        A = _host_matrix(A)
        ilu = spla.spilu(A.tocsc(), drop_tol=drop_tol, fill_factor=fill_factor,
                         diag_pivot_thresh=0.0, options={'ColPerm': 'NATURAL'})
        U = ilu.U
        n = A.shape[0]
        scale = np.reciprocal(np.sqrt(U.diagonal()))
        del ilu
        # row scaling written as the reference writes it (dia * csc) so that the
        # stored order and every rounded product are the same
        self._Lt = (sp.dia_matrix((scale, [0]), shape=(n, n)) * U).tocsr()
        self._L = self._Lt.transpose().tocsr()
        self._dL = DeviceTrsv(self._L, lower=True)
        self._dLt = DeviceTrsv(self._Lt, lower=False)
        h = C.c_void_p()
        nat.check(nat.lib().psb_ic_create(self._dL.handle, self._dLt.handle, C.byref(h)),
                  'psb_ic_create')
        self._dev = DevicePrec(h, n, keep=(self._dL, self._dLt))

    def applyRight(self, vec):
        return self._dev.apply_host(vec)

    def right_device_handle(self):
        return self._dev.handle

    def device_prec(self):
        return self._dev


class LeftILUT(PreconditionerType):
    def __init__(self, drop_tol=0.001, fill_factor=15):
        self.drop_tol = drop_tol
        self.fill_factor = fill_factor

    def form(self, A):
        return LeftILUTPreconditioner(A, drop_tol=self.drop_tol,
                                      fill_factor=self.fill_factor)


class RightILUT(PreconditionerType):
    def __init__(self, drop_tol=0.001, fill_factor=15):
        self.drop_tol = drop_tol
        self.fill_factor = fill_factor

    def form(self, A):
        return RightILUTPreconditioner(A, drop_tol=self.drop_tol,
                                       fill_factor=self.fill_factor)


class ILUTPreconditioner(Preconditioner):
    """SuperLU ILUTP factorisation (default COLAMD column permutation,
    diag_pivot_thresh=0; ILUTPreconditioner.py:51-53), applied on the device as
    x = Pc U^-1 L^-1 Pr v."""

    def __init__(self, A, drop_tol=0.001, fill_factor=15):
        A = _host_matrix(A)
        self._ILU = spla.spilu(A.tocsc(), drop_tol=drop_tol,
                               fill_factor=fill_factor, diag_pivot_thresh=0.0)
        # applied with the dense trailing blocks of L and U inverted once (device.DeviceSplitLU):
        # systems of up to 8 192 rows become two triangular GEMVs, for DH-15 the dependency levels
        # left for the sparse triangular solves halve (993 / 891 -> 476 / 456)
        self._dev = DeviceSplitLU(self._ILU)

    def ILU(self):
        return self._ILU

    def device_prec(self):
        return self._dev


class LeftILUTPreconditioner(ILUTPreconditioner, LeftPreconditioner):
    def __init__(self, A, drop_tol=0.001, fill_factor=15):
        ILUTPreconditioner.__init__(self, A, drop_tol=drop_tol, fill_factor=fill_factor)

    def applyLeft(self, vec):
        return self._dev.apply_host(vec)


class RightILUTPreconditioner(ILUTPreconditioner, RightPreconditioner):
    def __init__(self, A, drop_tol=0.001, fill_factor=15):
        ILUTPreconditioner.__init__(self, A, drop_tol=drop_tol, fill_factor=fill_factor)

    def applyRight(self, vec):
        return self._dev.apply_host(vec)

    def right_device_handle(self):
        return self._dev.handle
