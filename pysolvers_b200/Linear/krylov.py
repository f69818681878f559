"""PCG and GMRES solver classes (host side) over the device-resident loops.

Same factories, constructor arguments, ``solve(A, b) -> SolveStatus`` contract,
prints and exit conventions as PySolvers/Linear/PCGSolver.py:25-142 and
PySolvers/Linear/GMRESSolver.py:27-180 -- but ``solve`` uploads the operands
once, runs the WHOLE Krylov loop in one C-ABI call (psb_pcg_solve /
psb_gmres_solve) and downloads the solution and the residual history.
``reportIter`` is then replayed for every iteration, so overriding it captures
the history exactly as it does on the reference.
"""
import ctypes as C

import numpy as np
import torch

from .. import _native as nat
from ..core import CommonSolverArgs, SolveStatus
from ..device import DeviceCSR, current_stream_ptr, ptr, require_cuda, to_device
from .base import IterativeLinearSolver, IterativeLinearSolverType
from .precond import IdentityPreconditionerType, right_handle_of


def _upload_matrix(A):
    return A if isinstance(A, DeviceCSR) else DeviceCSR(A)


def _device_norm(v):
    """Euclidean norm of a CUDA fp64 vector via the deterministic device dot."""
    out = torch.empty(1, dtype=torch.float64, device=v.device)
    nat.check(nat.lib().psb_dot(v.numel(), ptr(v), ptr(v), ptr(out),
                                current_stream_ptr()), 'psb_dot')
    return float(np.sqrt(out.item()))


class PCG(IterativeLinearSolverType):
    """Factory for PCGSolver (PCGSolver.py:25-36)."""

    def __init__(self, control=CommonSolverArgs(),
                 precond=IdentityPreconditionerType(), name='PCG'):
        super().__init__(control=control, precond=precond, name=name)

    def makeSolver(self, name=None):
        return PCGSolver(self.control(), precond=self.precond(),
                         name=self.name() if name is None else name)


class PCGSolver(IterativeLinearSolver):
    """Preconditioned conjugate gradients for SPD ``A`` (not checked).

    ``A``: scipy sparse matrix, dense 2-D ndarray (converted to CSR) or a
    ``DeviceCSR``; ``b``: 1-D float64 ndarray.  Inputs are never modified.
    The preconditioner formed on the first solve is kept on ``self.precond``
    and reused while ``freezePrec()`` is in force (PCGSolver.py:92-94).
    """

    def __init__(self, control=CommonSolverArgs(),
                 precond=IdentityPreconditionerType(), name='PCG'):
        super().__init__(control=control, precond=precond, name=name)
        self.precond = None
        self.last_history = None      # ||r_k|| of the most recent solve

    def solve(self, A, b):
        n = self._check_system(A, b)
        self._require_euclidean_norm()
        require_cuda()
        on_device = isinstance(b, torch.Tensor)      # device vector in -> device vector out
        if not on_device:
            b = np.asarray(b)
        b_d = to_device(b)
        zeros = (lambda: torch.zeros_like(b_d)) if on_device else (lambda: np.zeros_like(b))

        # b == 0 short cut, before any preconditioner work (PCGSolver.py:86-88)
        if n == 0 or _device_norm(b_d) == 0.0:
            return self.handleConvergence(0, zeros(), 0, 0)

        print('prec frozen = ', self.precFrozen())
        if self.precond is None or not self.precFrozen():
            print('building prec')
            self.precond = self.precondType().form(A)
        prec_h = right_handle_of(self.precond)

        dA = _upload_matrix(A)
        maxiter = int(self.maxiter())
        lib = nat.lib()
        wbytes = int(lib.psb_pcg_workspace_bytes(n, 1 if prec_h else 0))
        work = torch.empty(wbytes, dtype=torch.uint8, device=b_d.device)
        x_d = torch.empty(n, dtype=torch.float64, device=b_d.device)
        hist_d = torch.empty(max(maxiter, 1), dtype=torch.float64, device=b_d.device)
        res = nat.SolveResult()
        nat.check(lib.psb_pcg_solve(
            dA.handle, prec_h, ptr(b_d), ptr(x_d), ptr(work), wbytes, maxiter,
            float(self.tau()), 1 if self.failOnMaxiter() else 0, ptr(hist_d),
            C.byref(res), current_stream_ptr()), 'psb_pcg_solve')

        hist = hist_d[:res.n_hist].cpu().numpy()
        self.last_history = hist
        normB = res.norm_b
        for k in range(res.n_hist):
            self.reportIter(k, hist[k], normB)

        st = res.status
        if st == nat.TRIVIAL:
            return self.handleConvergence(0, zeros(), 0, 0)
        if st == nat.BREAKDOWN_UR:
            return self.handleBreakdown(0, 'breakdown dot(u,r)==0')
        if st == nat.BREAKDOWN_PAP:
            return self.handleBreakdown(res.k, 'breakdown dot(p, Ap)==0')
        x = x_d if on_device else self._to_host(x_d, b)
        if st == nat.CONVERGED:
            return self.handleConvergence(res.k, x, res.norm_r, normB)
        return self.handleMaxiter(res.k, x, res.norm_r, normB)


class GMRES(IterativeLinearSolverType):
    """Factory for GMRESSolver (GMRESSolver.py:27-40).  Added, optional knobs:
    ``orth`` = 'cgs2' (default, batched) or 'mgs' (the reference's
    orthogonalisation order); ``honorFreeze`` = keep the preconditioner while
    ``freezePrec()`` is in force (default False: rebuilt on every solve like the
    reference)."""

    def __init__(self, control=CommonSolverArgs(),
                 precond=IdentityPreconditionerType(), name='GMRES', orth='cgs2',
                 honorFreeze=False):
        super().__init__(control=control, precond=precond, name=name)
        self.orth = orth
        self.honorFreeze = honorFreeze

    def makeSolver(self, name=None):
        return GMRESSolver(self.control(), precond=self.precond(),
                           name=self.name() if name is None else name,
                           orth=self.orth, honorFreeze=self.honorFreeze)


class GMRESSolver(IterativeLinearSolver):
    """Right-preconditioned GMRES without restart: the Krylov dimension is
    ``maxiter`` (GMRESSolver.py:75-83).

    Reference quirks kept: the preconditioner is rebuilt on EVERY solve --
    ``freezePrec`` has no effect, because the reference assigns the formed
    preconditioner to a local variable (GMRESSolver.py:71-72); a left
    preconditioner is a no-op (only applyRight is used, :107,160).  Reference
    bugs not kept: ``self.precond`` is initialised here (the reference raises
    AttributeError without the harness workaround), and reaching maxiter
    returns ``handleMaxiter`` with the last iterate instead of the NameError
    of GMRESSolver.py:180.
    """

    def __init__(self, control=CommonSolverArgs(),
                 precond=IdentityPreconditionerType(), name='GMRES', orth='cgs2',
                 honorFreeze=False):
        super().__init__(control=control, precond=precond, name=name)
        self.precond = None
        self.orth = orth
        # added, optional: keep the formed preconditioner while ``freezePrec()`` is in force, as
        # PCGSolver does (PCGSolver.py:92-94) -- what "AMG V-cycle preconditioner reuse" in
        # configs[4] intends.  Default False = the reference's behaviour (rebuild on every solve).
        self.honorFreeze = honorFreeze
        self.last_history = None

    def solve(self, A, b):
        n = self._check_system(A, b)
        self._require_euclidean_norm()
        require_cuda()
        on_device = isinstance(b, torch.Tensor)
        if not on_device:
            b = np.asarray(b)
        b_d = to_device(b)
        zeros = (lambda: torch.zeros_like(b_d)) if on_device else (lambda: np.zeros_like(b))
        if n == 0 or _device_norm(b_d) == 0.0:
            return self.handleConvergence(0, zeros(), 0, 0)

        if self.honorFreeze and self.precond is not None and self.precFrozen():
            precond = self.precond
        else:
            precond = self.precondType().form(A)      # always rebuilt by default, see class doc
            if self.honorFreeze:
                self.precond = precond
        prec_h = right_handle_of(precond)
        orth = {'cgs2': nat.ORTH_CGS2, 'mgs': nat.ORTH_MGS}[self.orth]

        dA = _upload_matrix(A)
        maxiter = int(self.maxiter())
        lib = nat.lib()
        wbytes = int(lib.psb_gmres_workspace_bytes(n, maxiter))
        work = torch.empty(wbytes, dtype=torch.uint8, device=b_d.device)
        x_d = torch.empty(n, dtype=torch.float64, device=b_d.device)
        hist_d = torch.empty(max(maxiter, 1), dtype=torch.float64, device=b_d.device)
        res = nat.SolveResult()
        nat.check(lib.psb_gmres_solve(
            dA.handle, prec_h, ptr(b_d), ptr(x_d), ptr(work), wbytes, maxiter,
            float(self.tau()), 1 if self.failOnMaxiter() else 0, orth, ptr(hist_d),
            C.byref(res), current_stream_ptr()), 'psb_gmres_solve')

        hist = hist_d[:res.n_hist].cpu().numpy()
        self.last_history = hist
        normB = res.norm_b
        for k in range(res.n_hist):
            self.reportIter(k, hist[k], normB)
        if res.status == nat.TRIVIAL:
            return self.handleConvergence(0, zeros(), 0, 0)
        x = x_d if on_device else self._to_host(x_d, b)
        if res.status == nat.CONVERGED:
            return self.handleConvergence(res.k, x, res.norm_r, normB)
        if res.status == nat.GMRES_FALSE_CONV:
            return SolveStatus(
                success=False, iters=res.k + 1, soln=x, resid=res.norm_r,
                msg='GMRES failure: true residual %12.5g did not meet tolerance '
                    'tau=%12.5g. Recursive residual was %12.5g.' % (
                        res.norm_r, self.tau(), res.norm_r_rec))
        return self.handleMaxiter(res.k, x, res.norm_r_rec, normB)
