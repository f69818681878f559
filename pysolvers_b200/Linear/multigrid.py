"""Smoothed-aggregation AMG: host-side classes over the device V-cycle.

API mirror of PySolvers/Linear/{MLHierarchy.py, SmoothedAggregation.py,
ClassicSmoothers.py, VCycleManager.py, VCycleSolver.py, AMGPreconditioner.py}.
Setup (aggregation, prolongator smoothing, Galerkin products, coarse LU) runs
on the host -- the reference's own algorithm restated in O(nnz), see
amg_setup.py -- and is uploaded once; every V-cycle runs on the device
(csrc/amg.cu).

Smoother plug-in protocol (ClassicSmoothers.py:6,10: ``cls(A)``,
``.apply(f, x, nu)``): on the GPU path the smoother CLASS selects a device
smoother.  ``JacobiSmoother`` (omega = 1, undamped as in the reference),
``DampedJacobiSmoother`` (omega = 2/3) or any class with a numeric ``omega``
attribute map to the fused Jacobi sweep; ``GaussSeidelSmoother`` maps to the
triu(A) triangular solve.  Other classes raise NotImplementedError (no CPU
fallback).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

from .. import _native as nat
from ..core import CommonSolverArgs, Tab
from ..device import (DeviceCSR, DevicePrec, DeviceSplitLU, DeviceTrsv, current_stream_ptr, ptr,
                      require_cuda, to_device, to_host)
from . import amg_setup
from .base import IterativeLinearSolver, IterativeLinearSolverType
from .precond import GenericPreconditioner, PreconditionerType


# ------------------------------------------------------------------ smoothers --
class _DeviceSmoother:
    kind = None
    omega = 1.0

    def __init__(self, A):
        self.A = A
        self._dev = None

    def _build(self):
        raise NotImplementedError

    def apply(self, f, x, nu):
        """nu sweeps on the device; numpy in, numpy out."""
        if self._dev is None:
            self._build()
        return self._sweeps(np.asarray(f, dtype=np.float64), np.asarray(x, dtype=np.float64), nu)


class JacobiSmoother(_DeviceSmoother):
    """x <- x + omega D^-1 (f - A x); omega = 1 as in ClassicSmoothers.py:5-16."""
    kind = nat.SMOOTH_JACOBI
    omega = 1.0

    def _build(self):
        self.DInv = np.reciprocal(self.A.diagonal())
        self._dev = (DeviceCSR(self.A), to_device(self.DInv))

    def _sweeps(self, f, x, nu):
        dA, dinv = self._dev
        fd, a = to_device(f), to_device(x)
        b = torch.empty_like(a)
        for _ in range(nu):
            nat.check(nat.lib().psb_jacobi_sweep(dA.handle, ptr(dinv), float(self.omega), ptr(fd),
                                                 ptr(a), ptr(b), current_stream_ptr()), 'psb_jacobi_sweep')
            a, b = b, a
        return a.cpu().numpy()


class DampedJacobiSmoother(JacobiSmoother):
    """Damped Jacobi, omega = 2/3 (north_star item 5; not in the reference)."""
    omega = 2.0 / 3.0


class GaussSeidelSmoother(_DeviceSmoother):
    """x <- x + triu(A)^-1 (f - A x)  (ClassicSmoothers.py:20-36)."""
    kind = nat.SMOOTH_GS

    def _build(self):
        self.U = sp.triu(self.A).tocsr()
        self._dev = (DeviceCSR(self.A), DeviceTrsv(self.U, lower=False))

    def _sweeps(self, f, x, nu):
        dA, dU = self._dev
        fd, xd = to_device(f), to_device(x)
        r = torch.empty_like(xd)
        for _ in range(nu):
            nat.check(nat.lib().psb_spmv_residual(dA.handle, ptr(xd), ptr(fd), ptr(r),
                                                  current_stream_ptr()), 'psb_spmv_residual')
            dx = dU.solve(r)
            nat.check(nat.lib().psb_vec_add(xd.numel(), ptr(dx), ptr(xd), current_stream_ptr()), 'psb_vec_add')
        dU.check()
        return xd.cpu().numpy()


def _smoother_spec(smoother):
    """(device smoother kind, omega) selected by a smoother class."""
    kind = getattr(smoother, 'kind', None)
    if kind in (nat.SMOOTH_JACOBI, nat.SMOOTH_GS):
        return kind, float(getattr(smoother, 'omega', 1.0))
    name = getattr(smoother, '__name__', type(smoother).__name__).lower()
    if 'jacobi' in name and isinstance(getattr(smoother, 'omega', None), (int, float)):
        return nat.SMOOTH_JACOBI, float(smoother.omega)
    raise NotImplementedError(
        'smoother %r has no device implementation (use JacobiSmoother, '
        'DampedJacobiSmoother or GaussSeidelSmoother); the GPU solve path has '
        'no CPU fallback' % (smoother,))


# ------------------------------------------------------------------ hierarchy --
class MLHierarchy:
    """Level operators of a multilevel method; level 0 is the coarsest
    (MLHierarchy.py:6-54)."""

    def __init__(self, numLevels=2, normalize=True):
        print('numLevels={}, type={}'.format(numLevels.__repr__(), type(numLevels)))
        self._numLevels = numLevels
        self._ops = [None] * numLevels
        self._updates = [None] * numLevels
        self._downdates = [None] * numLevels
        self._normalize = normalize

    def numLevels(self):
        return self._numLevels

    def update(self, k):
        return self._updates[k]

    def downdate(self, k):
        return self._downdates[k]

    def matrix(self, k):
        return self._ops[k]


def makeRestrictionOp(I_up, normalize=True):
    return amg_setup.restriction_of(I_up, normalize)


class SmoothedAggregationMLHierarchy(MLHierarchy):
    """Hierarchy built by smoothed aggregation (SmoothedAggregation.py:13-31);
    same operators as the reference, bit for bit, in O(nnz)."""

    def __init__(self, A_fine, numLevels=2, tol=None, normalize=True):
        super().__init__(numLevels=numLevels, normalize=normalize)
        # ``tol`` is accepted and stored exactly like the reference does -- and, exactly like
        # there, it changes nothing: SA_coarsen calls BuildAggregates WITHOUT it
        # (SmoothedAggregation.py:218, so the strength test always uses 0.08 * 0.5^(lvl-1)) and
        # BuildFilteredMatrix takes it as an argument it never reads (:156-183).
        self.tol = tol
        self.normalize = normalize
        tab = Tab()
        for lev in reversed(range(numLevels - 1)):
            print('{}making prolongator from level {} to {}'.format(tab, lev, lev + 1))
        if isinstance(A_fine, DeviceCSR):
            A_fine = A_fine.to_scipy()       # setup runs on the host, as in the reference
        ops, ups, downs = amg_setup.build_hierarchy(sp.csr_matrix(A_fine), numLevels, normalize)
        self._ops, self._updates, self._downdates = ops, ups, downs


def SA_coarsen(A, tol=None, lvl=1):
    """(P, aggregates) of one coarsening step (SmoothedAggregation.py:208-229);
    aggregates as a list of sets."""
    # tol: accepted, without effect -- as in the reference (see SmoothedAggregationMLHierarchy)
    P, agg_of = amg_setup.sa_coarsen(A, lvl=lvl)
    n_agg = P.shape[1]
    order = np.argsort(agg_of, kind='stable')
    bounds = np.searchsorted(agg_of[order], np.arange(n_agg + 1))
    return P, [set(order[bounds[j]:bounds[j + 1]].tolist()) for j in range(n_agg)]


COARSE_PERMC_SPEC = 'MMD_AT_PLUS_A'


class DeviceAMG:
    """The uploaded hierarchy + smoother data + coarse LU = one psb_amg handle."""

    def __init__(self, mlh, smoother, nuPre, nuPost, numIters, tau=1.0e-8):
        require_cuda()
        kind, omega = _smoother_spec(smoother)
        nlev = mlh.numLevels()
        self.mlh = mlh
        self.A = [DeviceCSR(mlh.matrix(k)) for k in range(nlev)]
        self.P = [DeviceCSR(mlh.update(k)) for k in range(nlev - 1)]
        self.R = [DeviceCSR(mlh.downdate(k)) for k in range(nlev - 1)]
        self.dinv = [None] * nlev
        self.gsU = [None] * nlev
        for k in range(1, nlev):
            Ak = sp.csr_matrix(mlh.matrix(k))
            if kind == nat.SMOOTH_JACOBI:
                self.dinv[k] = to_device(np.reciprocal(Ak.diagonal()))
            else:
                self.gsU[k] = DeviceTrsv(sp.triu(Ak).tocsr(), lower=False)
        # coarsest level: factor once on the host (the reference re-factorises inside spsolve on
        # every cycle, VCycleManager.py:34-37).  Same exact solve, but with the minimum-degree
        # ordering on A^T + A instead of spsolve's default COLAMD: for these symmetric-structure
        # Galerkin operators it halves the dependency levels of both factors (713 vs 1 361 at
        # Bratu m = 256) and cuts the fill by 40 %, and the device solve is bound by levels.
        A0 = sp.csc_matrix(mlh.matrix(0))
        lu = spla.splu(A0, permc_spec=COARSE_PERMC_SPEC)
        self.coarse = DeviceSplitLU(lu)

        def harr(objs):
            arr = (C.c_void_p * max(len(objs), 1))()
            for i, o in enumerate(objs):
                arr[i] = None if o is None else (o.handle if hasattr(o, 'handle') else o.data_ptr())
            return arr
        h = C.c_void_p()
        nat.check(nat.lib().psb_amg_create(
            nlev, harr(self.A), harr(self.P), harr(self.R), harr(self.dinv), harr(self.gsU),
            self.coarse.handle, kind, omega, int(nuPre), int(nuPost), int(numIters), float(tau),
            C.byref(h)), 'psb_amg_create')
        self.n = self.A[-1].shape[0]
        self.prec = DevicePrec(h, self.n, keep=(self.A, self.P, self.R, self.dinv, self.gsU, self.coarse))

    @property
    def handle(self):
        return self.prec.handle

    def solve(self, b, maxiter, tau, keep_on_device=False):
        """(x, result, hist) of up to maxiter V-cycles from x0 = b."""
        b_d = to_device(b)
        x_d = torch.empty_like(b_d)
        hist = torch.zeros(max(maxiter, 1), dtype=torch.float64, device=b_d.device)
        res = nat.SolveResult()
        nat.check(nat.lib().psb_amg_solve(self.handle, ptr(b_d), ptr(x_d), int(maxiter), float(tau),
                                          ptr(hist), C.byref(res), current_stream_ptr()), 'psb_amg_solve')
        return (x_d if keep_on_device else to_host(x_d)), res, hist[:res.n_hist].cpu().numpy()


class VCycleManager:
    """Runs single V-cycles on a hierarchy (VCycleManager.py:10-62)."""

    def __init__(self, mlh, nuPre=2, nuPost=2, smoother=GaussSeidelSmoother):
        assert isinstance(mlh, MLHierarchy)
        self._numLevels = mlh.numLevels()
        self._mlh = mlh
        self._nuPre, self._nuPost = nuPre, nuPost
        self._dev = DeviceAMG(mlh, smoother, nuPre, nuPost, 1)

    def device(self):
        return self._dev


# ------------------------------------------------------------------- solvers --
class AMGVCycle(IterativeLinearSolverType):
    def __init__(self, control=CommonSolverArgs(), numLevels=2, nuPre=2, nuPost=2,
                 smoother=GaussSeidelSmoother, name='AMGVCycle'):
        super().__init__(control=control, precond=None)
        self.numLevels, self.nuPre, self.nuPost = numLevels, nuPre, nuPost
        self.smoother = smoother

    def makeSolver(self, name=None):
        return AMGVCycleSolver(name=self.name() if name is None else name,
                               control=self.control(), numLevels=self.numLevels,
                               nuPre=self.nuPre, nuPost=self.nuPost, smoother=self.smoother)


class AMGVCycleSolver(IterativeLinearSolver):
    """V-cycle iteration x <- V(b, x) from x0 = b (VCycleSolver.py:52-95); the
    hierarchy is rebuilt on every solve unless ``freezeMatrix()`` is set."""

    def __init__(self, control=CommonSolverArgs(), numLevels=2, nuPre=2, nuPost=2,
                 smoother=GaussSeidelSmoother, name='AMGVCycle'):
        super().__init__(control=control, name=name, precond=None)
        self.numLevels, self.nuPre, self.nuPost = numLevels, nuPre, nuPost
        self.smoother = smoother
        self._cycleMgr = None
        self.last_history = None

    def device_amg(self, A):
        if self._cycleMgr is None or not self.matrixFrozen():
            mlh = SmoothedAggregationMLHierarchy(A, numLevels=self.numLevels)
            self._cycleMgr = VCycleManager(mlh, nuPre=self.nuPre, nuPost=self.nuPost,
                                           smoother=self.smoother)
        return self._cycleMgr.device()

    def solve(self, A, b):
        n = self._check_system(A, b)
        self._require_euclidean_norm()
        # device right-hand sides stay on the device, like PCG / GMRES (DeviceFDBratu2D hands them in)
        on_device = isinstance(b, torch.Tensor)
        if not on_device:
            b = np.asarray(b)
        zeros = (lambda: torch.zeros_like(b)) if on_device else (lambda: np.zeros_like(b))
        if n == 0 or not bool((b != 0).any()):
            return self.handleConvergence(0, zeros(), 0, 0)
        dev = self.device_amg(A)
        x, res, hist = dev.solve(b, int(self.maxiter()), float(self.tau()), keep_on_device=on_device)
        self.last_history = hist
        for k in range(res.n_hist):
            self.reportIter(k, hist[k], res.norm_b)
        if res.status == nat.TRIVIAL:
            return self.handleConvergence(0, zeros(), 0, 0)
        if res.status == nat.CONVERGED:
            return self.handleConvergence(res.k, x, res.norm_r, res.norm_b)
        return self.handleMaxiter(res.k, x, res.norm_r, res.norm_b)


# ------------------------------------------------------------ preconditioner --
class AMG(PreconditionerType):
    def __init__(self, numIters=5, numLevels=2, nuPre=2, nuPost=2,
                 smoother=GaussSeidelSmoother):
        self.numIters, self.numLevels = numIters, numLevels
        self.nuPre, self.nuPost = nuPre, nuPost
        self.smoother = smoother

    def form(self, A):
        return AMGPreconditioner(A, numIters=self.numIters, numLevels=self.numLevels,
                                 nuPre=self.nuPre, nuPost=self.nuPost, smoother=self.smoother)


class AMGPreconditioner(GenericPreconditioner):
    """numIters V-cycles with a frozen hierarchy (AMGPreconditioner.py:25-51).
    The inner solver's control is CommonSolverArgs(maxiter=numIters,
    failOnMaxiter=False): default tau 1e-8 and the strict '<' early exit."""

    def __init__(self, A, numIters=5, numLevels=2, nuPre=2, nuPost=2,
                 smoother=GaussSeidelSmoother):
        self._A = A
        self._numIters = numIters
        control = CommonSolverArgs(maxiter=numIters, failOnMaxiter=False)
        self._solver = AMGVCycleSolver(control=control, numLevels=numLevels, nuPre=nuPre,
                                       nuPost=nuPost, smoother=smoother, name='AMG prec')
        self._solver.freezeMatrix()
        mlh = SmoothedAggregationMLHierarchy(A, numLevels=numLevels)
        self._dev = DeviceAMG(mlh, smoother, nuPre, nuPost, numIters, tau=control.tau)
        self._solver._cycleMgr = _FrozenManager(self._dev)

    def apply(self, vec):
        result = self._solver.solve(self._A, vec)
        if result.success():
            return result.soln()
        raise RuntimeError('ML preconditioner failed: {}'.format(result))

    def right_device_handle(self):
        return self._dev.handle

    def device_amg(self):
        return self._dev


class _FrozenManager:
    def __init__(self, dev):
        self._dev = dev

    def device(self):
        return self._dev
