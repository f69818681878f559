"""Row-partitioned GMRES, AMG V-cycle preconditioner and Bratu problem: the multi-GPU form of
configs[4] (Newton + inexact GMRES + AMG on FDBratu2D), one process per GPU.

The reference is single-process; the contract is SURVEY.md section 8e -- "AMG with Jacobi
smoothing shards like SpMV (coarse solve replicated)" -- and the loops being sharded are
PySolvers/Linear/GMRESSolver.py:104-125 (Arnoldi: every dot product all-reduced) and
PySolvers/Linear/VCycleManager.py:31-62 (every level operator a row block with its own halo
plan).  Arithmetic per row is unchanged -- Jacobi sweeps, residuals, restriction and
prolongation sum each row in stored order exactly as on one GPU -- so a row-partitioned solve
differs from the single-GPU one only through the summation order of the all-reduced scalars.

Setup stays on the host as in the reference and is REPLICATED: every rank builds the same
hierarchy (bit-identical, Linear/amg_setup.py) from the global matrix, slices its rows, and
factors the coarsest operator for the replicated coarse solve.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

from . import _native as nat
from .core import CommonSolverArgs, SolveStatus
from .device import DeviceCSR, DevicePrec, DeviceSplitLU, current_stream_ptr, ptr, to_device, to_host
from .dist import DistCSR, row_starts
from .Linear.base import IterativeLinearSolver, IterativeLinearSolverType
from .Linear.multigrid import (COARSE_PERMC_SPEC, DampedJacobiSmoother, SmoothedAggregationMLHierarchy,
                               _smoother_spec)
from .Linear.precond import IdentityPreconditionerType, PreconditionerType


def dist_norm(x_loc):
    """2-norm of a row-partitioned vector from its local slice (CUDA tensor): the local sum
    of squares on the device, all-reduced over the ranks."""
    import torch.distributed as dist
    s = torch.dot(x_loc, x_loc).reshape(1)
    dist.all_reduce(s)
    return float(torch.sqrt(s).item())


def _block(M, lo, hi):
    B = sp.csr_matrix(M)[lo:hi, :]
    return B.indptr.astype(np.int32), B.indices.astype(np.int32), B.data.astype(np.float64)


class DistAMGPreconditioner:
    """numIters V-cycles of the smoothed-aggregation hierarchy of the GLOBAL matrix ``A``
    (scipy CSR, the same on every rank), applied to row-partitioned vectors
    (AMGPreconditioner.py:25-51 / VCycleManager.py:31-62 sharded).  Jacobi smoothing only."""

    def __init__(self, comm, A, numIters=5, numLevels=2, nuPre=2, nuPost=2,
                 smoother=DampedJacobiSmoother, tau=1.0e-8):
        import contextlib
        import io
        kind, omega = _smoother_spec(smoother)
        if kind != nat.SMOOTH_JACOBI:
            raise NotImplementedError('the row-partitioned V-cycle supports Jacobi smoothing only: Gauss-Seidel '
                                      'is one global dependency chain (SURVEY.md section 8e: replicas only)')
        if numLevels < 2:
            raise ValueError('numLevels must be >= 2')
        self.comm = comm
        world, rank = comm.world, comm.rank
        with contextlib.redirect_stdout(io.StringIO()):
            mlh = SmoothedAggregationMLHierarchy(sp.csr_matrix(A), numLevels=numLevels)
        self.mlh = mlh
        nlev = numLevels
        sizes = [mlh.matrix(k).shape[0] for k in range(nlev)]
        starts = [row_starts(n, world) for n in sizes]
        self.starts = starts
        self.A = [None] * nlev
        self.R = [None] * nlev
        self.P = [None] * nlev
        self.dinv = [None] * nlev
        for k in range(1, nlev):
            lo, hi = int(starts[k][rank]), int(starts[k][rank + 1])
            Ak = sp.csr_matrix(mlh.matrix(k))
            self.A[k] = DistCSR(comm, *_block(Ak, lo, hi), lo, hi, sizes[k], p2p=False)
            self.dinv[k] = to_device(np.reciprocal(Ak.diagonal()[lo:hi]))
            # restriction to level k-1: rows of level k-1, input = level-k vector
            clo, chi = int(starts[k - 1][rank]), int(starts[k - 1][rank + 1])
            self.R[k - 1] = DistCSR(comm, *_block(mlh.downdate(k - 1), clo, chi), clo, chi, sizes[k - 1],
                                    n_cols=sizes[k], p2p=False)
            if k - 1 >= 1:     # prolongator from level k-1: rows of level k, input = level-(k-1) vector
                self.P[k - 1] = DistCSR(comm, *_block(mlh.update(k - 1), lo, hi), lo, hi, sizes[k],
                                        n_cols=sizes[k - 1], p2p=False)
        lo1, hi1 = int(starts[1][rank]), int(starts[1][rank + 1])
        self.P0 = DeviceCSR(sp.csr_matrix(mlh.update(0))[lo1:hi1, :])      # global coarse columns
        lu = spla.splu(sp.csc_matrix(mlh.matrix(0)), permc_spec=COARSE_PERMC_SPEC)
        self.coarse = DeviceSplitLU(lu)

        def harr(objs):
            arr = (C.c_void_p * max(len(objs), 1))()
            for i, o in enumerate(objs):
                arr[i] = None if o is None else (o.handle if hasattr(o, 'handle') else o.data_ptr())
            return arr
        st0 = (C.c_int64 * (world + 1))(*[int(v) for v in starts[0]])
        h = C.c_void_p()
        nat.check(nat.lib().psb_dist_amg_create(
            comm.handle, nlev, harr(self.A), harr(self.P), harr(self.R), self.P0.handle, harr(self.dinv),
            self.coarse.handle, st0, float(omega), int(nuPre), int(nuPost), int(numIters), float(tau),
            C.byref(h)), 'psb_dist_amg_create')
        self.n_loc = int(starts[nlev - 1][rank + 1] - starts[nlev - 1][rank])
        self.prec = DevicePrec(h, self.n_loc, keep=(self.A, self.P, self.R, self.P0, self.dinv, self.coarse))

    @property
    def handle(self):
        return self.prec.handle

    def right_device_handle(self):
        return self.prec.handle

    def apply(self, r_loc, out=None):
        """z_loc = M^-1 r (slices in, slices out); collective."""
        return self.prec.apply(r_loc, out)

    def solve(self, b_loc, maxiter, tau):
        """(x_loc, result, hist) of up to maxiter V-cycles from x0 = b (VCycleSolver.py:52-95)."""
        b_d = to_device(b_loc)
        x_d = torch.empty_like(b_d)
        hist = torch.zeros(max(maxiter, 1), dtype=torch.float64, device=b_d.device)
        res = nat.SolveResult()
        nat.check(nat.lib().psb_amg_solve(self.handle, ptr(b_d), ptr(x_d), int(maxiter), float(tau),
                                          ptr(hist), C.byref(res), current_stream_ptr()), 'psb_amg_solve')
        return x_d, res, hist[:res.n_hist].cpu().numpy()

    def close(self):
        for group in (self.A, self.R, self.P):
            for d in group:
                if d is not None:
                    d.close()


class DistAMG(PreconditionerType):
    """Factory with the reference's AMG(...) arguments (AMGPreconditioner.py:8-21); ``form`` takes
    the GLOBAL matrix (host scipy CSR).  Default smoother here: damped Jacobi (north_star item 5)."""

    def __init__(self, comm, numIters=5, numLevels=2, nuPre=2, nuPost=2, smoother=DampedJacobiSmoother):
        self.comm = comm
        self.numIters, self.numLevels = numIters, numLevels
        self.nuPre, self.nuPost, self.smoother = nuPre, nuPost, smoother

    def form(self, A_global):
        return DistAMGPreconditioner(self.comm, A_global, numIters=self.numIters, numLevels=self.numLevels,
                                     nuPre=self.nuPre, nuPost=self.nuPost, smoother=self.smoother)


class DistributedGMRES(IterativeLinearSolverType):
    """Factory of DistributedGMRESSolver; arguments as GMRES(...) (GMRESSolver.py:27-40)."""

    def __init__(self, control=CommonSolverArgs(), precond=IdentityPreconditionerType(), name='GMRES',
                 orth='cgs2'):
        super().__init__(control=control, precond=precond, name=name)
        self.orth = orth

    def makeSolver(self, name=None):
        return DistributedGMRESSolver(self.control(), precond=self.precond(),
                                      name=self.name() if name is None else name, orth=self.orth)


class DistributedGMRESSolver(IterativeLinearSolver):
    """Right-preconditioned, un-restarted GMRES on a row-partitioned system (GMRESSolver.py:75-174
    with every reduction all-reduced).  ``solve(D, b_local)``: ``D`` a dist.DistCSR -- or an object
    with ``.dist`` (the DistCSR) and ``.global_matrix()`` (host scipy CSR, only needed to form an AMG
    preconditioner) such as DistFDBratu2D's Jacobian; returns the local slice of x as a CUDA
    tensor.  The preconditioner is formed once and kept while ``freezePrec()`` is in force
    (what "preconditioner reuse" in configs[4] intends); otherwise on every solve like the
    reference."""

    def __init__(self, control=CommonSolverArgs(), precond=IdentityPreconditionerType(), name='GMRES',
                 orth='cgs2'):
        super().__init__(control=control, precond=precond, name=name)
        self.precond = None
        self.orth = orth
        self.last_history = None

    def solve(self, A, b):
        self._require_euclidean_norm()
        D = getattr(A, 'dist', A)
        b_d = to_device(b)
        n = D.n_loc
        assert b_d.numel() == n
        if dist_norm(b_d) == 0.0:
            return self.handleConvergence(0, torch.zeros_like(b_d), 0, 0)
        ptype = self.precondType()
        prec_h = None
        if not isinstance(ptype, IdentityPreconditionerType):
            if self.precond is None or not self.precFrozen():
                if self.precond is not None and hasattr(self.precond, 'close'):
                    self.precond.close()
                if not hasattr(A, 'global_matrix'):
                    raise TypeError('forming a preconditioner needs the global matrix: pass a '
                                    'dist_krylov.DistOperator(dist_csr, global_matrix) instead of the bare DistCSR')
                self.precond = ptype.form(A.global_matrix())
            prec_h = self.precond.right_device_handle()
        lib = nat.lib()
        maxiter = int(self.maxiter())
        wbytes = int(lib.psb_dist_gmres_workspace_bytes(n, D.n_halo, maxiter))
        work = torch.empty(wbytes, dtype=torch.uint8, device=b_d.device)
        x_d = torch.empty(max(n, 1), dtype=torch.float64, device=b_d.device)
        hist_d = torch.empty(max(maxiter, 1), dtype=torch.float64, device=b_d.device)
        res = nat.SolveResult()
        orth = {'cgs2': nat.ORTH_CGS2, 'mgs': nat.ORTH_MGS}[self.orth]
        nat.check(lib.psb_dist_gmres_solve(
            D.handle, prec_h, ptr(b_d), ptr(x_d), ptr(work), wbytes, maxiter, float(self.tau()),
            1 if self.failOnMaxiter() else 0, orth, ptr(hist_d), C.byref(res), current_stream_ptr()),
            'psb_dist_gmres_solve')
        hist = hist_d[:res.n_hist].cpu().numpy()
        self.last_history = hist
        for k in range(res.n_hist):
            self.reportIter(k, hist[k], res.norm_b)
        x = x_d[:n]
        if res.status == nat.TRIVIAL:
            return self.handleConvergence(0, torch.zeros_like(b_d), 0, 0)
        if res.status == nat.CONVERGED:
            return self.handleConvergence(res.k, x, res.norm_r, res.norm_b)
        if res.status == nat.GMRES_FALSE_CONV:
            return SolveStatus(
                success=False, iters=res.k + 1, soln=x, resid=res.norm_r,
                msg='GMRES failure: true residual %12.5g did not meet tolerance '
                    'tau=%12.5g. Recursive residual was %12.5g.' % (res.norm_r, self.tau(), res.norm_r_rec))
        return self.handleMaxiter(res.k, x, res.norm_r_rec, res.norm_b)


class DistOperator:
    """A row block (dist.DistCSR) together with the global matrix it is a block of (host scipy CSR,
    or a callable returning it): what DistributedGMRESSolver needs when it has to FORM a
    preconditioner whose setup is global (the smoothed-aggregation hierarchy)."""

    def __init__(self, dist_csr, global_matrix):
        self.dist = dist_csr
        self._global = global_matrix
        self.shape = (dist_csr.n, dist_csr.n_cols)

    def global_matrix(self):
        return self._global() if callable(self._global) else self._global


class _DistJacobian:
    """What DistFDBratu2D.evalJ returns: the row block (DistCSR, values updated in place) plus a
    way to get the global matrix on the host when a preconditioner has to be formed."""

    def __init__(self, problem):
        self._p = problem
        self.dist = problem._DJ
        self.shape = (problem.n, problem.n)

    def global_matrix(self):
        return self._p.global_jacobian()


class DistFDBratu2D:
    """FDBratu2D (examples/FDBratu2D.py:10-29) with u, F and the Jacobian row-partitioned over the
    ranks of ``comm``: evalF / evalJ take and return local CUDA slices / a row-block Jacobian
    whose diagonal values are rewritten in place.  Drop it into NewtonSolver with
    ``CommonSolverArgs(norm=dist_norm)`` and a DistributedGMRES linear solver."""

    def __init__(self, comm, m=4, alpha=0.5):
        from .problems import fd_laplacian_2d
        self.comm, self.m, self.alpha = comm, m, alpha
        self.n = m * m
        starts = row_starts(self.n, comm.world)
        self.lo, self.hi = int(starts[comm.rank]), int(starts[comm.rank + 1])
        self.starts = starts
        blk = -fd_laplacian_2d(-1.0, 1.0, m, row_lo=self.lo, row_hi=self.hi)     # my rows, global columns
        self._blk = blk
        n_loc = self.hi - self.lo
        rows = np.repeat(np.arange(self.lo, self.hi, dtype=np.int64), np.diff(blk.indptr))
        dpos = np.flatnonzero(rows == blk.indices)
        assert dpos.size == n_loc
        ip, ix = blk.indptr.astype(np.int32), blk.indices.astype(np.int32)
        self._DA = DistCSR(comm, ip, ix, blk.data.copy(), self.lo, self.hi, self.n, p2p=False)
        self._DJ = DistCSR(comm, ip.copy(), ix.copy(), blk.data.copy(), self.lo, self.hi, self.n, p2p=False)
        assert self._DJ.A is not self._DA.A
        self._diag_pos = torch.from_numpy(dpos).cuda()
        self._a_diag = to_device(blk.data[dpos])
        self._u_last = None

    def initialU(self):
        return torch.ones(self.hi - self.lo, dtype=torch.float64, device='cuda')

    def evalF(self, u):
        Au = self._DA.matvec(u)
        F = torch.empty_like(u)
        nat.check(nat.lib().psb_bratu_residual(u.numel(), ptr(Au), ptr(u), float(self.alpha), ptr(F),
                                               current_stream_ptr()), 'psb_bratu_residual')
        return F

    def evalJ(self, u):
        nat.check(nat.lib().psb_bratu_jacobian(u.numel(), ptr(self._diag_pos), ptr(self._a_diag), ptr(u),
                                               float(self.alpha), ptr(self._DJ.A.data), current_stream_ptr()),
                  'psb_bratu_jacobian')
        self._u_last = u
        return _DistJacobian(self)

    def global_jacobian(self):
        """J(u_last) as a host scipy CSR on every rank: the Laplacian plus the gathered diagonal
        values exactly as the device computed them.  Only used to FORM the AMG preconditioner,
        i.e. once per Newton solve with freezePrec."""
        import torch.distributed as dist
        from .problems import fd_laplacian_2d
        mine = self._DJ.A.data[self._diag_pos].cpu().numpy()
        parts = [None] * self.comm.world
        dist.all_gather_object(parts, mine)
        J = -fd_laplacian_2d(-1.0, 1.0, self.m)
        J.setdiag(np.concatenate(parts))
        return J

    def close(self):
        self._DA.close()
        self._DJ.close()
