"""Oracle restatement of the reference AMG solve phase (TEST INFRASTRUCTURE ONLY).

* smoothers      -- PySolvers/Linear/ClassicSmoothers.py:5-36
* V-cycle        -- PySolvers/Linear/VCycleManager.py:26-62
* V-cycle solver -- PySolvers/Linear/VCycleSolver.py:52-95
* preconditioner -- PySolvers/Linear/AMGPreconditioner.py:46-51

The hierarchy is passed in as plain lists ``ops[k]`` (level matrices, level 0
coarsest), ``ups[k]`` (level k -> k+1) and ``downs[k]`` (level k+1 -> k), i.e.
whatever ``MLHierarchy.matrix/update/downdate`` return (MLHierarchy.py:35-48).
"""
import numpy as np
import numpy.linalg as npla
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class Jacobi:
    """x <- x + omega * D^-1 (f - A x); the reference has omega = 1
    (ClassicSmoothers.py:5-16).  omega != 1 is the harness-supplied damped
    variant with the same plug-in protocol."""

    def __init__(self, A, omega=1.0):
        self.A = A
        self.omega = omega
        self.dinv = np.reciprocal(A.diagonal())

    def apply(self, f, x, nu):
        for _ in range(nu):
            r = f - self.A * x
            if self.omega == 1.0:
                x = x + np.multiply(self.dinv, r)
            else:
                x = x + self.omega * np.multiply(self.dinv, r)
        return x


class GaussSeidel:
    """x <- x + triu(A)^-1 (f - A x) (ClassicSmoothers.py:20-36; the reference
    calls the general sparse direct solver on the triangular matrix)."""

    def __init__(self, A):
        self.A = A
        self.U = sp.triu(A).tocsr()

    def apply(self, f, x, nu):
        for _ in range(nu):
            r = f - self.A * x
            x = x + spla.spsolve(self.U, r)
        return x


def run_level(ops, ups, downs, smoothers, f, x, lev, nu_pre, nu_post):
    # VCycleManager.py:31-62
    if lev == 0:
        return spla.spsolve(ops[0], f)
    x = smoothers[lev].apply(f, x, nu_pre)
    r = f - ops[lev] * x
    r2 = downs[lev - 1] * r
    e2 = run_level(ops, ups, downs, smoothers, r2, np.zeros_like(r2),
                   lev - 1, nu_pre, nu_post)
    x = x + ups[lev - 1] * e2
    return smoothers[lev].apply(f, x, nu_post)


def vcycle_solve(ops, ups, downs, b, maxiter=100, tau=1.0e-8,
                 fail_on_maxiter=True, nu_pre=2, nu_post=2,
                 smoother=GaussSeidel):
    """AMGVCycleSolver.solve with a prebuilt hierarchy (VCycleSolver.py:52-95).
    Note x0 = b (:69) and the strict '<' test (:90)."""
    nlev = len(ops)
    A = ops[nlev - 1]
    hist = []
    norm_b = npla.norm(b)
    if norm_b == 0.0:
        return dict(success=True, iters=1, soln=np.zeros_like(b), resid=0,
                    hist=np.asarray(hist))
    smoothers = [smoother(ops[k]) for k in range(nlev)]
    x = np.copy(b)
    k = -1
    norm_r = None
    for k in range(maxiter):
        x = run_level(ops, ups, downs, smoothers, b, x, nlev - 1,
                      nu_pre, nu_post)
        r = b - A * x
        norm_r = npla.norm(r)
        hist.append(norm_r)
        if norm_r < tau * norm_b:
            return dict(success=True, iters=k + 1, soln=x, resid=norm_r,
                        hist=np.asarray(hist))
    # handleMaxiter (IterativeSolver.py:117-129)
    return dict(success=not fail_on_maxiter, iters=k, soln=x, resid=norm_r,
                hist=np.asarray(hist))


def amg_apply(ops, ups, downs, v, num_iters=5, nu_pre=2, nu_post=2,
              smoother=GaussSeidel):
    """AMGPreconditioner.apply: num_iters V-cycles, failOnMaxiter=False,
    default tau=1e-8 of CommonSolverArgs (AMGPreconditioner.py:39-51)."""
    res = vcycle_solve(ops, ups, downs, v, maxiter=num_iters, tau=1.0e-8,
                       fail_on_maxiter=False, nu_pre=nu_pre, nu_post=nu_post,
                       smoother=smoother)
    return res['soln']
