"""Oracle restatement of the reference Krylov loops (TEST INFRASTRUCTURE ONLY).

Follows, statement by statement and with the same numpy/scipy calls so that
rounding is identical:

* ``pcg``   -- PySolvers/Linear/PCGSolver.py:79-142
* ``gmres`` -- PySolvers/Linear/GMRESSolver.py:60-180 with
  PySolvers/Linear/Givens.py:7-34 for the rotations
* result conventions -- PySolvers/IterativeSolver.py:101-129

Results are plain dicts: ``success, iters, soln, resid, msg, hist`` where
``hist[k]`` is the residual norm the reference hands to ``reportIter`` at
iteration k.
"""
import numpy as np
import numpy.linalg as npla
import scipy.sparse as sp


def _matvec(A, x):
    # PySolvers/Linear/IterativeLinearSolver.py:94-106 (mvmult)
    return A * x if sp.issparse(A) else np.dot(A, x)


def _result(success, iters, soln, resid, hist, msg=None):
    return dict(success=success, iters=iters, soln=soln, resid=resid,
                msg=msg, hist=np.asarray(hist, dtype=np.float64))


def pairwise_dot(a, b):
    """Same dot product, different (pairwise) summation order.  Running the
    oracle with this instead of ``np.dot`` measures how far a residual history
    moves when ONLY the rounding of the reductions changes -- the noise floor
    any re-implementation with its own reduction tree lives on."""
    return np.sum(a * b)


def pcg(A, b, prec=None, maxiter=100, tau=1.0e-8, fail_on_maxiter=True,
        dot=np.dot):
    """Preconditioned CG.  ``prec`` is a callable r -> M^{-1} r or None.
    ``dot`` is np.dot as in the reference; tests pass ``pairwise_dot`` to
    measure rounding sensitivity.

    With ``prec=None`` the identity returns its argument, so u aliases r exactly
    as in the reference (Preconditioner.py:58-68).
    """
    apply_prec = prec if prec is not None else (lambda v: v)
    n, nc = A.shape
    assert n == nc and n == len(b)
    hist = []

    norm = npla.norm if dot is np.dot else (lambda v: np.sqrt(dot(v, v)))
    norm_b = norm(b)                                        # PCGSolver.py:86
    if norm_b == 0.0:                                       # :87-88
        return _result(True, 1, np.zeros_like(b), 0, hist)

    r = np.copy(b)                                          # :97
    p = apply_prec(r)                                       # :98
    u = np.copy(p)                                          # :99
    x = np.zeros_like(b)                                    # :100
    u_dot_r = dot(u, r)                                     # :102
    if u_dot_r == 0.0:                                      # :104-105
        return _result(False, 0, None, None, hist, 'breakdown dot(u,r)==0')

    k = -1
    norm_r = None
    for k in range(maxiter):                                # :109
        Ap = _matvec(A, p)                                  # :111
        pAp = dot(p, Ap)                                    # :113
        if pAp == 0.0:                                      # :114-115
            return _result(False, k, None, None, hist,
                           'breakdown dot(p, Ap)==0')
        alpha = u_dot_r / pAp                               # :118
        x = x + alpha * p                                   # :121
        r = r - alpha * Ap                                  # :122
        u = apply_prec(r)                                   # :123
        norm_r = norm(r)                                    # :125
        hist.append(norm_r)                                 # :126 reportIter
        if (norm_r <= tau * norm_b) or ((not fail_on_maxiter)
                                        and k == maxiter - 1):   # :129-131
            return _result(True, k + 1, x, norm_r, hist)
        new_u_dot_r = dot(u, r)                             # :134
        beta = new_u_dot_r / u_dot_r                        # :135
        u_dot_r = new_u_dot_r                               # :136
        p = u + beta * p                                    # :138

    # :142 handleMaxiter(k, ...) with the loop variable's last value
    if fail_on_maxiter:
        return _result(False, k, x, norm_r, hist, 'failure to converge')
    return _result(True, k, x, norm_r, hist)


def givens_coefficients(v, i):
    # Givens.py:7-12 -- plain sqrt of the sum of squares, not hypot
    hyp = np.sqrt(v[i + 1] * v[i + 1] + v[i] * v[i])
    return v[i] / hyp, v[i + 1] / hyp


def givens_apply(v, c, s, i):
    # Givens.py:28-34
    a, bb = v[i], v[i + 1]
    v[i] = c * a + s * bb
    v[i + 1] = -s * a + c * bb


def gmres(A, b, prec=None, maxiter=100, tau=1.0e-8, fail_on_maxiter=True,
          dot=np.dot):
    """Right-preconditioned, un-restarted MGS GMRES (GMRESSolver.py:55-180).

    The reference's exit at maxiter raises NameError (``norm_k`` undefined,
    GMRESSolver.py:180); the oracle returns what ``handleMaxiter`` would have
    produced from the last implicit residual instead and marks it in ``msg``.
    """
    apply_prec = prec if prec is not None else (lambda v: v)
    n, nc = A.shape
    assert n == nc and n == len(b)
    hist = []

    norm = npla.norm if dot is np.dot else (lambda v: np.sqrt(dot(v, v)))
    norm_b = norm(b)                                        # :66
    if norm_b == 0.0:
        return _result(True, 1, np.zeros_like(b), 0, hist)

    Q = np.zeros([n, maxiter + 1])                          # :77
    H = np.zeros([maxiter + 1, maxiter])                    # :80
    CS = np.zeros([maxiter, 2])                             # :83
    beta = norm(b)                                          # :90
    Q[:, 0] = b / beta                                      # :91
    g = np.zeros(maxiter + 1)
    g[0] = 1.0
    g = beta * g                                            # :95-97
    breakdown = False
    norm_r_k = None
    k = -1
    for k in range(maxiter):                                # :104
        u = _matvec(A, apply_prec(Q[:, k]))                 # :107
        for j in range(k + 1):                              # :110-112 (MGS)
            H[j, k] = dot(Q[:, j], u)
            u -= H[j, k] * Q[:, j]
        H[k + 1, k] = norm(u)                               # :115
        col_norm = npla.norm(H[0:k + 1, k])                 # :121
        if abs(H[k + 1, k]) <= 1.0e-16 * col_norm:          # :122-123
            breakdown = True
        else:
            Q[:, k + 1] = u / H[k + 1, k]                   # :125
        for j in range(k):                                  # :133-135
            givens_apply(H[:, k], CS[j, 0], CS[j, 1], j)
        CS[k, :] = givens_coefficients(H[:, k], k)          # :140
        givens_apply(H[:, k], CS[k, 0], CS[k, 1], k)        # :145
        givens_apply(g, CS[k, 0], CS[k, 1], k)              # :148
        norm_r_k = np.abs(g[k + 1])                         # :152
        hist.append(norm_r_k)                               # :155
        if breakdown or (norm_r_k <= tau * norm_b):         # :158
            y = npla.solve(H[0:k + 1, 0:k + 1], g[0:k + 1])     # :159
            x = apply_prec(np.dot(Q[:, 0:k + 1], y))        # :160
            resid = b - _matvec(A, x)                       # :163
            norm_true = npla.norm(resid)                    # :164
            if norm_true <= tau * norm_b:                   # :165-166
                return _result(True, k + 1, x, norm_true, hist)
            return _result(False, k + 1, x, norm_true, hist,
                           'GMRES failure: true residual did not meet tolerance')
    msg = 'failure to converge (reference raises NameError here)'
    if fail_on_maxiter:
        return _result(False, k, None, norm_r_k, hist, msg)
    return _result(True, k, None, norm_r_k, hist, msg)
