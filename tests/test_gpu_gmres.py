"""GMRES parity on the B200 (config 2: DH matrices, GMRES(maxiter=30) + RightILUT)
against the histories the reference itself produced, plus MGS/CGS2 modes,
the left-preconditioner no-op and the exit conventions."""
import contextlib
import io

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _run(solver, A, b):
    hist = []
    solver.reportIter = lambda k, nr, nb: hist.append(nr)
    with contextlib.redirect_stdout(io.StringIO()):
        st = solver.solve(A, b)
    return st, np.asarray(hist)


def _dh(lev):
    from pysolvers_b200.problems import load_dh_matrix
    A = load_dh_matrix(lev)
    return A, A @ np.random.default_rng(2024).random(A.shape[0])


@pytest.mark.parametrize('lev', list(range(16)))
@pytest.mark.parametrize('orth', ['cgs2', 'mgs'])
def test_gmres_ilut_history_vs_reference_golden(cuda, golden, lev, orth):
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import GMRES, RightILUT
    A, b = _dh(lev)
    s = GMRES(CommonSolverArgs(maxiter=30, tau=1e-8), precond=RightILUT(), orth=orth).makeSolver()
    st, hist = _run(s, A, b)
    key = 'gmres_ilut/dh%d' % lev
    g = golden[key + '/hist']
    assert st.success()
    assert abs(st.iters() - int(golden[key + '/iters'])) <= 1
    k = min(len(hist), len(g))
    # residuals below 1e-13 ||b|| are rounding noise of the recursion itself
    floor = 1e-13 * np.linalg.norm(b)
    sel = g[:k] > floor
    assert rel_err(hist[:k][sel], g[:k][sel]) < 1e-10
    if lev <= 12:
        gx = golden[key + '/x']
        assert np.linalg.norm(st.soln() - gx) <= 1e-8 * np.linalg.norm(gx)
    true_r = np.linalg.norm(b - A @ st.soln())
    assert abs(st.resid() - true_r) <= 1e-6 * true_r + 1e-14 * np.linalg.norm(b)


@pytest.mark.parametrize('lev', [5, 8])
def test_gmres_unpreconditioned_mgs_and_left_noop(cuda, golden, lev):
    from oracle import krylov
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import GMRES, LeftILUT
    A, b = _dh(lev)
    g = golden['gmres/dh%d/hist' % lev]
    st, hist = _run(GMRES(CommonSolverArgs(maxiter=100, tau=1e-8), orth='mgs').makeSolver(), A, b)
    assert st.success() and abs(st.iters() - int(golden['gmres/dh%d/iters' % lev])) <= 1
    k = min(len(hist), len(g))
    # Un-preconditioned MGS Arnoldi on these FE matrices amplifies 1-ulp changes of the
    # dot products.  The noise floor of THIS input is measured by running
    # the oracle against itself with only the summation order of its dots changed; the GPU
    # (same MGS order as the reference, own reduction tree) must stay within 10x of it.
    alt = krylov.gmres(A, b, maxiter=100, tau=1e-8, dot=krylov.pairwise_dot)
    ka = min(k, len(alt['hist']))
    floor = np.maximum.accumulate(np.abs(alt['hist'][:ka] - g[:ka]) / g[:ka])
    err = np.abs(hist[:ka] - g[:ka]) / g[:ka]
    assert np.all(err <= np.maximum(1e-10, 10.0 * floor)), float(np.max(err))
    assert rel_err(hist[:5], g[:5]) < 1e-10
    st2, hist2 = _run(GMRES(CommonSolverArgs(maxiter=100, tau=1e-8), orth='cgs2').makeSolver(), A, b)
    assert st2.success() and abs(st2.iters() - st.iters()) <= 1
    gx = golden['gmres/dh%d/x' % lev]
    for s in (st, st2):
        assert np.linalg.norm(s.soln() - gx) <= 1e-8 * np.linalg.norm(gx)
    # LeftILUT is the identity from the right: identical history to no preconditioner
    st3, hist3 = _run(GMRES(CommonSolverArgs(maxiter=100, tau=1e-8), precond=LeftILUT(),
                            orth='mgs').makeSolver(), A, b)
    assert np.array_equal(hist3, hist)


def test_gmres_exits(cuda):
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import GMRES
    A, b = _dh(9)
    st, h = _run(GMRES(CommonSolverArgs(maxiter=5, tau=1e-12)).makeSolver(), A, b)
    assert (st.success(), st.iters(), st.msg()) == (False, 4, 'failure to converge')
    assert len(h) == 5 and st.soln() is not None
    st, h = _run(GMRES(CommonSolverArgs(maxiter=5)).makeSolver(), A, np.zeros(A.shape[0]))
    assert (st.success(), st.iters(), st.resid()) == (True, 1, 0)
    # lucky breakdown: b is an eigenvector of a diagonal matrix -> one iteration
    import scipy.sparse as sp
    D = sp.diags(np.arange(1.0, 51.0)).tocsr()
    e = np.zeros(50)
    e[7] = 3.0
    st, h = _run(GMRES(CommonSolverArgs(maxiter=10)).makeSolver(), D, e)
    assert st.success() and st.iters() == 1
    assert np.allclose(st.soln(), e / 8.0, rtol=1e-14)


def test_gmres_preconditioner_always_rebuilt(cuda):
    """freezePrec has no effect for GMRES in the reference (GMRESSolver.py:71-72)."""
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import GMRES, RightILUT

    class Counting(RightILUT):
        built = 0

        def form(self, A):
            Counting.built += 1
            return super().form(A)
    A, b = _dh(6)
    s = GMRES(CommonSolverArgs(maxiter=30), precond=Counting()).makeSolver()
    s.freezePrec()
    _run(s, A, b)
    _run(s, A, b)
    assert Counting.built == 2


def test_gmres_honor_freeze_is_opt_in(cuda):
    """The reference's GMRES rebuilds its preconditioner on every solve (freezePrec has no effect,
    GMRESSolver.py:71-72) and so does ours by default; ``honorFreeze=True`` keeps it while frozen,
    like PCGSolver (PCGSolver.py:92-94)."""
    import contextlib
    import io
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import GMRES, RightILUT
    from pysolvers_b200.problems import load_dh_matrix
    A = load_dh_matrix(8)
    b = A @ np.ones(A.shape[0])
    built = []
    ptype = RightILUT()
    form0 = ptype.form
    ptype.form = lambda M: built.append(1) or form0(M)
    for honor, want in ((False, 3), (True, 1)):
        del built[:]
        s = GMRES(CommonSolverArgs(maxiter=30, tau=1e-8), precond=ptype, honorFreeze=honor).makeSolver()
        s.freezePrec()
        outs = []
        for _ in range(3):
            with contextlib.redirect_stdout(io.StringIO()):
                outs.append(s.solve(A, b))
        assert len(built) == want
        assert all(o.success() and o.iters() == outs[0].iters() for o in outs)
        assert np.array_equal(outs[0].soln(), outs[2].soln())
