"""Generate the golden fixtures by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

It imports the unmodified reference from /root/reference with two stand-in
modules for its un-vendored dependencies (tests/golden/_shims/PyTab.py,
PyTimer.py), runs the reference classes on seeded inputs and stores residual
histories, iteration counts, solutions and AMG hierarchy pieces in
tests/golden/reference_golden.npz, plus the DH test matrices as COO triplets in
tests/golden/matrices/.  Harness workarounds (no edits to the reference):
``GMRESSolver.precond = None`` before solve (GMRESSolver.py:71 reads an
attribute the constructor never sets); histories captured by overriding
``reportIter``; stdout silenced; every RNG seeded.
"""
import os
os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')   # ddot order depends on threads
import contextlib
import io
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path[:0] = [os.path.join(HERE, '_shims'), REF, os.path.join(REF, 'examples')]

import PySolvers  # noqa: E402
from PySolvers import CommonSolverArgs  # noqa: E402
from PySolvers.Linear import (PCG, GMRES, RightIC, RightILUT, LeftILUT, AMG,  # noqa: E402
                              AMGVCycle)
from PySolvers.Linear.ClassicSmoothers import JacobiSmoother, GaussSeidelSmoother  # noqa: E402
from PySolvers.Linear.SmoothedAggregation import (SmoothedAggregationMLHierarchy,  # noqa: E402
                                                  BuildAggregates)
from PySolvers.Nonlinear import NewtonSolver  # noqa: E402
from FDLaplacian2D import FDLaplacian2D  # noqa: E402
from FDBratu2D import FDBratu2D  # noqa: E402
from scipy.io import mmread  # noqa: E402

G = {}


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def run(solver, A, b):
    hist = []
    solver.reportIter = lambda k, nr, nb: hist.append(nr)
    st = quiet(solver.solve, A, b)
    return st, np.asarray(hist, dtype=np.float64)


def put(prefix, st, hist, keep_x=True, x_stride=1):
    G[prefix + '/hist'] = hist
    G[prefix + '/iters'] = np.int64(-1 if st.iters() is None else st.iters())
    G[prefix + '/success'] = np.bool_(st.success())
    if st.resid() is not None:
        G[prefix + '/resid'] = np.float64(st.resid())
    if keep_x and st.soln() is not None:
        G[prefix + '/x'] = np.asarray(st.soln())[::x_stride]
        G[prefix + '/xnorm'] = np.float64(np.linalg.norm(st.soln()))


def csr_put(prefix, M):
    M = sp.csr_matrix(M)
    G[prefix + '/indptr'] = M.indptr.astype(np.int32)
    G[prefix + '/indices'] = M.indices.astype(np.int32)
    G[prefix + '/data'] = M.data.astype(np.float64)
    G[prefix + '/shape'] = np.asarray(M.shape, dtype=np.int64)


class DampedJacobi:
    """Harness-supplied smoother with the reference's plug-in protocol
    (cls(A), .apply(f, x, nu), ClassicSmoothers.py:6,10); omega = 2/3."""
    omega = 2.0 / 3.0

    def __init__(self, A):
        self.A = A
        self.DInv = np.reciprocal(A.diagonal())

    def apply(self, f, x, nu):
        for _ in range(nu):
            r = f - self.A * x
            x = x + self.omega * np.multiply(self.DInv, r)
        return x


def main():
    # ---------------- config 1: un-preconditioned PCG on the 2-D Laplacian ---
    for m in (16, 64, 256):
        A = -quiet(FDLaplacian2D, 0.0, 1.0, m)
        n = A.shape[0]
        for tag, b in (('ones', np.ones(n)),
                       ('rand', A @ np.random.default_rng(12345).random(n))):
            s = PCG(CommonSolverArgs(maxiter=5000, tau=1e-8)).makeSolver()
            st, h = run(s, A, b)
            put('pcg/lap2d_m%d_%s' % (m, tag), st, h, x_stride=1 if m <= 64 else 97)
            print('pcg m=%d %s: iters=%d' % (m, tag, st.iters()))
    # maxiter exits (fail / not fail) and trivial rhs
    A = -quiet(FDLaplacian2D, 0.0, 1.0, 16)
    b = np.ones(A.shape[0])
    st, h = run(PCG(CommonSolverArgs(maxiter=7, tau=1e-12)).makeSolver(), A, b)
    put('pcg/maxiter_fail', st, h)
    st, h = run(PCG(CommonSolverArgs(maxiter=7, tau=1e-12, failOnMaxiter=False)).makeSolver(), A, b)
    put('pcg/maxiter_ok', st, h)
    st, h = run(PCG(CommonSolverArgs(maxiter=7)).makeSolver(), A, np.zeros(A.shape[0]))
    put('pcg/zero_rhs', st, h)

    # ---------------- IC-preconditioned PCG -----------------------------------
    for m in (16, 32, 64):
        A = -quiet(FDLaplacian2D, 0.0, 1.0, m)
        b = np.ones(A.shape[0])
        s = PCG(CommonSolverArgs(maxiter=500, tau=1e-8), precond=RightIC()).makeSolver()
        st, h = run(s, A, b)
        put('icpcg/lap2d_m%d' % m, st, h)
        print('ic-pcg m=%d: iters=%d' % (m, st.iters()))
        if m == 32:
            v = np.random.default_rng(7).random(A.shape[0])
            G['ic/apply_m32_in'] = v
            G['ic/apply_m32_out'] = s.precond.applyRight(v)
            csr_put('ic/L_m32', s.precond._L)
            csr_put('ic/Lt_m32', s.precond._Lt)

    # ---------------- config 2: GMRES + RightILUT on the DH matrices ----------
    os.makedirs(os.path.join(HERE, 'matrices'), exist_ok=True)
    for lev in range(16):
        coo = sp.coo_matrix(mmread(os.path.join(REF, 'TestMatrices', 'DH-Matrix-%d.mtx' % lev)))
        np.savez_compressed(os.path.join(HERE, 'matrices', 'DH-Matrix-%d.npz' % lev),
                            row=coo.row.astype(np.int32), col=coo.col.astype(np.int32),
                            data=coo.data, shape=np.asarray(coo.shape, dtype=np.int64))
        A = coo.tocsr()
        n = A.shape[0]
        b = A @ np.random.default_rng(2024).random(n)
        s = GMRES(CommonSolverArgs(maxiter=30, tau=1e-8), precond=RightILUT()).makeSolver()
        s.precond = None
        st, h = run(s, A, b)
        put('gmres_ilut/dh%d' % lev, st, h, keep_x=(lev <= 12))
        print('gmres+ilut dh%d: iters=%d success=%s' % (lev, st.iters(), st.success()))
    # un-preconditioned GMRES (MGS-sensitive) and LeftILUT no-op
    for lev in (5, 8):
        A = sp.coo_matrix(mmread(os.path.join(REF, 'TestMatrices', 'DH-Matrix-%d.mtx' % lev))).tocsr()
        b = A @ np.random.default_rng(2024).random(A.shape[0])
        s = GMRES(CommonSolverArgs(maxiter=100, tau=1e-8)).makeSolver()
        s.precond = None
        st, h = run(s, A, b)
        put('gmres/dh%d' % lev, st, h)
        s = GMRES(CommonSolverArgs(maxiter=100, tau=1e-8), precond=LeftILUT()).makeSolver()
        s.precond = None
        st, h = run(s, A, b)
        put('gmres_leftilut/dh%d' % lev, st, h)
    # known-answer assertions of the (stale) reference tests, on the current API:
    # tests/TestPCG.py:28-40 (IC-PCG, tau=1e-10) and tests/TestGMRES.py:28-40
    # (PCG + RightILUT, tau=1e-12), DH level 10, ||x - x_ex|| <= 1e-8.
    A = sp.coo_matrix(mmread(os.path.join(REF, 'TestMatrices', 'DH-Matrix-10.mtx'))).tocsr()
    xex = np.random.default_rng(99).random(A.shape[0])
    b = A @ xex
    st, h = run(PCG(CommonSolverArgs(maxiter=100, tau=1e-10), precond=RightIC()).makeSolver(), A, b)
    put('kat/pcg_ic_dh10', st, h)
    G['kat/pcg_ic_dh10/err'] = np.float64(np.linalg.norm(st.soln() - xex))
    st, h = run(PCG(CommonSolverArgs(maxiter=100, tau=1e-12), precond=RightILUT()).makeSolver(), A, b)
    put('kat/pcg_ilut_dh10', st, h)
    G['kat/pcg_ilut_dh10/err'] = np.float64(np.linalg.norm(st.soln() - xex))
    # ILUT apply + factors for DH-8
    A = sp.coo_matrix(mmread(os.path.join(REF, 'TestMatrices', 'DH-Matrix-8.mtx'))).tocsr()
    pre = quiet(RightILUT().form, A)
    v = np.random.default_rng(8).random(A.shape[0])
    G['ilut/apply_dh8_in'] = v
    G['ilut/apply_dh8_out'] = pre.applyRight(v)

    # ---------------- AMG hierarchy + V-cycle -----------------------------------
    for m, nlev in ((16, 2), (32, 2), (32, 3)):
        A = -quiet(FDLaplacian2D, 0.0, 1.0, m)
        mlh = quiet(SmoothedAggregationMLHierarchy, A, numLevels=nlev)
        tag = 'amg/m%d_L%d' % (m, nlev)
        for k in range(nlev):
            csr_put('%s/A%d' % (tag, k), mlh.matrix(k))
        for k in range(nlev - 1):
            csr_put('%s/P%d' % (tag, k), mlh.update(k))
            csr_put('%s/R%d' % (tag, k), mlh.downdate(k))
        aggs, nbrs = quiet(BuildAggregates, A, lvl=nlev - 1)
        G[tag + '/agg_sizes'] = np.asarray([len(a) for a in aggs], dtype=np.int64)
        G[tag + '/agg_flat'] = np.asarray([i for a in aggs for i in sorted(a)], dtype=np.int64)
        b = np.ones(A.shape[0])
        for sm_name, sm in (('gs', GaussSeidelSmoother), ('jac', JacobiSmoother),
                            ('djac', DampedJacobi)):
            s = AMGVCycle(CommonSolverArgs(maxiter=12, tau=1e-8, failOnMaxiter=False),
                          numLevels=nlev, smoother=sm).makeSolver()
            st, h = run(s, A, b)
            put('%s/vcycle_%s' % (tag, sm_name), st, h)
            pre = quiet(AMG(numIters=5, numLevels=nlev, smoother=sm).form, A)
            v = np.random.default_rng(11).random(A.shape[0])
            G['%s/apply_%s_in' % (tag, sm_name)] = v
            G['%s/apply_%s_out' % (tag, sm_name)] = quiet(pre.apply, v)
        for sm_name, sm in (('gs', GaussSeidelSmoother), ('djac', DampedJacobi)):
            s = PCG(CommonSolverArgs(maxiter=100, tau=1e-8),
                    precond=AMG(numIters=5, numLevels=nlev, smoother=sm)).makeSolver()
            st, h = run(s, A, b)
            put('%s/pcg_amg_%s' % (tag, sm_name), st, h)
            print('%s pcg+amg(%s): iters=%d' % (tag, sm_name, st.iters()))

    # hierarchies of irregular FE matrices (weak connections -> filtered-matrix lumping,
    # phase-2 tie breaking) and of a Bratu Jacobian
    extra = [('amg/dh7_L2', sp.coo_matrix(mmread(os.path.join(REF, 'TestMatrices', 'DH-Matrix-7.mtx'))).tocsr(), 2),
             ('amg/dh9_L3', sp.coo_matrix(mmread(os.path.join(REF, 'TestMatrices', 'DH-Matrix-9.mtx'))).tocsr(), 3)]
    fb = quiet(FDBratu2D, m=20)
    extra.append(('amg/bratuJ_m20_L2', fb.evalJ(np.linspace(0.5, 2.0, 400)), 2))
    rng = np.random.default_rng(5)
    W = sp.random(300, 300, density=0.02, random_state=rng)
    W = (W + W.T).tocsr()
    W.data = -np.abs(W.data) * rng.choice([1.0, 0.01], size=W.nnz)   # strong and weak couplings
    Wd = sp.diags(np.asarray(np.abs(W).sum(axis=1)).ravel() + 0.1)
    extra.append(('amg/rand300_L2', (W + Wd).tocsr(), 2))
    for tag, A, nlev in extra:
        mlh = quiet(SmoothedAggregationMLHierarchy, A, numLevels=nlev)
        csr_put(tag + '/Afine', A)
        for k in range(nlev):
            csr_put('%s/A%d' % (tag, k), mlh.matrix(k))
        for k in range(nlev - 1):
            csr_put('%s/P%d' % (tag, k), mlh.update(k))
            csr_put('%s/R%d' % (tag, k), mlh.downdate(k))
        print(tag, [mlh.matrix(k).shape[0] for k in range(nlev)])

    # ---------------- config 5 (small): Newton + PCG + AMG on Bratu -------------
    for m in (16, 32):
        for sm_name, sm in (('gs', GaussSeidelSmoother), ('djac', DampedJacobi)):
            func = quiet(FDBratu2D, m=m)
            lin_iters = []
            newton = NewtonSolver(control=CommonSolverArgs(tau=1.0e-12, maxiter=10),
                                  solver=PCG(precond=AMG(numIters=5, smoother=sm)),
                                  fixLinTol=False, minLinTol=1.0e-6, freezePrec=True)
            inner = newton.solver
            orig = inner.solve

            def spy(J, rhs, _orig=orig, _acc=lin_iters):
                r = _orig(J, rhs)
                _acc.append(r.iters())
                return r
            inner.solve = spy
            st, h = run(newton, func, func.initialU())
            put('newton/bratu_m%d_%s' % (m, sm_name), st, h)
            G['newton/bratu_m%d_%s/lin_iters' % (m, sm_name)] = np.asarray(lin_iters, dtype=np.int64)
            print('newton m=%d %s: iters=%d lin=%s' % (m, sm_name, st.iters(), lin_iters))
            # restore the shared default control object's tolerance
            CommonSolverArgs.__init__.__defaults__  # (defaults are per-def singletons)

    out = os.path.join(HERE, 'reference_golden.npz')
    np.savez_compressed(out, **G)
    print('wrote', out, '%.1f KB' % (os.path.getsize(out) / 1024.0), len(G), 'arrays')


if __name__ == '__main__':
    main()
