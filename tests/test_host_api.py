"""Host-side mirror of the reference API, on CPU: control / result protocol, report formats,
the preconditioner protocol, the Newton driver with the direct-solver passthrough, and the
error behaviour of the device solvers when the GPU path cannot serve a request."""
import contextlib
import io
import os
import sys

import numpy as np
import numpy.linalg as npla
import pytest
import scipy.sparse as sp

import pysolvers_b200 as P
from pysolvers_b200 import CommonSolverArgs, SolveStatus
from pysolvers_b200.core import IterativeSolver
from pysolvers_b200.Linear import (PCG, GMRES, PCGSolver, IdentityPreconditioner, IdentityPreconditionerType,
                                   LeftPreconditioner, RightPreconditioner, GenericPreconditioner,
                                   PreconditionerType, DefaultDirect, AMGVCycle, LinearSolver)
from pysolvers_b200.Linear.precond import right_handle_of
from pysolvers_b200.Nonlinear import NewtonSolver, FuncAdapter1D, SimpleBacktrack, PreconditionerFreeze


def _out(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def test_package_surface_matches_reference():
    # PySolvers/__init__.py:1-3 and PySolvers/Linear/__init__.py:1-12
    for name in ('Linear', 'Nonlinear', 'CommonSolverArgs'):
        assert hasattr(P, name)
    for name in ('IterativeLinearSolver', 'mvmult', 'DefaultDirect', 'DefaultDirectSolver', 'PCG', 'PCGSolver',
                 'GMRES', 'GMRESSolver', 'IdentityPreconditionerType', 'IdentityPreconditioner',
                 'ICRightPreconditioner', 'RightIC', 'LeftILUT', 'RightILUT', 'AMGVCycle', 'AMGVCycleSolver',
                 'AMGPreconditioner', 'AMG'):
        assert hasattr(P.Linear, name), name
    for name in ('NewtonSolver', 'FuncAdapter1D'):
        assert hasattr(P.Nonlinear, name)


def test_control_defaults_and_shared_default_instance():
    c = CommonSolverArgs()
    assert (c.maxiter, c.failOnMaxiter, c.tau, c.showIters, c.showFinal, c.interval) == (100, True, 1e-8, True, True, 1)
    assert c.norm is npla.norm
    # def-time singleton shared between solvers built with defaults
    a, b = PCG().makeSolver(), PCG().makeSolver()
    assert a._control is b._control
    old = a.tau()
    a.setTolerance(1e-3)
    assert b.tau() == 1e-3
    a.setTolerance(old)


def test_exit_conventions_and_report_text():
    s = IterativeSolver(CommonSolverArgs(maxiter=7), name='S')
    st, text = _out(s.handleConvergence, 3, 'x', 1e-9, 2.0)
    assert (st.success(), st.iters(), st.soln(), st.resid(), st.msg()) == (True, 4, 'x', 1e-9, None)
    assert text == 'S solve succeeded: iters=%7d, ||r||/r0=%12.5g\n' % (4, 1e-9 / 2.0)
    st, text = _out(s.handleBreakdown, 2, 'why')
    assert (st.success(), st.iters(), st.soln(), st.resid(), st.msg()) == (False, 2, None, None, 'why')
    assert text == 'S solve broke down: why\n'
    st, text = _out(s.handleMaxiter, 6, 'x', 0.5, 2.0)
    assert (st.success(), st.iters(), st.msg()) == (False, 6, 'failure to converge')
    assert text == 'S solve FAILED: iters=%7d, ||r||/r0=%12.5g\n' % (6, 0.25)
    s2 = IterativeSolver(CommonSolverArgs(maxiter=7, failOnMaxiter=False), name='S')
    st, _ = _out(s2.handleMaxiter, 6, 'x', 0.5, 2.0)
    assert (st.success(), st.iters()) == (True, 6)
    _, text = _out(s.reportIter, 4, 3.0, 6.0)
    assert text == 'S iter=%7d ||r||=%12.5g ||r||/r0=%12.5g\n' % (4, 3.0, 0.5)
    quiet = IterativeSolver(CommonSolverArgs(showIters=False, showFinal=False), name='S')
    assert _out(quiet.reportIter, 0, 1.0, 1.0)[1] == '' and _out(quiet.handleConvergence, 0, 0, 1.0, 1.0)[1] == ''
    assert str(SolveStatus(True, None, 0.5, 3)) == 'SolverState(success=True, resid=0.5, iters=3)'


def test_preconditioner_protocol():
    v = np.arange(3.0)
    ident = IdentityPreconditionerType().form(None)
    assert ident.applyLeft(v) is v and ident.applyRight(v) is v and right_handle_of(ident) is None

    class MyLeft(LeftPreconditioner):
        def applyLeft(self, vec):
            return 2 * vec
    assert MyLeft().applyRight(v) is v and right_handle_of(MyLeft()) is None

    class HostOnly(RightPreconditioner):
        def applyRight(self, vec):
            return vec / 2
    with pytest.raises(NotImplementedError):
        right_handle_of(HostOnly())            # no CPU fallback on the solve path

    class Unknown:
        def applyRight(self, vec):
            return vec
    with pytest.raises(NotImplementedError):
        right_handle_of(Unknown())


def test_shape_asserts_and_unsupported_norm():
    A = sp.identity(4, format='csr')
    with pytest.raises(AssertionError):
        PCG().makeSolver().solve(sp.csr_matrix((3, 4)), np.ones(3))
    with pytest.raises(AssertionError):
        GMRES().makeSolver().solve(A, np.ones(5))
    s = PCG(CommonSolverArgs(norm=lambda x: np.max(np.abs(x)))).makeSolver()
    with pytest.raises(NotImplementedError):
        s.solve(A, np.ones(4))


def test_missing_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from pysolvers_b200 import _native
    with pytest.raises(_native.NativeError):
        _out(PCG().makeSolver().solve, sp.identity(4, format='csr'), np.ones(4))


class _Root2(FuncAdapter1D):
    def _evalF(self, x):
        return x * x - 2

    def _evalJ(self, x):
        return 2.0 * x


class _ArcTan(FuncAdapter1D):
    def _evalF(self, x):
        return np.arctan(x)

    def _evalJ(self, x):
        return 1.0 / (1.0 + x * x)


def test_newton_examples_with_direct_solver():
    """examples/NewtonExample_Root2.py and NewtonExample_ArcTan.py (CPU: DefaultDirect)."""
    solver = NewtonSolver(control=CommonSolverArgs(tau=1.0e-15, maxiter=10), solver=DefaultDirect())
    st, text = _out(solver.solve, _Root2(), np.array([3.0]))
    assert st.success() and abs(st.soln()[0] - np.sqrt(2.0)) < 1e-14
    assert 'Newton iter=' in text
    solver = NewtonSolver(control=CommonSolverArgs(tau=1.0e-15, maxiter=10), solver=DefaultDirect(), freezePrec=False)
    st, text = _out(solver.solve, _ArcTan(), np.array([10.0]))
    assert st.success() and abs(st.soln()[0]) < 1e-14
    assert 'k=   0 t=' in text                       # the line search had to backtrack from x0 = 10


def test_newton_linear_failure_is_a_breakdown():
    class Failing(LinearSolver):
        def solve(self, A, b):
            return SolveStatus(False, None, None, None, 'nope')

    class Type:
        def makeSolver(self):
            return Failing()
    st, _ = _out(NewtonSolver(solver=Type()).solve, _Root2(), np.array([3.0]))
    assert not st.success() and 'solve for Newton step failed with msg=nope' in st.msg()


def test_preconditioner_freeze_never_unfreezes():
    s = PCG().makeSolver()
    PreconditionerFreeze(s, True)
    assert s.precFrozen()                             # and stays so: __def__ typo reproduced
    s.unfreezePrec()
    PreconditionerFreeze(s, False)
    assert not s.precFrozen()


def test_backtracking_line_search():
    ls = SimpleBacktrack(report=False)
    ls.setNorm(npla.norm)
    ok, x, F, nF = ls.search(np.array([10.0]), abs(np.arctan(10.0)), np.array([-10.0 * 101 * np.arctan(10.0) / 10.0]), _ArcTan())
    assert ok and nF < abs(np.arctan(10.0))


def test_dense_block_by_height_keeps_factors_triangular():
    """Host logic of the split LU (no GPU): psb_tri_levels equals the numpy level analysis; moving the
    rows at the top of the elimination tree to the end keeps L and U triangular; the rest has no more
    levels than the same number of trailing rows would leave; the permuted solve is the same solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from oracle import precond
    from pysolvers_b200.device import _dense_block_by_level, tri_heights_upper, tri_levels
    from pysolvers_b200.problems import fd_laplacian_2d
    rng = np.random.default_rng(8)
    n = 30 * 30
    A = sp.csc_matrix(-fd_laplacian_2d(0.0, 1.0, 30)) + sp.random(n, n, density=0.003, random_state=rng, format='csc')
    lu = spla.splu(A, permc_spec='MMD_AT_PLUS_A')
    L, U = lu.L.tocsr(), lu.U.tocsr()
    assert np.array_equal(tri_levels(L, True), precond.level_sets(L, lower=True)[0])
    assert np.array_equal(tri_levels(U, False), precond.level_sets(U, lower=False)[0])
    assert np.array_equal(tri_heights_upper(U), precond.level_sets(sp.csr_matrix(U.T), lower=True)[0])
    ipr = np.empty(n, dtype=np.int64)
    ipr[lu.perm_r] = np.arange(n)
    ipc = np.empty(n, dtype=np.int64)
    ipc[lu.perm_c] = np.arange(n)
    for tail in (1, 5, 120, n - 1):
        qL, n1L = _dense_block_by_level(L, True, tail)
        qU, n1U = _dense_block_by_level(U, False, tail)
        assert sorted(qL.tolist()) == list(range(n)) and sorted(qU.tolist()) == list(range(n))
        assert 1 <= n - n1L <= tail and 1 <= n - n1U <= tail
        Lp, Up = L[qL][:, qL].tocsr(), U[qU][:, qU].tocsr()
        assert sp.triu(Lp, 1).nnz == 0 and sp.tril(Up, -1).nnz == 0
        if n1L > 0 and n - n1L > 1:
            by_height = tri_levels(Lp[:n1L, :n1L], True).max() + 1
            trailing = tri_levels(L[:n1L, :n1L], True).max() + 1
            assert by_height <= trailing
        v = rng.standard_normal(n)
        y = spla.spsolve_triangular(Lp, v[ipr[qL]], lower=True, unit_diagonal=True)
        posL = np.empty(n, dtype=np.int64)
        posL[qL] = np.arange(n)
        x = spla.spsolve_triangular(Up, y[posL[qU]], lower=False)
        out = np.empty(n)
        out[ipc[qU]] = x
        ref = lu.solve(v)
        assert np.linalg.norm(out - ref) <= 1e-12 * np.linalg.norm(ref)


def test_bench_reference_arm_contract():
    """``bench.py --impl reference`` (the reference's CPU path, oracle port) prints one JSON line with
    the contract's keys; run here on a small grid so that it takes seconds."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PSB_BENCH_M='192', PSB_REF_ITERS_PER_STEP='3')
    p = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '2',
                        '--warmup', '1'], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'pcg_iterations_per_second'
    assert line['unit'] == 'iter/s' and line['higher_is_better'] is True and line['value'] > 0
    assert line['steps'] == 2 and line['warmup'] == 1 and line['n_gpus'] == 1
    assert line['e2e']['value'] == line['value'] and line['e2e']['h2d_bytes_per_step'] == 0
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == line['value'] and 'sample' in cb
    assert 'workload' in line['config']
    # ranks other than 0 of a torchrun launch exit 0 without output
    p2 = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference'], capture_output=True,
                        text=True, env=dict(env, RANK='1', WORLD_SIZE='2'), timeout=120)
    assert p2.returncode == 0 and p2.stdout.strip() == ''
