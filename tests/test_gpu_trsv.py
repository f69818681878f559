"""Sparse triangular solves and the IC / ILUT preconditioners on the B200.

Level sets must match the numpy oracle bit-exactly (north_star); the solve
follows the row-wise stored-order summation of oracle.precond.trsv_rowwise, so
on small cases it is compared bit for bit with that restatement and to
rounding with scipy's spsolve_triangular / SuperLU.solve (what the reference
calls, ICPreconditioner.py:58-63, ILUTPreconditioner.py:66-78)."""
import contextlib
import io

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import rel_err, golden_csr

pytestmark = pytest.mark.gpu


def _lap(m):
    from pysolvers_b200.problems import fd_laplacian_2d
    return -fd_laplacian_2d(0.0, 1.0, m)


def _factors(m):
    from oracle import precond
    return precond.ic_factor(_lap(m))


@pytest.mark.parametrize('m', [8, 24])
def test_levels_bit_exact_and_solve_matches_rowwise(cuda, m):
    from oracle import precond
    from pysolvers_b200.device import DeviceTrsv, to_device
    L, Lt = _factors(m)
    v = np.random.default_rng(m).standard_normal(L.shape[0])
    for T, lower in ((L, True), (Lt, False)):
        dT = DeviceTrsv(T, lower=lower)
        level, lptr, lrows = precond.level_sets(T, lower=lower)
        got_ptr, got_rows = dT.levels()
        assert np.array_equal(got_ptr, lptr)
        assert np.array_equal(got_rows, lrows)
        assert dT.info()['levels'] == len(lptr) - 1
        x = dT.solve(to_device(v)).cpu().numpy()
        dT.check()
        assert np.array_equal(x, precond.trsv_rowwise(T, v, lower=lower))
        ref = spla.spsolve_triangular(T, v, lower=lower)
        assert np.linalg.norm(x - ref) <= 1e-13 * np.linalg.norm(ref)


def test_trsv_edge_cases(cuda):
    from pysolvers_b200.device import DeviceTrsv, to_device
    rng = np.random.default_rng(0)
    # diagonal only (one level), unit lower with explicit ones, single row, dense-ish random lower
    cases = []
    cases.append((sp.diags(rng.random(77) + 1.0).tocsr(), True, False))
    Lr = sp.tril(sp.random(300, 300, density=0.05, random_state=rng), k=-1) + sp.identity(300)
    cases.append((Lr.tocsr(), True, True))
    cases.append((sp.csr_matrix(np.array([[2.5]])), True, False))
    Ur = sp.triu(sp.random(257, 257, density=0.1, random_state=rng), k=1) + sp.diags(rng.random(257) + 2.0)
    cases.append((Ur.tocsr(), False, False))
    # long rows (warp-per-row path): dense triangles and an exact sparse LU with fill
    Ld = np.tril(rng.random((150, 150))) + 150.0 * np.eye(150)
    cases.append((sp.csr_matrix(Ld), True, False))
    cases.append((sp.csr_matrix(Ld.T), False, False))
    from pysolvers_b200.problems import fd_laplacian_2d
    lu = spla.splu(sp.csc_matrix(-fd_laplacian_2d(0.0, 1.0, 40)))
    cases.append((lu.L.tocsr(), True, True))
    cases.append((lu.U.tocsr(), False, False))
    # long chain: bidiagonal -> n levels with one row each
    n = 2000
    chain = sp.diags([np.full(n - 1, -0.5), np.full(n, 1.5)], [-1, 0]).tocsr()
    cases.append((chain, True, False))
    for T, lower, unit in cases:
        v = rng.standard_normal(T.shape[0])
        dT = DeviceTrsv(T, lower=lower, unit_diag=unit)
        x = dT.solve(to_device(v)).cpu().numpy()
        dT.check()
        ref = spla.spsolve_triangular(T.tocsr(), v, lower=lower, unit_diagonal=unit)
        assert np.linalg.norm(x - ref) <= 1e-11 * np.linalg.norm(ref), (T.shape, lower, unit)
    assert DeviceTrsv(chain, lower=True).info()['levels'] == n


def _both_kernels(T, lower, unit, v):
    """Solve with the grid-wide kernel (hand-over through L2), the one-CTA kernel (hand-over through
    the shared-memory window) and, where the analysis allows it, the 4-CTA cluster kernel (window
    replicated through distributed shared memory): the results must be the same bits."""
    from pysolvers_b200.device import DeviceTrsv, to_device
    dT = DeviceTrsv(T, lower=lower, unit_diag=unit)
    out = {}
    kernels = ['grid', 'cta'] + (['cluster'] if dT.info2()['cluster_ok'] else [])
    for kern in kernels:
        dT.set_kernel(kern)
        assert dT.info2()['forced'] == {'grid': 0, 'cta': 1, 'cluster': 2}[kern]
        xs = [dT.solve(to_device(v)).cpu().numpy() for _ in range(2)]    # twice: state carried over
        dT.check()
        assert np.array_equal(xs[0], xs[1])
        out[kern] = xs[0]
    for kern in kernels[1:]:
        assert np.array_equal(out['grid'], out[kern]), kern
    return dT, out['cta']


def test_both_kernels_bit_identical(cuda, monkeypatch):
    from oracle import precond
    # one packing for all three kernels: without the grid-only 8-lanes-per-row chunks that factors with
    # wide levels get by default (test_wide_levels_subwarp_rows covers those)
    monkeypatch.setenv('PSB_TRSV_NO_SUBWARP', '1')
    rng = np.random.default_rng(5)
    cases = []
    for m in (24, 140):          # 140^2 = 19 600 rows: the window wraps around
        L, Lt = _factors(m)
        cases += [(L, True, False), (Lt, False, False)]
    # random dependencies all over the vector: most of them are "far" (older than the window)
    n = 40000
    R = sp.tril(sp.random(n, n, density=3.0 / n, random_state=rng), k=-1) + sp.diags(rng.random(n) + 2.0)
    cases.append((R.tocsr(), True, False))
    band = sp.diags([np.full(n - 1, -0.4), np.full(n - 9000, 0.1), np.full(n - 15000, 0.2), np.full(n, 1.5)],
                    [-1, -9000, -15000, 0]).tocsr()          # a chain with far dependencies on it
    cases.append((band, True, False))
    cases.append((band.T.tocsr(), False, False))
    # wide levels (Gauss-Seidel triangle of a 200 x 200 grid: 100 rows per level on average, 40 000
    # rows): analysed for the cluster kernel, the window wraps around
    gs = sp.triu(_lap(200)).tocsr()
    cases.append((gs, False, False))
    lu2 = spla.splu(sp.csc_matrix(_lap(132)))                 # 17 424 rows: long rows AND a wrapping window
    cases += [(lu2.L.tocsr(), True, True), (lu2.U.tocsr(), False, False)]
    lu = spla.splu(sp.csc_matrix(_lap(48)))                   # long rows (warp per row)
    cases += [(lu.L.tocsr(), True, True), (lu.U.tocsr(), False, False)]
    for T, lower, unit in cases:
        v = rng.standard_normal(T.shape[0])
        dT, x = _both_kernels(T, lower, unit, v)
        i2 = dT.info2()
        assert T.shape[0] <= i2['wslots'] <= 16384 or i2['wslots'] in (8192, 16384)
        if T.shape[0] <= 20000 and all(T is not c[0] for c in cases[-4:]):
            # thread-per-row chunks follow the row-wise restatement bit for bit (a long row summed by
            # a whole warp does not: shuffle tree)
            assert np.array_equal(x, precond.trsv_rowwise(T, v, lower=lower, unit_diagonal=unit))
        ref = spla.spsolve_triangular(T.tocsr(), v, lower=lower, unit_diagonal=unit)
        assert np.linalg.norm(x - ref) <= 1e-11 * np.linalg.norm(ref), (T.shape, lower, unit)
    from pysolvers_b200.device import DeviceTrsv as _DT
    i3 = _DT(gs, lower=False).info2()
    assert i3['cluster_ok'] and i3['kernel'] == 'cluster' and i3['wslots'] == 16384
    # the banded case must really exercise the far path
    from pysolvers_b200.device import DeviceTrsv
    i2 = DeviceTrsv(band, lower=True).info2()
    assert i2['n_far'] > 0 and i2['max_dist'] >= 15000


def test_wide_levels_subwarp_rows(cuda):
    """Factors with WIDE levels and medium / long rows (the leading blocks of the AMG coarse LU factors):
    the analysis packs them for the grid kernel with a thread per row up to 8 entries, 8 lanes per row up
    to 128, a warp per row beyond.  Against scipy to rounding, deterministic, and -- with the sub-warp
    classes switched off (thread per row up to 32 entries) -- equal to the old packing to rounding."""
    from pysolvers_b200.device import DeviceTrsv, to_device
    rng = np.random.default_rng(17)
    n_lev, per = 24, 2500
    n = n_lev * per
    rows, cols, vals = [], [], []
    for i in range(per, n):
        lv = i // per
        k = int(rng.choice([3, 7, 12, 30, 50, 64, 65, 150, 400], p=[.2, .15, .15, .15, .1, .05, .05, .1, .05]))
        c = rng.choice(lv * per, size=min(k, lv * per), replace=False)
        c[0] = (lv - 1) * per + int(rng.integers(per))         # at least one dependency on the previous level
        c = np.unique(c)
        rows.append(np.full(c.size, i)); cols.append(c); vals.append(rng.standard_normal(c.size) / (4.0 * c.size))
    L = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)) \
        + sp.diags(rng.random(n) + 1.0)
    L = L.tocsr()
    v = rng.standard_normal(n)
    for T, lower in ((L, True), (L.T.tocsr(), False)):
        dT = DeviceTrsv(T, lower=lower)
        info = dT.info()
        assert info['levels'] == n_lev and dT.info2()['kernel'] == 'grid'
        x1 = dT.solve(to_device(v)).cpu().numpy()
        x2 = dT.solve(to_device(v)).cpu().numpy()
        dT.check()
        assert np.array_equal(x1, x2)
        ref = spla.spsolve_triangular(T, v, lower=lower)
        assert np.linalg.norm(x1 - ref) <= 1e-12 * np.linalg.norm(ref)
        with pytest.raises(Exception):
            dT.set_kernel('cta')                                 # 8-lanes-per-row chunks are the grid kernel's format
        # per-chunk time stamps of the grid kernel (psb_trsv_set_trace, tools/trsv_levels.py): every chunk
        # claimed before it is done, its first item inside the level-major order, no level done before
        # the chunks it depends on were claimed; the traced solve gives the same bits
        import ctypes as C
        import torch
        from pysolvers_b200 import _native as nat
        from pysolvers_b200.device import ptr
        g = info['groups']
        buf = torch.zeros(3 * g, dtype=torch.int64, device='cuda')
        nat.check(nat.lib().psb_trsv_set_trace(dT.handle, ptr(buf)), 'psb_trsv_set_trace')
        x3 = dT.solve(to_device(v)).cpu().numpy()
        nat.check(nat.lib().psb_trsv_set_trace(dT.handle, None), 'psb_trsv_set_trace')
        assert np.array_equal(x3, x1)
        t = buf.cpu().numpy().reshape(g, 3)
        assert np.all(t[:, 0] > 0) and np.all(t[:, 1] >= t[:, 0])
        lp = np.zeros(n_lev + 1, dtype=np.int32)
        lr = np.zeros(n, dtype=np.int32)
        nat.check(nat.lib().psb_trsv_get_levels(dT.handle, lp.ctypes.data_as(C.c_void_p), lr.ctypes.data_as(C.c_void_p)), 'levels')
        lev = np.searchsorted(lp, t[:, 2], side='right') - 1
        assert lev.min() == 0 and lev.max() == n_lev - 1 and np.all(np.diff(lev) >= 0)     # chunks are level-major
        # (no assertion on the order of stamps taken on different SMs: %globaltimer advances in steps of
        # a fraction of a microsecond, the same size as a hand-over)
        done = np.array([t[lev == l, 1].max() for l in range(n_lev)])
        assert done.max() - t[:, 0].min() < 1e9                 # nanoseconds: one solve, well under a second


def test_ic_apply_vs_reference_golden(cuda, golden):
    from pysolvers_b200.Linear import RightIC
    A = _lap(32)
    pre = RightIC().form(A)
    gL = golden_csr(golden, 'ic/L_m32')
    gLt = golden_csr(golden, 'ic/Lt_m32')
    # the host setup is the reference's: same factors, same stored order, bit for bit
    for mine, ref in ((pre._L, gL), (pre._Lt, gLt)):
        assert np.array_equal(mine.indptr, ref.indptr)
        assert np.array_equal(mine.indices, ref.indices)
        assert np.array_equal(mine.data, ref.data)
    v = golden['ic/apply_m32_in']
    out = pre.applyRight(v)
    assert rel_err(out, golden['ic/apply_m32_out']) < 1e-11
    assert pre.applyLeft(v) is v


def test_ilut_apply_vs_reference_golden(cuda, golden):
    from pysolvers_b200.Linear import RightILUT, LeftILUT
    from pysolvers_b200.problems import load_dh_matrix
    A = load_dh_matrix(8)
    v = golden['ilut/apply_dh8_in']
    pre = RightILUT().form(A)
    out = pre.applyRight(v)
    g = golden['ilut/apply_dh8_out']
    assert np.linalg.norm(out - g) <= 1e-12 * np.linalg.norm(g)
    left = LeftILUT().form(A)
    assert left.applyRight(v) is v                       # no-op from the right
    assert left.right_device_handle() is None
    assert np.linalg.norm(left.applyLeft(v) - g) <= 1e-12 * np.linalg.norm(g)


def _run(solver, A, b):
    hist = []
    solver.reportIter = lambda k, nr, nb: hist.append(nr)
    with contextlib.redirect_stdout(io.StringIO()):
        st = solver.solve(A, b)
    return st, np.asarray(hist)


@pytest.mark.parametrize('m', [16, 32, 64])
def test_icpcg_history_vs_reference_golden(cuda, golden, m):
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import PCG, RightIC
    A = _lap(m)
    st, hist = _run(PCG(CommonSolverArgs(maxiter=500, tau=1e-8), precond=RightIC()).makeSolver(),
                    A, np.ones(A.shape[0]))
    key = 'icpcg/lap2d_m%d' % m
    assert st.success() and abs(st.iters() - int(golden[key + '/iters'])) <= 1
    k = min(len(hist), len(golden[key + '/hist']))
    assert rel_err(hist[:k], golden[key + '/hist'][:k]) < 1e-10
    gx = golden[key + '/x']
    assert np.linalg.norm(st.soln() - gx) <= 1e-8 * np.linalg.norm(gx)


@pytest.mark.parametrize('m', [128, 256])
def test_icpcg_history_at_size_vs_reference(cuda, golden_large, m):
    """IC-PCG at m = 128 / 256 against histories the reference itself produced
    (tests/golden/make_golden_large.py): 1e-10 per iteration, iterations +-1, solution 1e-8."""
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import PCG, RightIC
    A = _lap(m)
    st, hist = _run(PCG(CommonSolverArgs(maxiter=500, tau=1e-8), precond=RightIC()).makeSolver(),
                    A, np.ones(A.shape[0]))
    key = 'icpcg/lap2d_m%d' % m
    assert st.success() and abs(st.iters() - int(golden_large[key + '/iters'])) <= 1
    k = min(len(hist), len(golden_large[key + '/hist']))
    assert rel_err(hist[:k], golden_large[key + '/hist'][:k]) < 1e-10
    gx = golden_large[key + '/x']
    assert np.linalg.norm(st.soln()[::97] - gx) <= 1e-8 * np.linalg.norm(gx)


def test_known_answer_dh10(cuda, golden):
    """The assertions of the reference's own (stale) tests on the current API:
    tests/TestPCG.py:28-40 and tests/TestGMRES.py:28-40 -- ||x - x_ex|| <= 1e-8."""
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import PCG, RightIC, RightILUT
    from pysolvers_b200.problems import load_dh_matrix
    A = load_dh_matrix(10)
    xex = np.random.default_rng(99).random(A.shape[0])
    b = A @ xex
    st, h = _run(PCG(CommonSolverArgs(maxiter=100, tau=1e-10), precond=RightIC()).makeSolver(), A, b)
    assert st.success() and np.linalg.norm(st.soln() - xex) <= 1e-8
    assert abs(st.iters() - int(golden['kat/pcg_ic_dh10/iters'])) <= 1
    st, h = _run(PCG(CommonSolverArgs(maxiter=100, tau=1e-12), precond=RightILUT()).makeSolver(), A, b)
    assert st.success() and np.linalg.norm(st.soln() - xex) <= 1e-8
    assert abs(st.iters() - int(golden['kat/pcg_ilut_dh10/iters'])) <= 1


def test_prec_frozen_reuse(cuda):
    """PCG keeps the formed preconditioner while frozen (PCGSolver.py:92-94)."""
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import PCG, RightIC
    A = _lap(12)
    s = PCG(CommonSolverArgs(maxiter=100, tau=1e-8), precond=RightIC()).makeSolver()
    _run(s, A, np.ones(A.shape[0]))
    first = s.precond
    s.freezePrec()
    _run(s, A, np.ones(A.shape[0]))
    assert s.precond is first
    s.unfreezePrec()
    _run(s, A, np.ones(A.shape[0]))
    assert s.precond is not first


def test_rows_longer_than_the_staging_buffer(cuda):
    """Wide levels AND long rows: the upper triangle of a 4 x 6-point stencil on a 200 x 200 grid
    (23 dependencies per row, ~100 rows per level, 40 000 rows).  The analysis picks the cluster
    configuration (16 384-slot window, 15 staged entries per lane), so every chunk needs two staging
    rounds -- in all three kernels the same bits, and the same as the row-wise restatement."""
    from oracle import precond
    from pysolvers_b200.device import DeviceTrsv
    m = 200
    rng = np.random.default_rng(31)
    B1 = sp.diags([np.full(m - a, 1.0) for a in range(4)], list(range(4)), shape=(m, m))
    B2 = sp.diags([np.full(m - b, 1.0) for b in range(6)], list(range(6)), shape=(m, m))
    U = sp.kron(B2, B1).tocsr()
    U.data = -0.04 * rng.random(U.nnz) - 0.001
    U.setdiag(1.0 + rng.random(m * m))
    U = U.tocsr()
    assert sp.tril(U, -1).nnz == 0 and np.diff(U.indptr).max() == 24
    v = rng.standard_normal(m * m)
    dT, x = _both_kernels(U, False, False, v)
    i2 = dT.info2()
    assert i2['cluster_ok'] and i2['wslots'] == 16384 and i2['stage_len'] < 23
    assert dT.info()['levels'] == 2 * m - 1
    assert np.array_equal(x, precond.trsv_rowwise(U, v, lower=False))
    L = sp.csr_matrix(U.T)
    dT2, x2 = _both_kernels(L, True, False, v)
    assert np.array_equal(x2, precond.trsv_rowwise(L, v, lower=True))
