"""Run under torchrun (one rank per GPU): checks the row-partitioned SpMV and
PCG against scipy / the oracle.  Used by tests/test_gpu_dist.py and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29517 tests/dist_worker.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import krylov  # noqa: E402
from pysolvers_b200 import CommonSolverArgs  # noqa: E402
from pysolvers_b200 import dist as pdist  # noqa: E402
from pysolvers_b200.problems import fd_laplacian_2d, fd_laplacian_3d, load_dh_matrix  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def check_gmres_amg(comm, rank, world, fails):
    """Row-partitioned GMRES + AMG V-cycle (damped Jacobi) + Newton on Bratu against the SAME
    solvers on one GPU (this rank's, whole system): rectangular halo plans bit-exact vs scipy,
    V-cycle and GMRES histories to 1e-10, identical iteration counts."""
    from pysolvers_b200 import dist_krylov as dk
    from pysolvers_b200.Linear import AMG, GMRES, DampedJacobiSmoother
    from pysolvers_b200.Nonlinear import NewtonSolver
    from pysolvers_b200.problems import DeviceFDBratu2D, FDBratu2D

    def rel(a, b):
        a, b = np.asarray(a), np.asarray(b)
        k = min(len(a), len(b))
        return float(np.max(np.abs(a[:k] - b[:k]) / np.abs(b[:k]))) if k else 1.0

    for m, nlev in ((48, 2), (40, 3)):
        tag = 'bratu%d_L%d' % (m, nlev)
        J = sp.csr_matrix(FDBratu2D(m, 0.5).evalJ(np.linspace(0.5, 2.0, m * m)))
        n = J.shape[0]
        starts = pdist.row_starts(n, world)
        lo, hi = int(starts[rank]), int(starts[rank + 1])
        pre = dk.DistAMGPreconditioner(comm, J, numIters=5, numLevels=nlev, smoother=DampedJacobiSmoother)
        # (1) rectangular row blocks: restriction / prolongation products bit-identical to scipy
        for k in range(nlev - 1):
            R, P = sp.csr_matrix(pre.mlh.downdate(k)), sp.csr_matrix(pre.mlh.update(k))
            xf = np.random.default_rng(5 + k).standard_normal(R.shape[1])
            sf, sc = pre.starts[k + 1], pre.starts[k]
            y = pre.R[k].matvec(torch.from_numpy(xf[int(sf[rank]):int(sf[rank + 1])]).cuda()).cpu().numpy()
            if not np.array_equal(y, (R @ xf)[int(sc[rank]):int(sc[rank + 1])]):
                fails.append('%s: restriction block %d differs from scipy' % (tag, k))
            if pre.P[k] is not None:
                xc = np.random.default_rng(9 + k).standard_normal(P.shape[1])
                y = pre.P[k].matvec(torch.from_numpy(xc[int(sc[rank]):int(sc[rank + 1])]).cuda()).cpu().numpy()
                if not np.array_equal(y, (P @ xc)[int(sf[rank]):int(sf[rank + 1])]):
                    fails.append('%s: prolongation block %d differs from scipy' % (tag, k))
        # (2) V-cycle iteration and preconditioner application vs the single-GPU hierarchy
        b = np.random.default_rng(21).random(n)
        one = quiet(AMG(numIters=5, numLevels=nlev, smoother=DampedJacobiSmoother).form, J)
        dev1 = one.device_amg()
        x1, res1, h1 = dev1.solve(b, 8, 1e-10)
        xd, resd, hd = pre.solve(b[lo:hi], 8, 1e-10)
        if resd.n_hist != res1.n_hist or rel(hd, h1) > 1e-10:
            fails.append('%s: V-cycle history differs (%d vs %d cycles, rel %.2e)' % (tag, resd.n_hist, res1.n_hist, rel(hd, h1)))
        if np.linalg.norm(xd.cpu().numpy() - x1[lo:hi]) > 1e-10 * np.linalg.norm(x1):
            fails.append('%s: V-cycle solution differs' % tag)
        z1 = dev1.prec.apply_host(b)
        zd = pre.apply(torch.from_numpy(b[lo:hi]).cuda()).cpu().numpy()
        if np.linalg.norm(zd - z1[lo:hi]) > 1e-11 * np.linalg.norm(z1):
            fails.append('%s: preconditioner application differs (%.2e)' % (tag, np.linalg.norm(zd - z1[lo:hi]) / np.linalg.norm(z1)))
        # (3) GMRES + AMG
        ctl = dict(maxiter=60, tau=1e-9, showIters=False, showFinal=False)
        g1 = GMRES(CommonSolverArgs(**ctl), precond=AMG(numIters=5, numLevels=nlev, smoother=DampedJacobiSmoother)).makeSolver()
        s1 = quiet(g1.solve, J, b)
        gd = dk.DistributedGMRES(CommonSolverArgs(**ctl), precond=dk.DistAMG(comm, numIters=5, numLevels=nlev)).makeSolver()
        blk = J[lo:hi, :]
        D = pdist.DistCSR(comm, blk.indptr, blk.indices, blk.data, lo, hi, n, p2p=False)
        sd = quiet(gd.solve, dk.DistOperator(D, J), b[lo:hi])
        if sd.iters() != s1.iters() or sd.success() != s1.success():
            fails.append('%s: GMRES+AMG iters %s vs %s' % (tag, sd.iters(), s1.iters()))
        if rel(gd.last_history, g1.last_history) > 1e-9:
            fails.append('%s: GMRES+AMG history rel %.2e' % (tag, rel(gd.last_history, g1.last_history)))
        if np.linalg.norm(sd.soln().cpu().numpy() - s1.soln()[lo:hi]) > 1e-8 * np.linalg.norm(s1.soln()):
            fails.append('%s: GMRES+AMG solution differs' % tag)
        # un-preconditioned, MGS order
        ctl2 = dict(maxiter=40, tau=1e-6, showIters=False, showFinal=False)
        g1 = GMRES(CommonSolverArgs(**ctl2), orth='mgs').makeSolver()
        s1 = quiet(g1.solve, J, b)
        gd2 = dk.DistributedGMRES(CommonSolverArgs(**ctl2), orth='mgs').makeSolver()
        sd = quiet(gd2.solve, D, b[lo:hi])
        if sd.iters() != s1.iters() or rel(gd2.last_history, g1.last_history) > 1e-9:
            fails.append('%s: GMRES(mgs) iters %s vs %s, rel %.2e' % (tag, sd.iters(), s1.iters(), rel(gd2.last_history, g1.last_history)))
        if rank == 0:
            print('%s: levels %s  V-cycle hist rel %.1e  GMRES+AMG iters %d  GMRES(mgs) iters %d'
                  % (tag, [pre.mlh.matrix(k).shape[0] for k in range(nlev)], rel(hd, h1), len(gd.last_history),
                     len(gd2.last_history)), flush=True)
        D.close()
        gd.precond.close()
        pre.close()
    # (4) Newton + inexact GMRES + AMG with preconditioner reuse (configs[4]) on Bratu
    m = 48
    lin1, lind = [], []

    def spy(solver, acc):
        orig = solver.solve

        def f(J_, rhs):
            r = orig(J_, rhs)
            acc.append(r.iters())
            return r
        solver.solve = f
    nctl = dict(tau=1.0e-12, maxiter=10, showIters=False, showFinal=False)
    lctl = dict(maxiter=60, showIters=False, showFinal=False)
    n1 = NewtonSolver(control=CommonSolverArgs(**nctl),
                      solver=GMRES(CommonSolverArgs(**lctl), precond=AMG(numIters=5, smoother=DampedJacobiSmoother),
                                   honorFreeze=True),
                      fixLinTol=False, minLinTol=1.0e-6, freezePrec=True)
    spy(n1.solver, lin1)
    f1 = DeviceFDBratu2D(m=m)
    h1 = []
    n1.reportIter = lambda k, nr, nb: h1.append(nr)
    st1 = quiet(n1.solve, f1, f1.initialU())
    nd = NewtonSolver(control=CommonSolverArgs(norm=dk.dist_norm, **nctl),
                      solver=dk.DistributedGMRES(CommonSolverArgs(**lctl), precond=dk.DistAMG(comm, numIters=5)),
                      fixLinTol=False, minLinTol=1.0e-6, freezePrec=True)
    spy(nd.solver, lind)
    fd = dk.DistFDBratu2D(comm, m=m)
    hd = []
    nd.reportIter = lambda k, nr, nb: hd.append(nr)
    std = quiet(nd.solve, fd, fd.initialU())
    if std.success() != st1.success() or std.iters() != st1.iters() or lind != lin1:
        fails.append('newton: iters %s lin %s vs 1 GPU iters %s lin %s' % (std.iters(), lind, st1.iters(), lin1))
    # ||F|| ends at the rounding floor of evaluating F (~1e-13 ||F_0||): compare down to that floor
    ha, hb = np.asarray(hd), np.asarray(h1)
    if ha.shape != hb.shape or np.any(np.abs(ha - hb) > 1e-8 * hb + 1e-11 * hb[0]):
        fails.append('newton: ||F|| history differs %s vs %s' % (ha, hb))
    lo, hi = fd.lo, fd.hi
    u1 = st1.soln().cpu().numpy() if hasattr(st1.soln(), 'cpu') else st1.soln()
    if np.linalg.norm(std.soln().cpu().numpy() - u1[lo:hi]) > 1e-8 * np.linalg.norm(u1):
        fails.append('newton: solution differs')
    if rank == 0:
        print('newton bratu m=%d: iters %d lin %s (1 GPU: %d %s) ||F|| rel %.1e' % (m, std.iters(), lind, st1.iters(), lin1, rel(hd, h1)), flush=True)
    fd.close()


def main():
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    comm = pdist.Comm()
    fails = []
    cases = [('lap2d', -fd_laplacian_2d(0.0, 1.0, 96), 1e-8, 600),
             ('lap3d', fd_laplacian_3d(0.0, 1.0, 24), 1e-9, 400),
             ('dh12', load_dh_matrix(12), 1e-8, 60)]
    for name, A, tau, maxiter in cases:
        A = sp.csr_matrix(A)
        n = A.shape[0]
        starts = pdist.row_starts(n, world)
        lo, hi = int(starts[rank]), int(starts[rank + 1])
        blk = A[lo:hi, :]
        D = pdist.DistCSR(comm, blk.indptr, blk.indices, blk.data, lo, hi, n)
        x = np.random.default_rng(3).standard_normal(n)
        y = D.matvec(torch.from_numpy(x[lo:hi]).cuda()).cpu().numpy()
        if not np.array_equal(y, (A @ x)[lo:hi]):
            fails.append('%s: dist spmv differs from scipy' % name)
        b = np.ones(n) if name != 'dh12' else A @ np.random.default_rng(1).random(n)
        solver = pdist.DistributedPCG(CommonSolverArgs(maxiter=maxiter, tau=tau))
        hist = []
        solver.reportIter = lambda k, nr, nb: hist.append(nr)
        with contextlib.redirect_stdout(io.StringIO()):
            st = solver.solve(D, b[lo:hi])
        ref = krylov.pcg(A, b, maxiter=maxiter, tau=tau)
        with contextlib.redirect_stdout(io.StringIO()):
            st_again = solver.solve(D, b[lo:hi])       # epochs / ping-pong buffers carry over
        if st_again.iters() != st.iters() or not np.array_equal(st_again.soln(), st.soln()):
            fails.append('%s: second solve differs from the first' % name)
        hist = np.asarray(hist)
        k = min(len(hist), len(ref['hist']))
        if name == 'dh12':
            # ill-conditioned + un-preconditioned: compare while the history is meaningful
            k = min(k, 20)
        rel = np.max(np.abs(hist[:k] - ref['hist'][:k]) / ref['hist'][:k])
        if st.success() != ref['success'] and name != 'dh12':
            fails.append('%s: success %s vs %s' % (name, st.success(), ref['success']))
        if name != 'dh12' and abs(st.iters() - ref['iters']) > 1:
            fails.append('%s: iters %d vs %d' % (name, st.iters(), ref['iters']))
        if rel > 1e-10:
            fails.append('%s: history rel err %.3e' % (name, rel))
        if name != 'dh12':
            err = np.linalg.norm(st.soln() - ref['soln'][lo:hi]) / np.linalg.norm(ref['soln'][lo:hi])
            if err > 1e-8:
                fails.append('%s: solution rel err %.3e' % (name, err))
        if rank == 0:
            print('%s: n=%d world=%d iters=%d (oracle %d) hist rel %.2e halo=%d r0=%d r1=%d p2p=%s'
                  % (name, n, world, st.iters(), ref['iters'], rel, D.n_halo, D.r0, D.r1, D.p2p), flush=True)
        del D
    check_gmres_amg(comm, rank, world, fails)
    flag = torch.tensor([len(fails)], device='cuda')
    dist.all_reduce(flag)
    if fails:
        print('rank %d FAILURES: %s' % (rank, fails), flush=True)
    comm.close()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == '__main__':
    main()
