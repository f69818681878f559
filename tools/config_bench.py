"""Measurements for the BASELINE.json configurations other than the bench.py headline
(C1, C2, IC-PCG, C5) on one B200, with the oracle port timed on the host beside them.

    python tools/config_bench.py [c1] [c2] [ic512] [ic1024] [c5_512] [c5_2048] > gpurun_out/configs.json
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import krylov, precond  # noqa: E402
from pysolvers_b200 import CommonSolverArgs, _native as nat  # noqa: E402
from pysolvers_b200.Linear import (PCG, GMRES, RightIC, RightILUT, AMG, DampedJacobiSmoother,  # noqa: E402
                                   GaussSeidelSmoother)
from pysolvers_b200.Nonlinear import NewtonSolver  # noqa: E402
from pysolvers_b200.device import DeviceCSR, to_device, ptr, current_stream_ptr  # noqa: E402
from pysolvers_b200.problems import fd_laplacian_2d, load_dh_matrix, FDBratu2D  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def run(solver, A, b):
    hist = []
    solver.reportIter = lambda k, nr, nb: hist.append(nr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = quiet(solver.solve, A, b)
    torch.cuda.synchronize()
    return st, np.asarray(hist), time.perf_counter() - t0


def time_gpu(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def c1():
    A = -fd_laplacian_2d(0.0, 1.0, 256)
    b = np.ones(A.shape[0])
    s = PCG(CommonSolverArgs(maxiter=5000, tau=1e-8)).makeSolver()
    run(s, A, b)
    st, hist, dt = run(s, A, b)
    t0 = time.perf_counter()
    ref = krylov.pcg(A, b, maxiter=5000, tau=1e-8)
    cpu = time.perf_counter() - t0
    k = min(len(hist), len(ref['hist']))
    return dict(config='C1: PCG, -FDLaplacian2D(0,1,256), b=1, tau=1e-8', iters=st.iters(), ref_iters=ref['iters'],
                gpu_solve_s=dt, gpu_it_per_s=st.iters() / dt, cpu_s=cpu, cpu_it_per_s=ref['iters'] / cpu,
                hist_max_rel_err=float(np.max(np.abs(hist[:k] - ref['hist'][:k]) / ref['hist'][:k])),
                soln_rel_err=float(np.linalg.norm(st.soln() - ref['soln']) / np.linalg.norm(ref['soln'])))


def c2():
    out = []
    for lev in (12, 15):
        A = load_dh_matrix(lev)
        b = A @ np.random.default_rng(2024).random(A.shape[0])
        t0 = time.perf_counter()
        pre = quiet(RightILUT().form, A)
        setup = time.perf_counter() - t0
        dev = pre.device_prec()
        v = to_device(b)
        z = torch.empty_like(v)
        apply_s = time_gpu(lambda: dev.apply(v, z), reps=20)
        lvL, lvU = dev.levels()
        s = GMRES(CommonSolverArgs(maxiter=30, tau=1e-8), precond=RightILUT()).makeSolver()
        run(s, A, b)
        st, hist, dt = run(s, A, b)
        ilu = precond.ilut_factor(A)
        t0 = time.perf_counter()
        ref = krylov.gmres(A, b, prec=lambda x: precond.ilut_apply(ilu, x), maxiter=30, tau=1e-8)
        cpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(20):
            precond.ilut_apply(ilu, b)
        cpu_apply = (time.perf_counter() - t0) / 20
        k = min(len(hist), len(ref['hist']))
        out.append(dict(config='C2: GMRES(30)+RightILUT, DH-Matrix-%d (n=%d)' % (lev, A.shape[0]),
                        iters=st.iters(), ref_iters=ref['iters'], ilut_setup_cpu_s=setup,
                        gpu_solve_incl_setup_s=dt, gpu_ilut_apply_s=apply_s, cpu_ilut_apply_s=cpu_apply,
                        levels_L11=lvL, levels_U11=lvU, dense_tail_rows=dev.n2,
                        us_per_level=1e6 * apply_s / max(lvL + lvU, 1),
                        cpu_solve_excl_setup_s=cpu,
                        hist_max_rel_err=float(np.max(np.abs(hist[:k] - ref['hist'][:k]) / ref['hist'][:k]))))
    return out


def ic(m):
    A = -fd_laplacian_2d(0.0, 1.0, m)
    n = A.shape[0]
    b = np.ones(n)
    t0 = time.perf_counter()
    pre = quiet(RightIC().form, A)
    setup = time.perf_counter() - t0
    dev = pre.device_prec()
    v = to_device(b)
    z = torch.empty_like(v)
    apply_s = time_gpu(lambda: dev.apply(v, z), reps=5, warm=1)
    iL, iLt = pre._dL.info(), pre._dLt.info()
    t0 = time.perf_counter()
    ref_z = precond.ic_apply(pre._L, pre._Lt, b)
    cpu_apply = time.perf_counter() - t0
    zz = z.cpu().numpy()
    # solve with the already-formed preconditioner
    s = PCG(CommonSolverArgs(maxiter=2000, tau=1e-8), precond=RightIC()).makeSolver()
    s.precond = pre
    s.freezePrec()
    st, hist, dt = run(s, A, b)
    bytes_ic = 12 * (iL['nnz_packed'] + iLt['nnz_packed']) + 2 * 24 * n
    return dict(config='IC-PCG, -FDLaplacian2D(0,1,%d) (n=%d), RightIC() defaults' % (m, n),
                ic_setup_cpu_s=setup, nnz_L=iL['nnz_off'] + n, levels_L=iL['levels'], levels_Lt=iLt['levels'],
                gpu_ic_apply_s=apply_s, cpu_ic_apply_s=cpu_apply, apply_speedup=cpu_apply / apply_s,
                us_per_level=1e6 * apply_s / (iL['levels'] + iLt['levels']),
                ic_apply_algorithmic_GBps=bytes_ic / apply_s / 1e9,
                apply_rel_err_vs_scipy=float(np.linalg.norm(zz - ref_z) / np.linalg.norm(ref_z)),
                pcg_iters=st.iters(), pcg_success=st.success(), gpu_pcg_solve_s=dt,
                est_cpu_pcg_solve_s=st.iters() * cpu_apply)


def c5(m, gmres_too=False):
    """Newton + Krylov + AMG on FDBratu2D(m): timings of the solve-phase pieces and one Newton
    run with GMRES + AMG (damped Jacobi).  NB: the reference's V-cycle preconditioner starts
    every application from x0 = b (VCycleSolver.py:69); as the oracle shows that
    makes PCG stagnate for m >= 128, so the linear solves are capped."""
    out = {'config': 'C5: FDBratu2D(m=%d) Newton + Krylov + AMG(5 V-cycles, 2 levels)' % m}
    func = FDBratu2D(m=m)
    n = m * m
    J = func.evalJ(func.initialU())
    F = func.evalF(func.initialU())
    for name, sm in (('djac', DampedJacobiSmoother), ('gs', GaussSeidelSmoother)):
        t0 = time.perf_counter()
        pre = quiet(AMG(numIters=5, smoother=sm).form, J)
        out['amg_setup_host_s_' + name] = time.perf_counter() - t0
        amg = pre.device_amg()
        v = to_device(-F)
        z = torch.empty_like(v)
        out['amg_apply_5_vcycles_s_' + name] = time_gpu(lambda: amg.prec.apply(v, z), reps=3, warm=1)
        cv = to_device(np.ones(amg.A[0].shape[0]))
        cz = torch.empty_like(cv)
        out['coarse_solve_s'] = time_gpu(lambda: amg.coarse.apply(cv, cz), reps=3, warm=1)
        out['coarse_levels_L11_U11'] = list(amg.coarse.levels())
        out['coarse_dense_tail_rows'] = amg.coarse.n2
        out['levels'] = [a.shape[0] for a in amg.A]
        if name == 'djac':
            dA, dinv = amg.A[-1], amg.dinv[-1]
            f = to_device(np.ones(n))
            xa, xb = torch.zeros_like(f), torch.empty_like(f)
            lib = nat.lib()
            sweep_s = time_gpu(lambda: nat.check(lib.psb_jacobi_sweep(dA.handle, ptr(dinv), 2.0 / 3.0, ptr(f), ptr(xa),
                                                                     ptr(xb), current_stream_ptr())), reps=20)
            out['jacobi_sweep_s'] = sweep_s
            out['jacobi_sweep_GBps'] = (12 * dA.nnz + 4 * (n + 1) + 32 * n) / sweep_s / 1e9
        del pre, amg
    if gmres_too:
        lin = []
        newton = NewtonSolver(control=CommonSolverArgs(tau=1.0e-12, maxiter=6),
                              solver=GMRES(control=CommonSolverArgs(maxiter=40),
                                           precond=AMG(numIters=5, smoother=DampedJacobiSmoother)),
                              fixLinTol=False, minLinTol=1.0e-6, freezePrec=True)
        inner = newton.solver
        orig = inner.solve

        def spy2(J_, rhs, _o=orig):
            r = _o(J_, rhs)
            lin.append(r.iters())
            return r
        inner.solve = spy2
        st, hist, dt = run(newton, func, func.initialU())
        out['newton_gmres_amg_djac'] = dict(newton_iters=st.iters(), success=st.success(), lin_iters=lin,
                                            total_s=dt, F_history=[float(h) for h in hist])
    return out


def newton_c5(m, lin_maxiter=None, device=True, reuse=False):
    """configs[4] as named: FDBratu2D(m, alpha = 0.5), u0 = 1, Newton (tau = 1e-12, inexact linear
    tolerance max(0.1 ||F||/||F0||, 1e-6)) with non-restarted GMRES + AMG(5 V-cycles, 2 levels,
    damped Jacobi) -- run to convergence.  GMRES rebuilds the hierarchy on every Newton step (the
    reference's behaviour), so the host setup is inside the time."""
    from pysolvers_b200.problems import DeviceFDBratu2D
    lin_maxiter = lin_maxiter or max(60, int(0.3 * m))
    func = DeviceFDBratu2D(m=m) if device else FDBratu2D(m=m)
    lin, lin_s = [], []
    newton = NewtonSolver(control=CommonSolverArgs(tau=1.0e-12, maxiter=12),
                          solver=GMRES(control=CommonSolverArgs(maxiter=lin_maxiter), honorFreeze=reuse,
                                       precond=AMG(numIters=5, smoother=DampedJacobiSmoother)),
                          fixLinTol=False, minLinTol=1.0e-6, freezePrec=True)
    inner = newton.solver
    orig = inner.solve

    def spy(J_, rhs, _o=orig):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = _o(J_, rhs)
        torch.cuda.synchronize()
        lin.append(int(r.iters()))
        lin_s.append(time.perf_counter() - t0)
        return r
    inner.solve = spy
    st, hist, dt = run(newton, func, func.initialU())
    return dict(config='C5: FDBratu2D(m=%d, alpha=0.5) Newton + GMRES(maxiter=%d, no restart) + AMG(5 V-cycles, damped Jacobi), %s, %s'
                       % (m, lin_maxiter, 'u / F / J resident in HBM' if device else 'host operands',
                          'hierarchy of the first Jacobian reused (honorFreeze)' if reuse else
                          'hierarchy rebuilt on every Newton step (reference behaviour)'),
                newton_iters=int(st.iters()), success=bool(st.success()), lin_iters=lin,
                lin_solve_s_incl_amg_setup=[round(x, 3) for x in lin_s], total_s=dt,
                F_history=[float(h) for h in hist])


def gmres_large(m, maxiter=30):
    """The GMRES kernels at the headline size: un-preconditioned GMRES(maxiter) on the m x m
    Laplacian, matrix and right-hand side resident in HBM, both orthogonalisation orders.
    Algorithmic bytes of iteration k (k = 0 .. maxiter-1, k+1 basis vectors):
      SpMV 12 nnz + 4(n+1) + 16 n;  CGS2: two passes of (k+1) dots + (k+1) axpys over the basis,
      i.e. 2 * [2 (k+1) + 3] * 8n;  MGS (axpy of step j-1 fused with the dot of step j: read
      q_{j-1}, q_j, w, write w): (k+1) * 32 n;  norm + scale: 24 n."""
    from pysolvers_b200.problems import device_fd_laplacian
    dA = device_fd_laplacian(2, 0.0, 1.0, m, negate=True)
    n, nnz = dA.shape[0], dA.nnz
    b = torch.ones(n, dtype=torch.float64, device='cuda')
    out = {'config': 'GMRES(%d), no preconditioner, 2-D 5-point Laplacian m=%d (n=%d), device-resident' % (maxiter, m, n)}
    spmv = 12 * nnz + 4 * (n + 1) + 16 * n
    for orth in ('cgs2', 'mgs'):
        s = GMRES(CommonSolverArgs(maxiter=maxiter, tau=1e-300, failOnMaxiter=False, showIters=False, showFinal=False),
                  orth=orth).makeSolver()
        quiet(s.solve, dA, b)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = quiet(s.solve, dA, b)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        iters = maxiter
        if orth == 'cgs2':
            byts = sum(spmv + 2 * (2 * (k + 1) + 3) * 8 * n + 24 * n for k in range(iters))
        else:
            byts = sum(spmv + (k + 1) * 32 * n + 24 * n for k in range(iters))
        out[orth] = {'solve_s': dt, 'iters_reported': int(st.iters()), 'ms_per_iteration': 1e3 * dt / iters,
                     'algorithmic_GB': byts / 1e9, 'achieved_GBps': byts / dt / 1e9,
                     'final_resid': float(st.resid()) if st.resid() is not None else None}
    return out


def main():
    which = sys.argv[1:] or ['c1', 'c2', 'ic512', 'c5_512']
    res = {}
    for w in which:
        t0 = time.perf_counter()
        if w == 'c1':
            res[w] = c1()
        elif w == 'c2':
            res[w] = c2()
        elif w.startswith('ic'):
            res[w] = ic(int(w[2:]))
        elif w.startswith('newtonreuse'):
            res[w] = newton_c5(int(w[11:]), reuse=True)
        elif w.startswith('newton'):
            res[w] = newton_c5(int(w[6:]))
        elif w.startswith('gmres'):
            res[w] = gmres_large(int(w[5:]))
        elif w.startswith('c5_'):
            res[w] = c5(int(w[3:]), gmres_too=(int(w[3:]) <= 512))
        print('%s done in %.1f s' % (w, time.perf_counter() - t0), file=sys.stderr, flush=True)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
