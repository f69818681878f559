"""configs[4] row-partitioned: FDBratu2D(m), Newton (tau = 1e-12) + inexact GMRES + AMG V-cycle
preconditioner (5 cycles, 2 levels, damped Jacobi) with preconditioner reuse, on 1 / 2 / 4 / 8 GPUs.

    python tools/dist_bratu.py --gridm 2048                                      # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/dist_bratu.py --gridm 2048

Prints one JSON line (rank 0): Newton / GMRES iteration counts, ||F|| history, time of the whole
Newton solve, of the (replicated, host) AMG setup inside it, and of one preconditioner application
and one Jacobian product on the device.  The iteration counts must not depend on N.
"""
import argparse
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

from pysolvers_b200 import CommonSolverArgs  # noqa: E402
from pysolvers_b200.Linear import AMG, GMRES, DampedJacobiSmoother  # noqa: E402
from pysolvers_b200.Nonlinear import NewtonSolver  # noqa: E402


def time_gpu(fn, reps=5, warm=2):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gridm', dest='m', type=int, default=256)
    ap.add_argument('--lin-maxiter', type=int, default=600)
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    m = args.m
    nctl = dict(tau=1.0e-12, maxiter=10, showIters=False, showFinal=False)
    lctl = dict(maxiter=args.lin_maxiter, showIters=False, showFinal=False)
    lin, hist = [], []
    comm = None
    if world == 1:
        from pysolvers_b200.problems import DeviceFDBratu2D
        func = DeviceFDBratu2D(m=m)
        newton = NewtonSolver(control=CommonSolverArgs(**nctl),
                              solver=GMRES(CommonSolverArgs(**lctl), precond=AMG(numIters=5, smoother=DampedJacobiSmoother),
                                           honorFreeze=True),
                              fixLinTol=False, minLinTol=1.0e-6, freezePrec=True)
    else:
        import torch.distributed as dist
        from pysolvers_b200 import dist as pdist
        from pysolvers_b200 import dist_krylov as dk
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        comm = pdist.Comm()
        func = dk.DistFDBratu2D(comm, m=m)
        newton = NewtonSolver(control=CommonSolverArgs(norm=dk.dist_norm, **nctl),
                              solver=dk.DistributedGMRES(CommonSolverArgs(**lctl), precond=dk.DistAMG(comm, numIters=5)),
                              fixLinTol=False, minLinTol=1.0e-6, freezePrec=True)
    inner = newton.solver
    orig = inner.solve
    t_lin = [0.0]

    def spy(J, rhs):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = orig(J, rhs)
        torch.cuda.synchronize()
        t_lin.append(time.perf_counter() - t0)
        lin.append(r.iters())
        return r
    inner.solve = spy
    newton.reportIter = lambda k, nr, nb: hist.append(float(nr))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        st = newton.solve(func, func.initialU())
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    pre = inner.precond
    out = {'config': 'FDBratu2D(m=%d) Newton + GMRES + AMG(5 V-cycles, 2 levels, damped Jacobi), reuse' % m,
           'n_gpus': world, 'newton_iters': int(st.iters()), 'success': bool(st.success()), 'gmres_iters': lin,
           'F_history': hist, 'total_s': total, 'linear_solves_s': t_lin[1:],
           'first_linear_solve_minus_second_s (~ host AMG setup)': (t_lin[1] - t_lin[2]) if len(t_lin) > 2 else None}
    # one preconditioner application / one Jacobian product, device-timed
    n_loc = func.initialU().numel()
    v = torch.ones(n_loc, dtype=torch.float64, device='cuda')
    z = torch.empty_like(v)
    if world == 1:
        dev = pre.device_amg()
        out['amg_apply_ms'] = 1e3 * time_gpu(lambda: dev.prec.apply(v, z))
    else:
        out['amg_apply_ms'] = 1e3 * time_gpu(lambda: pre.apply(v, z))
    if rank == 0:
        print(json.dumps(out), flush=True)
    if comm is not None:
        import torch.distributed as dist
        dist.barrier()
        pre.close()
        func.close()
        comm.close()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
