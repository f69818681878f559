"""configs[2] as close to its named size as a GPU call allows: PCG + incomplete Cholesky
(RightIC defaults, ICPreconditioner.py:45-63) on the 2-D 5-point Laplacian m x m.

    python tools/ic_large.py 2048 > gpurun_out/ic2048.json

The SuperLU setup is the reference's and runs on the host (about 7 - 12 min at m = 2048, one
core).  Reported: factor size, dependency levels, the apply with the kernel the analysis picks and
with each kernel forced (us per level), IC-PCG to tau = 1e-8 (iterations against the
BASELINE.md extrapolation 30 / 55 / 107 / ~208), the apply against two host
spsolve_triangular calls, and the solution's true residual.
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

from oracle import precond  # noqa: E402
from pysolvers_b200 import CommonSolverArgs  # noqa: E402
from pysolvers_b200.Linear import PCG, RightIC  # noqa: E402
from pysolvers_b200.device import to_device  # noqa: E402
from pysolvers_b200.problems import fd_laplacian_2d  # noqa: E402


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
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    A = -fd_laplacian_2d(0.0, 1.0, m)
    n = A.shape[0]
    b = np.ones(n)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        pre = RightIC().form(A)
    setup = time.perf_counter() - t0
    print('setup done in %.1f s' % setup, file=sys.stderr, flush=True)
    dev = pre.device_prec()
    v, z = to_device(b), torch.empty(n, dtype=torch.float64, device='cuda')
    iL, iLt = pre._dL.info(), pre._dLt.info()
    levels = iL['levels'] + iLt['levels']
    out = {'config': 'IC-PCG, -FDLaplacian2D(0,1,%d) (n=%d), RightIC() defaults' % (m, n),
           'ic_setup_host_s': setup, 'nnz_L': int(pre._L.nnz), 'levels_L': iL['levels'], 'levels_Lt': iLt['levels'],
           'chunks_per_level_L': iL['groups'] / max(iL['levels'], 1),
           'kernel_chosen': pre._dL.info2()['kernel'], 'info2_L': pre._dL.info2()}
    apply_s = time_gpu(lambda: dev.apply(v, z))
    z_auto = z.clone()
    out['apply_ms'] = 1e3 * apply_s
    out['us_per_level'] = 1e6 * apply_s / levels
    out['apply_algorithmic_GBps'] = (12 * (iL['nnz_packed'] + iLt['nnz_packed']) + 48 * n) / apply_s / 1e9
    for kern in ('cta', 'cluster', 'grid'):
        try:
            pre._dL.set_kernel(kern)
            pre._dLt.set_kernel(kern)
            s = time_gpu(lambda: dev.apply(v, z), reps=3, warm=1)
            out['forced_' + kern] = {'apply_ms': 1e3 * s, 'us_per_level': 1e6 * s / levels,
                                    'bit_identical_to_chosen': bool(torch.equal(z, z_auto))}
        except Exception as exc:                    # a kernel the analysis rules out for this factor
            out['forced_' + kern] = {'error': str(exc)[:160]}
    pre._dL.set_kernel(None)
    pre._dLt.set_kernel(None)
    t0 = time.perf_counter()
    ref_z = precond.ic_apply(pre._L, pre._Lt, b)
    out['cpu_ic_apply_s'] = time.perf_counter() - t0
    dev.apply(v, z)
    out['apply_rel_err_vs_scipy'] = float(np.linalg.norm(z.cpu().numpy() - ref_z) / np.linalg.norm(ref_z))
    s = PCG(CommonSolverArgs(maxiter=2000, tau=1e-8, showIters=False, showFinal=False), precond=RightIC()).makeSolver()
    s.precond = pre
    s.freezePrec()
    hist = []
    s.reportIter = lambda k, nr, nb: hist.append(nr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        st = s.solve(A, b)
    torch.cuda.synchronize()
    out['pcg_solve_s'] = time.perf_counter() - t0
    out['pcg_iters'] = int(st.iters())
    out['pcg_success'] = bool(st.success())
    out['pcg_iters_extrapolated_from_baseline'] = {256: 30, 512: 55, 1024: 107, 2048: 208, 4096: 405}.get(m)
    x = st.soln()
    out['true_rel_residual'] = float(np.linalg.norm(b - A @ x) / np.linalg.norm(b))
    out['est_cpu_pcg_solve_s'] = st.iters() * out['cpu_ic_apply_s']
    print(json.dumps(out))
    return 0


if __name__ == '__main__':
    sys.exit(main())
