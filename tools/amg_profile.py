"""AMG preconditioner application (configs[4]: Bratu Jacobian, 5 V-cycles, 2 levels, damped
Jacobi) timed on the device, with a profiler range around ONE application:

    python tools/amg_profile.py 2048                       # timings (CUDA events)
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/amg_launches.csv python tools/amg_profile.py 2048 --profile
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

from pysolvers_b200 import _native as nat  # noqa: E402
from pysolvers_b200.Linear import AMG, DampedJacobiSmoother  # noqa: E402
from pysolvers_b200.device import to_device  # noqa: E402
from pysolvers_b200.problems import FDBratu2D  # noqa: E402


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


def pack_sweep(m):
    """coarse LU solve with different row-class thresholds of the wide-level packing (A/B)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from pysolvers_b200.Linear import amg_setup
    from pysolvers_b200.device import DeviceSplitLU
    func = FDBratu2D(m=m)
    J = sp.csr_matrix(func.evalJ(func.initialU()))
    ops, _, _ = amg_setup.build_hierarchy(J, 2)
    lu = spla.splu(sp.csc_matrix(ops[0]), permc_spec='MMD_AT_PLUS_A')
    cv = to_device(np.ones(ops[0].shape[0]))
    cz = torch.empty_like(cv)
    out = {}
    for short, long_ in ((8, 64), (4, 32), (8, 32), (4, 64), (8, 128), (16, 64), (2, 32)):
        os.environ['PSB_TRSV_SHORT'], os.environ['PSB_TRSV_LONG'] = str(short), str(long_)
        c = DeviceSplitLU(lu)
        out['short%d_long%d' % (short, long_)] = round(1e3 * time_gpu(lambda: c.apply(cv, cz)), 4)
        del c
    print(json.dumps(out))
    return 0


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    profile = '--profile' in sys.argv
    if '--pack-sweep' in sys.argv:
        return pack_sweep(m)
    func = FDBratu2D(m=m)
    J = func.evalJ(func.initialU())
    F = func.evalF(func.initialU())
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        pre = AMG(numIters=5, smoother=DampedJacobiSmoother).form(J)
    setup = time.perf_counter() - t0
    amg = pre.device_amg()
    v = to_device(-F)
    z = torch.empty_like(v)
    if profile:
        amg.prec.apply(v, z)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        amg.prec.apply(v, z)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0
    if '--ops' in sys.argv:
        # every operator of the hierarchy under each SpMV kind it supports (plain y = M x)
        for name, M in (('A_fine', amg.A[-1]), ('R', amg.R[-1]), ('P', amg.P[-1])):
            xin = torch.rand(M.shape[1], dtype=torch.float64, device='cuda')
            yout = torch.empty(M.shape[0], dtype=torch.float64, device='cuda')
            info = M.info()
            r = dict(op=name, shape=list(M.shape), nnz=M.nnz, auto_kind=info['kind'], max_row=info['max_row'],
                     MB=round((12 * M.nnz + 4 * M.shape[0] + 8 * M.shape[0] + 8 * M.shape[1]) / 1e6, 1))
            for kind, kn in ((1, 'bulk'), (3, 'lsu'), (3 | 16, 'lsu512'), (2, 'vector'), (4, 'merge')):
                try:
                    M.set_kind(kind)
                    r[kn + '_us'] = round(1e6 * time_gpu(lambda: M.matvec(xin, yout), reps=20), 1)
                except Exception as e:
                    r[kn + '_us'] = None
            M.set_kind(info['kind'])
            print(json.dumps(r), flush=True)
    l0 = nat.launch_count()
    amg.prec.apply(v, z)
    launches = nat.launch_count() - l0
    out = {'m': m, 'levels': [a.shape[0] for a in amg.A], 'amg_setup_host_s': setup,
           'apply_5_vcycles_ms': 1e3 * time_gpu(lambda: amg.prec.apply(v, z)), 'launches_per_apply': int(launches)}
    cv = to_device(np.ones(amg.A[0].shape[0]))
    cz = torch.empty_like(cv)
    out['coarse_solve_ms'] = 1e3 * time_gpu(lambda: amg.coarse.apply(cv, cz))
    for la in os.environ.get('PSB_SWEEP_LOOKAHEAD', '').split():
        os.environ['PSB_TRSV_LOOKAHEAD'] = la
        out['coarse_solve_ms_lookahead_' + la] = 1e3 * time_gpu(lambda: amg.coarse.apply(cv, cz))
    os.environ.pop('PSB_TRSV_LOOKAHEAD', None)
    out['coarse_levels_L11_U11'] = list(amg.coarse.levels())
    out['coarse_dense_rows'] = amg.coarse.n2
    print(json.dumps(out))
    return 0


if __name__ == '__main__':
    sys.exit(main())
