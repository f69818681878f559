"""Per-level timeline of the grid SpTRSV kernel on the sparse leading blocks of the AMG coarse LU
factors (Bratu m^2, 2 levels): when the last chunk of every dependency level finished
(%globaltimer stamps per chunk, psb_trsv_set_trace), next to the level's rows / entries / longest row.

    python tools/trsv_levels.py 2048 [--tails 8192,16384] [--spmv] [--no-table]
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysolvers_b200 import _native as nat  # noqa: E402
from pysolvers_b200.Linear import amg_setup  # noqa: E402
from pysolvers_b200.device import DeviceSplitLU, ptr, to_device  # noqa: E402
from pysolvers_b200.problems import FDBratu2D  # noqa: E402


def time_gpu(fn, reps=10, warm=3):
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


def level_table(dT, name, apply):
    info = dT.info()
    g, n, nl = info['groups'], info['n'], info['levels']
    lp = np.zeros(nl + 1, dtype=np.int32)
    lr = np.zeros(n, dtype=np.int32)
    nat.check(nat.lib().psb_trsv_get_levels(dT.handle, lp.ctypes.data_as(C.c_void_p), lr.ctypes.data_as(C.c_void_p)), 'levels')
    buf = torch.zeros(3 * g, dtype=torch.int64, device='cuda')
    nat.check(nat.lib().psb_trsv_set_trace(dT.handle, ptr(buf)), 'trace')
    apply()
    apply()
    torch.cuda.synchronize()
    nat.check(nat.lib().psb_trsv_set_trace(dT.handle, None), 'trace')
    t = buf.cpu().numpy().reshape(g, 3)
    lev = np.searchsorted(lp, t[:, 2], side='right') - 1
    t0 = t[:, 0].min()
    claim, done = (t[:, 0] - t0) * 1e-3, (t[:, 1] - t0) * 1e-3
    print('%s: n %d, %d levels, %d chunks, kernel %.1f us' % (name, n, nl, g, done.max()))
    print('  level   rows chunks | first claim  last done  (+ us) | median chunk us (claim->done)')
    prev = 0.0
    rows = []
    for l in range(nl):
        s = lev == l
        if not s.any():
            continue
        d = done[s].max()
        rows.append((l, int(lp[l + 1] - lp[l]), int(s.sum()), float(claim[s].min()), float(d), float(d - prev),
                     float(np.median(done[s] - claim[s]))))
        prev = max(prev, d)
    for r in rows:
        print('  %5d %6d %6d | %10.1f %10.1f %7.2f | %8.2f' % r)
    return rows


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    tails = [8192]
    if '--tails' in sys.argv:
        tails = [int(v) for v in sys.argv[sys.argv.index('--tails') + 1].split(',')]
    func = FDBratu2D(m=m)
    J = sp.csr_matrix(func.evalJ(func.initialU()))
    ops, _, _ = amg_setup.build_hierarchy(J, 2)
    lu = spla.splu(sp.csc_matrix(ops[0]), permc_spec='MMD_AT_PLUS_A')
    cv = to_device(np.ones(ops[0].shape[0]))
    cz = torch.empty_like(cv)
    ref = None
    out = {}
    smins = [None]
    if '--smin' in sys.argv:                     # smallest supernode that is collapsed (device.COLLAPSE_MIN_ROWS)
        smins = [int(v) for v in sys.argv[sys.argv.index('--smin') + 1].split(',')]
    import pysolvers_b200.device as dev
    for tail, smin in [(t, sm) for t in tails for sm in smins]:
        if smin is not None:
            dev.COLLAPSE_MIN_ROWS = smin
        c = DeviceSplitLU(lu, tail=tail)
        ms = 1e3 * time_gpu(lambda: c.apply(cv, cz))
        z = cz.cpu().numpy()
        if ref is None:
            ref = lu.solve(np.ones(ops[0].shape[0]))
        err = float(np.abs(z - ref).max() / np.abs(ref).max())
        out['tail%d_smin%s' % (tail, smin)] = dict(coarse_solve_ms=round(ms, 4), levels=list(c.levels()), dense_rows=[c.n2, c.n2U],
                                    rel_err_vs_superlu=err, cond=[float(v) for v in c.cond_dense])
        print(json.dumps({('tail%d_smin%s' % (tail, smin)): out['tail%d_smin%s' % (tail, smin)]}), flush=True)
        if '--profile' in sys.argv:              # one coarse solve inside a profiler range (ncu --profile-from-start off)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            c.apply(cv, cz)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        if '--backoff' in sys.argv:                  # back-off of waiting warps in the grid kernel
            b1 = torch.ones(c.n1, dtype=torch.float64, device='cuda')
            o1 = torch.empty_like(b1)
            b2 = torch.ones(c.n1U, dtype=torch.float64, device='cuda')
            o2 = torch.empty_like(b2)
            for idle in (2, 5, 10, 20, 40):
                for cap in (128, 256, 512, 1024, 2048, 4096):
                    os.environ['PSB_TRSV_IDLE_TRIPS'], os.environ['PSB_TRSV_SLEEP_CAP'] = str(idle), str(cap)
                    r = dict(idle_trips=idle, sleep_cap=cap,
                             coarse_solve_ms=round(1e3 * time_gpu(lambda: c.apply(cv, cz)), 4),
                             L11_us=round(1e6 * time_gpu(lambda: c.L11.solve(b1, o1)), 1),
                             U11_us=round(1e6 * time_gpu(lambda: c.U11.solve(b2, o2)), 1))
                    print(json.dumps(r), flush=True)
            os.environ.pop('PSB_TRSV_IDLE_TRIPS'); os.environ.pop('PSB_TRSV_SLEEP_CAP')
        if '--spmv' in sys.argv:
            b = torch.ones(c.n1, dtype=torch.float64, device='cuda')
            xo = torch.empty_like(b)
            bu = torch.ones(c.n1U, dtype=torch.float64, device='cuda')
            xu = torch.empty_like(bu)
            # the two off-diagonal SpMVs of the split: kernel kinds
            y2 = torch.empty(c.n2, dtype=torch.float64, device='cuda')
            x2 = torch.ones(c.n2U, dtype=torch.float64, device='cuda')
            for name, M, xin, yout in (('L21', c.L21, b, y2), ('U12', c.U12, x2, xu)):
                r = dict(spmv=name, shape=list(M.shape), nnz=int(M.nnz), auto_kind=M.info()['kind'])
                for kind, kn in ((2, 'vector'), (4, 'merge')):
                    try:
                        M.set_kind(kind)
                        r[kn + '_us'] = round(1e6 * time_gpu(lambda: M.matvec(xin, yout)), 1)
                    except Exception as e:      # a kind the matrix does not support
                        r[kn + '_us'] = str(e)[:60]
                M.set_kind(r['auto_kind'])
                print(json.dumps(r), flush=True)
            auto = c.U12.info()['kind']
            for kind in (2, 4):
                c.U12.set_kind(kind)
                print(json.dumps({'U12_kind': kind, 'coarse_solve_ms': round(1e3 * time_gpu(lambda: c.apply(cv, cz)), 4)}), flush=True)
            c.U12.set_kind(auto)
        if tail == tails[0] and smin == smins[-1] and '--no-table' not in sys.argv:
            level_table(c.L11, 'L11', lambda: c.apply(cv, cz))
            level_table(c.U11, 'U11', lambda: c.apply(cv, cz))
        del c
    print(json.dumps(out))
    return 0


if __name__ == '__main__':
    sys.exit(main())
