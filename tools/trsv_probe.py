"""Per-level latency of the two SpTRSV kernels (grid-wide / one CTA with a shared-memory window)
on a pure chain, IC factors of the 2-D Laplacian, an ILUT factor pair and a Gauss-Seidel triangle."""
import os, sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysolvers_b200.device import DeviceTrsv, to_device
from oracle import precond
from pysolvers_b200.problems import fd_laplacian_2d, load_dh_matrix


def sweep(T, lower, name):
    """one-CTA kernel under different spin policies (PSB_TRSV_NEAR_LEVELS: levels that spin instead
    of sleeping; PSB_TRSV_SLEEP_NS: sleep per level of distance)"""
    dT = DeviceTrsv(T, lower=lower)
    v = to_device(np.ones(T.shape[0]))
    out = torch.empty_like(v)
    kern = os.environ.get('PSB_PROBE_KERNEL', 'cta')
    dT.set_kernel(kern)
    lv = dT.info()['levels']
    row = []
    nears = os.environ.get('PSB_PROBE_NEARS', '0.5 1.0 1.5 2.0 3.0 4.0').split()
    sleeps = os.environ.get('PSB_PROBE_SLEEPS', '32 64 128').split()
    for near in nears:
        for ns in sleeps:
            os.environ['PSB_TRSV_NEAR_LEVELS'], os.environ['PSB_TRSV_SLEEP_NS'] = near, ns
            dT.solve(v, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                dT.solve(v, out)
            e1.record()
            torch.cuda.synchronize()
            row.append('%s/%s: %.3f' % (near, ns, 1e3 * e0.elapsed_time(e1) / 10 / lv))
    os.environ.pop('PSB_TRSV_NEAR_LEVELS'); os.environ.pop('PSB_TRSV_SLEEP_NS')
    print('%-10s %s us/level by (levels that spin / sleep ns per level): %s' % (name, kern, '  '.join(row)), flush=True)


def bench(T, lower, name, unit=False, reps=5):
    if os.environ.get('PSB_PROBE_SWEEP'):
        return sweep(T, lower, name)
    dT = DeviceTrsv(T, lower=lower, unit_diag=unit)
    v = to_device(np.ones(T.shape[0]))
    out = torch.empty_like(v)
    res = {}
    kernels = ['grid', 'cta'] + (['cluster'] if dT.info2()['cluster_ok'] else [])
    for kern in kernels:
        dT.set_kernel(kern)
        dT.solve(v, out)
        torch.cuda.synchronize()
        dT.check()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dT.solve(v, out)
        e1.record()
        torch.cuda.synchronize()
        res[kern] = (e0.elapsed_time(e1) / reps, out.clone())
    for kern in kernels[1:]:
        assert torch.equal(res['grid'][1], res[kern][1]), (name, kern)
    i, i2 = dT.info(), dT.info2()
    lv = i['levels']
    us = lambda k: ('%7.3f' % (1e3 * res[k][0] / lv)) if k in res else '    n/a'
    print('%-14s n=%8d lev=%6d rows/lev=%7.1f chunks/lev=%5.1f far=%6d | us/level: grid %s  cta %s  cluster %s | auto=%-7s %8.3f ms'
          % (name, i['n'], lv, i['n'] / lv, i['groups'] / lv, i2['n_far'], us('grid'), us('cta'), us('cluster'),
             i2['kernel'], res[i2['kernel']][0]), flush=True)


which = sys.argv[1:] or ['chain', 'ic128', 'ic256', 'ic512', 'dh15', 'gs256', 'gs512', 'gs2048', 'lu']
for w in which:
    if w == 'chain':
        n = 20000
        bench(sp.diags([np.full(n - 1, -0.5), np.full(n, 1.5)], [-1, 0]).tocsr(), True, 'chain20000')
    elif w.startswith('ic'):
        m = int(w[2:])
        L, Lt = precond.ic_factor(-fd_laplacian_2d(0.0, 1.0, m))
        bench(L, True, 'IC%d-L' % m)
        bench(Lt, False, 'IC%d-Lt' % m)
    elif w.startswith('dh'):
        A = load_dh_matrix(int(w[2:]))
        ilu = spla.spilu(sp.csc_matrix(A), drop_tol=1e-3, fill_factor=15)
        bench(ilu.L.tocsr(), True, w + '-L', unit=True)
        bench(ilu.U.tocsr(), False, w + '-U')
    elif w.startswith('gs'):
        m = int(w[2:])
        bench(sp.triu(-fd_laplacian_2d(0.0, 1.0, m)).tocsr(), False, 'GS%d-triu' % m)
    elif w == 'lu':
        lu = spla.splu(sp.csc_matrix(-fd_laplacian_2d(0.0, 1.0, 160)))
        bench(lu.L.tocsr(), True, 'LU160-L', unit=True)
        bench(lu.U.tocsr(), False, 'LU160-U')
