"""Per-level latency of the SpTRSV kernel: a pure chain (one row per level) and an IC factor."""
import os, sys, time
import numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysolvers_b200.device import DeviceTrsv, to_device
from oracle import precond
from pysolvers_b200.problems import fd_laplacian_2d


def bench(T, lower, name, reps=5):
    dT = DeviceTrsv(T, lower=lower)
    v = to_device(np.ones(T.shape[0]))
    out = torch.empty_like(v)
    dT.solve(v, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dT.solve(v, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    lv = dT.info()['levels']
    print('%s tune=%s: %.2f ms, %d levels, %.2f us/level' % (name, os.environ.get('PSB_TRSV_TUNE', '0'), ms, lv, 1e3 * ms / lv))


n = 20000
chain = sp.diags([np.full(n - 1, -0.5), np.full(n, 1.5)], [-1, 0]).tocsr()
bench(chain, True, 'chain20000')
L, Lt = precond.ic_factor(-fd_laplacian_2d(0.0, 1.0, 256))
bench(L, True, 'IC256-L')
