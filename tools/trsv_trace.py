"""Where a chunk of the one-CTA SpTRSV kernel spends its time (clock64 stamps per chunk)."""
import os, sys
import numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysolvers_b200.device import DeviceTrsv, to_device, ptr
from pysolvers_b200 import _native as nat
from oracle import precond
from pysolvers_b200.problems import fd_laplacian_2d


def trace(T, lower, name):
    dT = DeviceTrsv(T, lower=lower)
    dT.set_kernel('cta')
    v = to_device(np.ones(T.shape[0]))
    out = torch.empty_like(v)
    dT.solve(v, out)
    g = dT.info()['groups']
    buf = torch.zeros(12 * g, dtype=torch.int64, device='cuda')
    nat.check(nat.lib().psb_trsv_set_trace(dT.handle, ptr(buf)), 'trace')
    dT.solve(v, out)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(g, 12).astype(np.float64)
    t0 = t[:, 0].min()
    for c in (0, 1, 2, 3, 4, 5, 8, 9, 10):
        t[:, c] = np.where(t[:, c] > 0, t[:, c] - t0, np.nan)
    lo = g // 4
    s = t[lo:lo + 20000]
    done_prev = np.concatenate([[np.nan], s[:-1, 5]])
    print('%s: %d chunks, %d levels; %.1f cycles/chunk' % (name, g, dT.info()['levels'], np.nanmax(t[:, 5]) / g))
    med = lambda a: np.nanmedian(a)
    print('   start->woke %.0f | woke->staged %.0f | woke->stored %.0f | stored->done %.0f | entries/lane %.1f'
          % (med(s[:, 2] - s[:, 0]), med(s[:, 3] - s[:, 2]), med(s[:, 4] - s[:, 2]), med(s[:, 5] - s[:, 4]), med(s[:, 7])))
    stored_prev = np.concatenate([[np.nan], s[:-1, 4]])
    print('   relative to stored(g-1): woke %.0f | stored %.0f   (= cycles per chunk on the critical path)'
          % (med(s[:, 2] - stored_prev), med(s[:, 4] - stored_prev)))
    print('   last arrivals relative to stored(g-1): third-last %.0f | second-last %.0f | last %.0f | stored %.0f'
          % (med(s[:, 10] - stored_prev), med(s[:, 9] - stored_prev), med(s[:, 8] - stored_prev), med(s[:, 4] - stored_prev)))
    for r in (s[100:106] - s[100, 0]):
        print('     ', np.nan_to_num(r).astype(np.int64))


n = 20000
trace(sp.diags([np.full(n - 1, -0.5), np.full(n, 1.5)], [-1, 0]).tocsr(), True, 'chain')
L, Lt = precond.ic_factor(-fd_laplacian_2d(0.0, 1.0, 256))
trace(L, True, 'IC256-L')
