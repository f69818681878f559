import sys, os, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pysolvers_b200.Linear import AMG, GaussSeidelSmoother, DampedJacobiSmoother, amg_setup
from pysolvers_b200.Linear.multigrid import DeviceAMG, SmoothedAggregationMLHierarchy
from pysolvers_b200.device import DeviceCSR, to_device
from pysolvers_b200.problems import fd_laplacian_2d
from oracle import multigrid as omg
A = -fd_laplacian_2d(0.0, 1.0, 32)
v = np.random.default_rng(11).random(A.shape[0])
for nlev in (2, 3):
    with contextlib.redirect_stdout(io.StringIO()):
        mlh = SmoothedAggregationMLHierarchy(A, numLevels=nlev)
    ops = [mlh.matrix(k) for k in range(nlev)]
    ups = [mlh.update(k) for k in range(nlev - 1)]
    downs = [mlh.downdate(k) for k in range(nlev - 1)]
    print('levels', [o.shape for o in ops], [DeviceCSR(o).info()['kind'] for o in ops],
          [DeviceCSR(o).info()['kind'] for o in ups], [DeviceCSR(o).info()['kind'] for o in downs])
    for M in ops + ups + downs:
        x = np.random.default_rng(1).random(M.shape[1])
        y = DeviceCSR(M).matvec(to_device(x)).cpu().numpy()
        print('   spmv exact:', np.array_equal(y, M @ x), M.shape, M.has_sorted_indices)
    for sm, osm, name in ((DampedJacobiSmoother, lambda A_: omg.Jacobi(A_, 2.0 / 3.0), 'djac'), (GaussSeidelSmoother, omg.GaussSeidel, 'gs')):
        for inp, iname in ((np.ones(A.shape[0]), 'ones'), (v, 'rand')):
            for ni in (1, 2, 5):
                dev = DeviceAMG(mlh, sm, 2, 2, ni)
                x, res, hist = dev.solve(inp, ni, 1e-8)
                ref = omg.vcycle_solve(ops, ups, downs, inp, maxiter=ni, tau=1e-8, fail_on_maxiter=False, smoother=osm)
                print(nlev, name, iname, ni, 'x rel', np.linalg.norm(x - ref['soln']) / np.linalg.norm(ref['soln']),
                      'hist', hist[-1], ref['hist'][-1])
