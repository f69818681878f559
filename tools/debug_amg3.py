import sys, os, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysolvers_b200 import CommonSolverArgs
from pysolvers_b200.Linear import AMG, AMGVCycle, GaussSeidelSmoother, DampedJacobiSmoother
from pysolvers_b200.problems import fd_laplacian_2d
from oracle import multigrid as omg
g = np.load('tests/golden/reference_golden.npz')
A = -fd_laplacian_2d(0.0, 1.0, 32)
v = g['amg/m32_L3/apply_djac_in']
go = g['amg/m32_L3/apply_djac_out']
s = AMGVCycle(CommonSolverArgs(maxiter=12, tau=1e-8, failOnMaxiter=False), numLevels=3, smoother=DampedJacobiSmoother).makeSolver()
with contextlib.redirect_stdout(io.StringIO()):
    st = s.solve(A, np.ones(A.shape[0]))
    pre = AMG(numIters=5, numLevels=3, smoother=DampedJacobiSmoother).form(A)
    out = pre.apply(v)
    x3, res, hist = pre.device_amg().solve(v, 5, 1e-8)
    x4, res4, hist4 = s._cycleMgr.device().solve(v, 5, 1e-8)
m1, m2 = s._cycleMgr.device().mlh, pre.device_amg().mlh
for k in range(3):
    a, b = m1.matrix(k), m2.matrix(k)
    print('A%d equal' % k, np.array_equal(a.data, b.data) and np.array_equal(a.indices, b.indices))
print('apply rel', np.linalg.norm(out - go) / np.linalg.norm(go), 'direct', np.linalg.norm(x3 - go) / np.linalg.norm(go), hist,
      'via first object', np.linalg.norm(x4 - go) / np.linalg.norm(go), hist4)
print('control tau of pre', pre._solver.tau(), pre._solver.maxiter(), 'tau of s', s.tau())
