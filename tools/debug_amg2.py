import sys, os, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysolvers_b200 import CommonSolverArgs
from pysolvers_b200.Linear import AMG, AMGVCycle, GaussSeidelSmoother, DampedJacobiSmoother
from pysolvers_b200.problems import fd_laplacian_2d
g = np.load('tests/golden/reference_golden.npz')
A = -fd_laplacian_2d(0.0, 1.0, 32)
v = g['amg/m32_L3/apply_djac_in']
go = g['amg/m32_L3/apply_djac_out']
for trial in range(3):
    with contextlib.redirect_stdout(io.StringIO()):
        pre = AMG(numIters=5, numLevels=3, smoother=DampedJacobiSmoother).form(A)
        out = pre.apply(v)
        out2 = pre.apply(v)
        x3, res, hist = pre.device_amg().solve(v, 5, 1e-8)
    d = np.abs(out - go)
    print(trial, 'rel', np.linalg.norm(out - go) / np.linalg.norm(go), 'max at', d.argmax(), d.max(),
          'again', np.linalg.norm(out2 - go) / np.linalg.norm(go), 'direct', np.linalg.norm(x3 - go) / np.linalg.norm(go),
          'cycles', res.n_hist, res.status, hist)
