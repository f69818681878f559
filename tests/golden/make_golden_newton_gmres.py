"""Golden fixtures for configs[4] as BASELINE.json names it -- Newton with inexact GMRES and the
AMG V-cycle preconditioner on FDBratu2D -- produced by RUNNING THE REFERENCE ITSELF (build
container only; same shims and workarounds as make_golden.py):

    python tests/golden/make_golden_newton_gmres.py   ->  tests/golden/newton_gmres_golden.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets up the reference import path and the shims)
from make_golden import (CommonSolverArgs, GMRES, AMG, NewtonSolver, FDBratu2D, DampedJacobi,  # noqa: E402
                         quiet, run)


def main():
    G = {}
    for m in (24, 40):
        func = quiet(FDBratu2D, m=m)
        lin_iters = []
        newton = NewtonSolver(control=CommonSolverArgs(tau=1.0e-12, maxiter=10),
                              solver=GMRES(control=CommonSolverArgs(maxiter=60),
                                           precond=AMG(numIters=5, smoother=DampedJacobi)),
                              fixLinTol=False, minLinTol=1.0e-6, freezePrec=True)
        inner = newton.solver
        inner.precond = None                      # GMRESSolver.py:71 reads an attribute never set
        orig = inner.solve

        def spy(J, rhs, _orig=orig, _acc=lin_iters):
            r = _orig(J, rhs)
            _acc.append(r.iters())
            return r
        inner.solve = spy
        st, h = run(newton, func, func.initialU())
        key = 'newton_gmres/bratu_m%d_djac' % m
        G[key + '/hist'] = h
        G[key + '/iters'] = np.int64(st.iters())
        G[key + '/success'] = np.bool_(st.success())
        G[key + '/x'] = np.asarray(st.soln(), dtype=np.float64)
        G[key + '/lin_iters'] = np.asarray(lin_iters, dtype=np.int64)
        print('newton+gmres+amg m=%d: newton iters=%d success=%s lin=%s |F|=%s'
              % (m, st.iters(), st.success(), lin_iters, ['%.2e' % v for v in h]))
    out = os.path.join(HERE, 'newton_gmres_golden.npz')
    np.savez_compressed(out, **G)
    print('wrote', out)


if __name__ == '__main__':
    main()
