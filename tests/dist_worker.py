"""Run under torchrun (one rank per GPU): checks the row-partitioned SpMV and
PCG against scipy / the oracle.  Used by tests/test_gpu_dist.py and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29517 tests/dist_worker.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import krylov  # noqa: E402
from pysolvers_b200 import CommonSolverArgs  # noqa: E402
from pysolvers_b200 import dist as pdist  # noqa: E402
from pysolvers_b200.problems import fd_laplacian_2d, fd_laplacian_3d, load_dh_matrix  # noqa: E402


def main():
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    comm = pdist.Comm()
    fails = []
    cases = [('lap2d', -fd_laplacian_2d(0.0, 1.0, 96), 1e-8, 600),
             ('lap3d', fd_laplacian_3d(0.0, 1.0, 24), 1e-9, 400),
             ('dh12', load_dh_matrix(12), 1e-8, 60)]
    for name, A, tau, maxiter in cases:
        A = sp.csr_matrix(A)
        n = A.shape[0]
        starts = pdist.row_starts(n, world)
        lo, hi = int(starts[rank]), int(starts[rank + 1])
        blk = A[lo:hi, :]
        D = pdist.DistCSR(comm, blk.indptr, blk.indices, blk.data, lo, hi, n)
        x = np.random.default_rng(3).standard_normal(n)
        y = D.matvec(torch.from_numpy(x[lo:hi]).cuda()).cpu().numpy()
        if not np.array_equal(y, (A @ x)[lo:hi]):
            fails.append('%s: dist spmv differs from scipy' % name)
        b = np.ones(n) if name != 'dh12' else A @ np.random.default_rng(1).random(n)
        solver = pdist.DistributedPCG(CommonSolverArgs(maxiter=maxiter, tau=tau))
        hist = []
        solver.reportIter = lambda k, nr, nb: hist.append(nr)
        with contextlib.redirect_stdout(io.StringIO()):
            st = solver.solve(D, b[lo:hi])
        ref = krylov.pcg(A, b, maxiter=maxiter, tau=tau)
        with contextlib.redirect_stdout(io.StringIO()):
            st_again = solver.solve(D, b[lo:hi])       # epochs / ping-pong buffers carry over
        if st_again.iters() != st.iters() or not np.array_equal(st_again.soln(), st.soln()):
            fails.append('%s: second solve differs from the first' % name)
        hist = np.asarray(hist)
        k = min(len(hist), len(ref['hist']))
        if name == 'dh12':
            # ill-conditioned + un-preconditioned: compare while the history is meaningful
            k = min(k, 20)
        rel = np.max(np.abs(hist[:k] - ref['hist'][:k]) / ref['hist'][:k])
        if st.success() != ref['success'] and name != 'dh12':
            fails.append('%s: success %s vs %s' % (name, st.success(), ref['success']))
        if name != 'dh12' and abs(st.iters() - ref['iters']) > 1:
            fails.append('%s: iters %d vs %d' % (name, st.iters(), ref['iters']))
        if rel > 1e-10:
            fails.append('%s: history rel err %.3e' % (name, rel))
        if name != 'dh12':
            err = np.linalg.norm(st.soln() - ref['soln'][lo:hi]) / np.linalg.norm(ref['soln'][lo:hi])
            if err > 1e-8:
                fails.append('%s: solution rel err %.3e' % (name, err))
        if rank == 0:
            print('%s: n=%d world=%d iters=%d (oracle %d) hist rel %.2e halo=%d r0=%d r1=%d p2p=%s'
                  % (name, n, world, st.iters(), ref['iters'], rel, D.n_halo, D.r0, D.r1, D.p2p), flush=True)
        del D
    flag = torch.tensor([len(fails)], device='cuda')
    dist.all_reduce(flag)
    if fails:
        print('rank %d FAILURES: %s' % (rank, fails), flush=True)
    comm.close()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == '__main__':
    main()
