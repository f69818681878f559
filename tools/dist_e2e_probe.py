"""Where an end-to-end step of the multi-GPU path spends its time (run under torchrun)."""
import os, sys, time, contextlib, io
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['PSB_DIST_TIMING'] = '1'
from pysolvers_b200 import CommonSolverArgs
from pysolvers_b200.dist import Comm, DistCSR, DistributedPCG, laplacian_block_device, row_starts

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
dist.init_process_group('nccl')
dev = torch.device('cuda', torch.cuda.current_device())
comm = Comm()
m = 4096; n = m * m
st = row_starts(n, world); lo, hi = int(st[rank]), int(st[rank + 1])
indptr, cols, data = laplacian_block_device(2, 0.0, 1.0, m, lo, hi, dev)
data = -data
pin = lambda t: t.cpu().pin_memory().numpy()
ip_h, cols_h, dt_h = pin(indptr), pin(cols), pin(data)
b_h = torch.ones(hi - lo, dtype=torch.float64).pin_memory().numpy()
solver = DistributedPCG(CommonSolverArgs(maxiter=200, tau=0.0, failOnMaxiter=False, showIters=False, showFinal=False))
for it in range(3):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    D = DistCSR(comm, ip_h, cols_h, dt_h, lo, hi, n)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        s = solver.solve(D, b_h)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    del D
    torch.cuda.synchronize(); t3 = time.perf_counter()
    if rank == 0:
        print('step %d: DistCSR %.1f ms, solve+download %.1f ms, destroy %.1f ms' % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
dist.destroy_process_group()
