"""Where does the end-to-end step (PCGSolver.solve with host operands) spend its time?"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pysolvers_b200 import CommonSolverArgs  # noqa: E402
from pysolvers_b200.Linear import PCG  # noqa: E402
from pysolvers_b200.device import DeviceCSR, to_device  # noqa: E402


def t(fn, reps=3):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, r


A, b = bench.build_problem(4096)
n = A.shape[0]
print('pinned?', torch.from_numpy(A.data).is_pinned(), torch.from_numpy(b).is_pinned())
dt, _ = t(lambda: to_device(A.data))
print('H2D vals %.1f MB: %.1f ms = %.1f GB/s' % (A.data.nbytes / 1e6, dt * 1e3, A.data.nbytes / dt / 1e9))
dt, _ = t(lambda: to_device(A.indices, torch.int32))
print('H2D cols %.1f MB: %.1f ms = %.1f GB/s' % (A.indices.nbytes / 1e6, dt * 1e3, A.indices.nbytes / dt / 1e9))
dt, dA = t(lambda: DeviceCSR(A))
print('DeviceCSR(A) total: %.1f ms' % (dt * 1e3))
x = torch.ones(n, dtype=torch.float64, device='cuda')
dt, _ = t(lambda: x.cpu().numpy())
print('D2H x pageable: %.1f ms = %.1f GB/s' % (dt * 1e3, 8 * n / dt / 1e9))
stage = torch.empty(n, dtype=torch.float64, pin_memory=True)
dt, _ = t(lambda: stage.copy_(x, non_blocking=True))
print('D2H x pinned: %.1f ms = %.1f GB/s' % (dt * 1e3, 8 * n / dt / 1e9))
dt, _ = t(lambda: torch.empty(n * 8 * 5, dtype=torch.uint8, device='cuda'))
print('torch.empty workspace: %.2f ms' % (dt * 1e3))
s = PCG(CommonSolverArgs(maxiter=200, tau=0.0, failOnMaxiter=False, showIters=False, showFinal=False)).makeSolver()


def solve():
    with contextlib.redirect_stdout(io.StringIO()):
        return s.solve(A, b)
dt, _ = t(solve)
print('solve(A, b) e2e: %.1f ms' % (dt * 1e3))
