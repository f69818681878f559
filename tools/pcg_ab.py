"""A/B of the PCG drivers (persistent kernel vs kernel-per-phase) on device-built stencils."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysolvers_b200 import _native as nat  # noqa: E402
from pysolvers_b200.device import DeviceCSR, ptr, current_stream_ptr  # noqa: E402
from pysolvers_b200.dist import laplacian_block_device  # noqa: E402


def main():
    dim, m, iters = int(sys.argv[1]), int(sys.argv[2]), 200
    n = m ** dim
    dev = torch.device('cuda')
    ip, cols, data = laplacian_block_device(dim, 0.0, 1.0, m, 0, n, dev)
    A = DeviceCSR(indptr=ip, indices=cols.to(torch.int32), data=data, shape=(n, n))
    del cols
    lib = nat.lib()
    b = torch.ones(n, dtype=torch.float64, device=dev)
    x = torch.empty_like(b)
    wb = int(lib.psb_pcg_workspace_bytes(n, 0))
    work = torch.empty(wb, dtype=torch.uint8, device=dev)
    hist = torch.empty(iters, dtype=torch.float64, device=dev)
    res = nat.SolveResult()

    def step():
        nat.check(lib.psb_pcg_solve(A.handle, None, ptr(b), ptr(x), ptr(work), wb, iters, 0.0, 0, ptr(hist),
                                    C.byref(res), current_stream_ptr()))
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 3 / iters * 1e3
    nbytes = 12 * A.nnz + 4 * (n + 1) + 88 * n
    print('dim=%d m=%d n=%d mega=%s: %.1f us/iter, %.0f GB/s (%.3f of 6560), last residual %.17g'
          % (dim, m, n, os.environ.get('PSB_PCG_MEGA', '1'), us, nbytes / us / 1e3, nbytes / us / 1e3 / 6560.3,
             float(hist[-1].item())))


if __name__ == '__main__':
    main()
