"""A/B timing of the SpMV kernel variants on the C3/C4-shaped matrices
(CUDA events, 30 launches after 5 warm-ups; inputs much larger than L2)."""
import os
import sys
import ctypes as C

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysolvers_b200 import _native as nat  # noqa: E402
from pysolvers_b200.device import DeviceCSR, ptr, current_stream_ptr  # noqa: E402
from pysolvers_b200.problems import fd_laplacian_2d, fd_laplacian_3d  # noqa: E402


def time_kind(dA, kind, x, y, dot, reps=30, epi='dot'):
    dA.set_kind(kind)
    lib = nat.lib()
    st = current_stream_ptr()

    def go():
        if epi == 'dot':
            nat.check(lib.psb_spmv_dot(dA.handle, ptr(x), ptr(y), ptr(dot), st))
        else:
            nat.check(lib.psb_spmv(dA.handle, ptr(x), ptr(y), st))
    for _ in range(5):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    which = sys.argv[1:] or ['2d']
    for w in which:
        if w == '2d':
            A = -fd_laplacian_2d(0.0, 1.0, 4096)
        elif w == '3d':
            A = fd_laplacian_3d(0.0, 1.0, 256)
        else:
            continue
        n = A.shape[0]
        dA = DeviceCSR(A)
        x = torch.rand(n, dtype=torch.float64, device='cuda')
        y = torch.empty(n, dtype=torch.float64, device='cuda')
        dot = torch.zeros(1, dtype=torch.float64, device='cuda')
        want = None
        bytes_ = 12 * A.nnz + 4 * (n + 1) + 16 * n
        for name, kind in (('bulk256', nat.SPMV_STREAM), ('bulk512', nat.SPMV_STREAM | nat.SPMV_TILE512),
                           ('lsu256', nat.SPMV_STREAM_LSU), ('lsu512', nat.SPMV_STREAM_LSU | nat.SPMV_TILE512),
                           ('vector', nat.SPMV_VECTOR)):
            ms = time_kind(dA, kind, x, y, dot)
            got = y.clone()
            if want is None:
                want = got
            same = bool(torch.equal(got, want)) if (kind & 15) != nat.SPMV_VECTOR else None
            print('%s %-8s %8.1f us  %7.1f GB/s  frac %.3f  bit-equal-to-first=%s'
                  % (w, name, ms * 1e3, bytes_ / ms / 1e6, bytes_ / ms / 1e6 / 6560.3, same))


if __name__ == '__main__':
    main()
