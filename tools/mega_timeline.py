"""Per-CTA phase timeline of the persistent PCG kernel (psb_debug_mega_timeline).

    python tools/mega_timeline.py [--gridm 1448] [--iters 200] [--first 100] [--count 8] [--out PREFIX]
    python -m torch.distributed.run --nproc-per-node N ... tools/mega_timeline.py --gridm 4096

Records %globaltimer at the five phase boundaries of iterations [first, first+count) for every
CTA (of every rank), prints where an iteration's time goes -- phase A work, wait at the p.Ap
barrier, phase B work, wait at the r.r barrier, split into "slowest CTA still working" and
"pure barrier latency" -- and saves the raw stamps as PREFIX.rank<r>.npy.
m = 1448 on one GPU has the per-GPU problem size of C3 on 8 GPUs (2.1 M rows).
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pysolvers_b200 import _native as nat  # noqa: E402
from pysolvers_b200 import dist as pdist  # noqa: E402
from pysolvers_b200.device import DeviceCSR, current_stream_ptr, ptr  # noqa: E402
from pysolvers_b200.problems import device_fd_laplacian  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gridm', dest='m', type=int, default=1448)
    ap.add_argument('--dim', type=int, default=2)
    ap.add_argument('--iters', type=int, default=200)
    ap.add_argument('--first', type=int, default=100)
    ap.add_argument('--count', type=int, default=8)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'mega_timeline'))
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    lib = nat.lib()
    dev = torch.device('cuda', local)
    n = args.m ** args.dim
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
        comm = pdist.Comm()
    starts = pdist.row_starts(n, world)
    lo, hi = int(starts[rank]), int(starts[rank + 1])
    n_loc = hi - lo
    ip, ix, dt = device_fd_laplacian(args.dim, 0.0, 1.0, args.m, negate=(args.dim == 2), row_lo=lo, row_hi=hi,
                                     raw=True)
    b = torch.ones(n_loc, dtype=torch.float64, device=dev)
    x = torch.empty(n_loc, dtype=torch.float64, device=dev)
    hist = torch.empty(args.iters, dtype=torch.float64, device=dev)
    res = nat.SolveResult()
    st = current_stream_ptr()
    if world == 1:
        A = DeviceCSR(indptr=ip, indices=ix, data=dt, shape=(n, n))
        wb = int(lib.psb_pcg_workspace_bytes(n, 0))
        work = torch.empty(wb, dtype=torch.uint8, device=dev)

        def step():
            nat.check(lib.psb_pcg_solve(A.handle, None, ptr(b), ptr(x), ptr(work), wb, args.iters, 0.0, 0,
                                        ptr(hist), C.byref(res), st))
    else:
        D = pdist.DistCSR(comm, ip, ix, dt, lo, hi, n)
        wb = int(lib.psb_dist_pcg_workspace_bytes(n_loc, D.n_halo))
        work = torch.empty(wb, dtype=torch.uint8, device=dev)

        def step():
            nat.check(lib.psb_dist_pcg_solve(D.handle, ptr(b), ptr(x), ptr(work), wb, args.iters, 0.0, 0,
                                             ptr(hist), C.byref(res), st))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    us_iter = 1e3 * e0.elapsed_time(e1) / (args.reps * args.iters)

    max_grid = 148 * 16
    buf = torch.zeros(args.count * max_grid * 6, dtype=torch.int64, device=dev)
    nat.check(lib.psb_debug_mega_timeline(ptr(buf), args.first, args.count))
    step()
    torch.cuda.synchronize()
    nat.check(lib.psb_debug_mega_timeline(None, 0, 0))
    raw = buf.cpu().numpy()
    # the grid size is not exported: find it as the largest g for which [count][g][6] has its slot 0 filled
    grid = None
    for g in range(max_grid, 0, -1):
        t = raw[:args.count * g * 6].reshape(args.count, g, 6)
        if np.all(t[:, :, 0] > 0) and np.all(t[:, :, 4] > 0):
            grid = g
            break
    if grid is None:
        print('rank %d: no stamps recorded (was the persistent kernel used?)' % rank)
        return 1
    t = raw[:args.count * grid * 6].reshape(args.count, grid, 6)[:, :, :5].astype(np.float64) * 1e-3   # us
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    np.save('%s.rank%d.npy' % (args.out, rank), t)
    # per iteration: everything relative to the earliest phase-A start
    rows = []
    for k in range(args.count - 1):
        a0, a1, r1, b1, r2 = (t[k, :, j] for j in range(5))
        nxt = t[k + 1, :, 0]
        base = a0.min()
        rows.append(dict(
            iter_us=nxt.min() - base,
            phaseA_med=np.median(a1 - a0), phaseA_max_end=a1.max() - base,
            red1_latency=np.median(r1) - a1.max(),          # last CTA done -> typical CTA released
            phaseB_med=np.median(b1 - r1), phaseB_max_end=b1.max() - np.median(r1),
            red2_latency=np.median(r2) - b1.max(),
            a_spread=a1.max() - np.median(a1), b_spread=b1.max() - np.median(b1)))
    keys = list(rows[0])
    med = {k: float(np.median([r[k] for r in rows])) for k in keys}
    print('rank %d/%d  n_loc=%d grid=%d  %.2f us/iteration (events)  timeline medians over %d iterations [us]:'
          % (rank, world, n_loc, grid, us_iter, len(rows)))
    print('   ' + '  '.join('%s=%.2f' % (k, med[k]) for k in keys), flush=True)
    if comm is not None:
        import torch.distributed as dist
        dist.barrier()
        if world > 1:
            D.close()
        comm.close()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
