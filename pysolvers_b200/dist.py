"""Row-partitioned (multi-GPU) PCG: host-side plumbing, one process per GPU.

The reference is single-process; the contract for
this path is SURVEY.md section 8e and is pinned by oracle/partition.py:

* contiguous block rows with the ``np.array_split`` boundaries;
* halo (receive) list of a rank = sorted unique global column ids outside its
  row range, grouped by owner; local columns = owned (global - lo) first, then
  halo columns in sorted-global order;
* send list owner -> rank = the same ids as local offsets of the owner.

``torch.distributed`` is used only to bootstrap (NCCL unique id, exchange of
the halo id lists); the solve itself -- halo send/recv overlapped with the
interior SpMV, two scalar all-reduces per iteration -- runs inside
libpysolv_b200 (csrc/dist.cu) on raw NCCL.
"""
import contextlib
import ctypes as C
import io
import json
import os
import time

import numpy as np
import torch

from . import _native as nat
from .core import CommonSolverArgs, IterativeSolver, SolveStatus
from .device import DeviceCSR, current_stream_ptr, ptr, to_device, to_host

TILE = 256     # interior range is aligned to SpMV tiles (and so to the 16-byte bulk copies)


def row_starts(n, nranks):
    """Block-row boundaries, identical to np.array_split(np.arange(n), nranks)."""
    base, rem = divmod(int(n), int(nranks))
    starts = np.zeros(nranks + 1, dtype=np.int64)
    for r in range(nranks):
        starts[r + 1] = starts[r] + base + (1 if r < rem else 0)
    return starts


def localize(indptr, indices, lo, hi, starts):
    """Renumber the columns of the row block [lo, hi) and list its halo.

    ``indptr`` / ``indices`` are torch tensors (CPU or CUDA) of the block with
    GLOBAL column ids.  Returns dict(local_indices int32, recv int64 (sorted
    global ids), recv_owner int64, r0, r1) with [r0, r1) the largest
    TILE-aligned row range around the middle whose rows touch no halo column.
    """
    dev = indices.device
    n_loc = int(hi - lo)
    cols = indices.to(torch.int64)
    off = (cols < lo) | (cols >= hi)
    recv = torch.unique(cols[off])                      # sorted ascending
    st = torch.as_tensor(starts, dtype=torch.int64, device=dev)
    owner = torch.searchsorted(st, recv, right=True) - 1
    halo_pos = torch.searchsorted(recv, cols) if recv.numel() else torch.zeros_like(cols)
    local = torch.where(off, n_loc + halo_pos, cols - lo).to(torch.int32)
    # rows that touch a halo column
    csum = torch.zeros(cols.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(off.to(torch.int64), dim=0, out=csum[1:])
    ip = indptr.to(torch.int64)
    per_row = csum[ip[1:]] - csum[ip[:-1]]
    touched = torch.nonzero(per_row > 0).flatten()
    mid = n_loc // 2
    low = touched[touched < mid]
    high = touched[touched >= mid]
    r0 = int(low.max().item()) + 1 if low.numel() else 0
    r1 = int(high.min().item()) if high.numel() else n_loc
    r0 = -(-r0 // TILE) * TILE
    if r1 != n_loc:
        r1 = (r1 // TILE) * TILE
    if r0 >= r1:
        r0 = r1 = 0                                       # no overlap window: all rows wait
    return dict(local_indices=local, recv=recv, recv_owner=owner, r0=r0, r1=r1)


def send_lists(rank, lo, all_recv):
    """From every rank's (recv ids, owners) derive what ``rank`` must send:
    {peer: int32 local offsets, ascending}."""
    out = {}
    for q, (ids, owners) in enumerate(all_recv):
        if q == rank:
            continue
        mine = np.asarray(ids)[np.asarray(owners) == rank]
        if mine.size:
            out[q] = (mine - lo).astype(np.int32)
    return out


class Comm:
    """NCCL communicator of the solve path (psb_comm_t), bootstrapped through
    torch.distributed."""

    def __init__(self):
        import torch.distributed as dist
        assert dist.is_initialized(), 'initialise torch.distributed first'
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        buf = (C.c_ubyte * 128)()
        if self.rank == 0:
            nat.check(nat.lib().psb_nccl_unique_id(buf), 'psb_nccl_unique_id')
        box = [bytes(buf)]
        dist.broadcast_object_list(box, src=0)
        ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
        self._h = C.c_void_p()
        nat.check(nat.lib().psb_comm_create(ident, self.rank, self.world, C.byref(self._h)),
                  'psb_comm_create')

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            nat.lib().psb_comm_destroy(self._h)
            self._h = C.c_void_p()


class DistCSR:
    """This rank's row block of a row-partitioned matrix plus its halo plan.

    ``indptr, indices, data``: the block's CSR arrays with GLOBAL column ids
    (numpy arrays or torch tensors, host or device); ``lo, hi``: its row range;
    ``n``: global size.
    """

    def __init__(self, comm, indptr, indices, data, lo, hi, n):
        import torch.distributed as dist
        import time as _time
        _timing = os.environ.get('PSB_DIST_TIMING', '0') == '1'
        _t = [_time.perf_counter()]

        def _mark(what):
            if _timing:
                torch.cuda.synchronize()
                now = _time.perf_counter()
                if comm.rank == 0:
                    print('[DistCSR] %-28s %7.2f ms' % (what, 1e3 * (now - _t[0])), flush=True)
                _t[0] = now
        self._mark = _mark
        self.comm = comm
        self.lo, self.hi, self.n = int(lo), int(hi), int(n)
        self.n_loc = self.hi - self.lo
        self.starts = row_starts(n, comm.world)
        assert self.starts[comm.rank] == self.lo and self.starts[comm.rank + 1] == self.hi
        dev = torch.device('cuda', torch.cuda.current_device())
        ip = torch.as_tensor(indptr).to(dev)
        ix = torch.as_tensor(indices).to(dev)
        _mark('upload indptr, indices')
        loc = localize(ip, ix, self.lo, self.hi, self.starts)
        _mark('localize (halo lists)')
        self.recv = loc['recv'].cpu().numpy()
        self.recv_owner = loc['recv_owner'].cpu().numpy()
        self.n_halo = int(self.recv.size)
        self.r0, self.r1 = loc['r0'], loc['r1']
        gathered = [None] * comm.world
        dist.all_gather_object(gathered, (self.recv, self.recv_owner))
        self.send = send_lists(comm.rank, self.lo, gathered)
        _mark('all_gather halo lists')
        self.A = DeviceCSR(indptr=ip.to(torch.int32), indices=loc['local_indices'],
                           data=torch.as_tensor(data).to(dev),
                           shape=(self.n_loc, self.n_loc + self.n_halo))
        del ix
        _mark('upload data, csr_create')
        # peers: union of the ranks we send to / receive from
        peers = sorted(set(self.send) | set(int(o) for o in np.unique(self.recv_owner)))
        k = len(peers)
        peer_rank = (C.c_int32 * max(k, 1))()
        send_off = (C.c_int64 * max(k, 1))()
        send_cnt = (C.c_int64 * max(k, 1))()
        recv_off = (C.c_int64 * max(k, 1))()
        recv_cnt = (C.c_int64 * max(k, 1))()
        idx_ptrs = (C.c_void_p * max(k, 1))()
        self._idx_keep = []
        for i, q in enumerate(peers):
            peer_rank[i] = q
            s = self.send.get(q)
            if s is not None and s.size:
                send_cnt[i] = s.size
                if np.array_equal(s, np.arange(s[0], s[0] + s.size, dtype=s.dtype)):
                    send_off[i] = int(s[0])              # contiguous slice: no pack kernel
                    idx_ptrs[i] = None
                else:
                    t = torch.from_numpy(s).to(dev)
                    self._idx_keep.append(t)
                    idx_ptrs[i] = t.data_ptr()
            sel = np.flatnonzero(self.recv_owner == q)
            if sel.size:
                recv_off[i] = int(sel[0])               # owners are grouped: contiguous
                recv_cnt[i] = int(sel.size)
        self._h = C.c_void_p()
        nat.check(nat.lib().psb_dist_create(
            comm.handle, self.A.handle, self.n_loc, self.n_halo, self.r0, self.r1, k,
            peer_rank, send_off, send_cnt, idx_ptrs, recv_off, recv_cnt, C.byref(self._h)),
            'psb_dist_create')

        _mark('psb_dist_create')
        self.p2p = False
        if os.environ.get('PSB_DIST_MODE', 'p2p') != 'nccl':
            self._enable_p2p(gathered)
        _mark('peer-memory mapping')

    def _enable_p2p(self, all_recv):
        """Map every rank's exported region (NVLink peer memory) so that the solve
        can fuse the halo exchange and the scalar all-reduces into its kernels.
        Needs contiguous send slices (true for slab partitions of banded
        matrices); otherwise the NCCL path stays in force."""
        import torch.distributed as dist
        comm = self.comm
        ok = all(np.array_equal(s, np.arange(s[0], s[0] + s.size, dtype=s.dtype))
                 for s in self.send.values()) and len(self.send) <= 4 and comm.world <= 32 \
            and self.A.info()['kind'] == nat.SPMV_STREAM
        flags = [None] * comm.world
        dist.all_gather_object(flags, bool(ok))
        if not all(flags):
            return
        handle = (C.c_ubyte * 64)()
        layout = (C.c_int64 * 6)()
        nat.check(nat.lib().psb_dist_p2p_alloc(self._h, handle, layout), 'psb_dist_p2p_alloc')
        info = [None] * comm.world
        dist.all_gather_object(info, (bytes(handle), [int(v) for v in layout]))
        handles = (C.c_ubyte * (64 * comm.world))()
        for q, (hb, _) in enumerate(info):
            handles[64 * q:64 * (q + 1)] = list(hb)
        targets = sorted(self.send)
        k = len(targets)
        push_rank = (C.c_int32 * max(k, 1))()
        send_off = (C.c_int64 * max(k, 1))()
        send_cnt = (C.c_int64 * max(k, 1))()
        roff0 = (C.c_int64 * max(k, 1))()
        roff1 = (C.c_int64 * max(k, 1))()
        roffr = (C.c_int64 * max(k, 1))()
        fidx = (C.c_int32 * max(k, 1))()
        for i, q in enumerate(targets):
            s = self.send[q]
            ids_q, owners_q = all_recv[q]
            owners_q = np.asarray(owners_q)
            first = int(np.flatnonzero(owners_q == comm.rank)[0])     # my slice inside q's halo
            lay = info[q][1]
            push_rank[i] = q
            send_off[i] = int(s[0])
            send_cnt[i] = int(s.size)
            roff0[i] = lay[0] + 8 * (lay[3] + first)
            roff1[i] = lay[1] + 8 * (lay[3] + first)
            roffr[i] = lay[2] + 8 * (lay[3] + first)
            fidx[i] = int(np.flatnonzero(np.unique(owners_q) == comm.rank)[0])
        nat.check(nat.lib().psb_dist_p2p_open(self._h, handles, k, push_rank, send_off, send_cnt,
                                              roff0, roff1, roffr, fidx), 'psb_dist_p2p_open')
        dist.barrier()
        self.p2p = True

    @property
    def handle(self):
        return self._h

    def matvec(self, x_loc):
        """y_loc = (A x)_loc for the distributed vector whose local slice is x_loc."""
        ext = torch.zeros(self.n_loc + self.n_halo, dtype=torch.float64, device=x_loc.device)
        ext[:self.n_loc] = x_loc
        y = torch.empty(self.n_loc, dtype=torch.float64, device=x_loc.device)
        nat.check(nat.lib().psb_dist_spmv(self._h, ptr(ext), ptr(y), current_stream_ptr()),
                  'psb_dist_spmv')
        torch.cuda.synchronize()
        return y

    def __del__(self):
        try:
            if self._h:
                nat.lib().psb_dist_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


class DistributedPCG(IterativeSolver):
    """Un-preconditioned PCG on a row-partitioned system; every rank calls
    ``solve(dist_csr, b_local)`` and gets a SolveStatus whose ``soln()`` is its
    local slice of x.  Same loop, exits and history as PCGSolver."""

    def __init__(self, control=CommonSolverArgs(), name='PCG'):
        super().__init__(control, name=name)
        self.last_history = None

    def solve(self, D, b_local, keep_on_device=False):
        self._require_euclidean_norm()
        b_d = to_device(b_local)
        n = D.n_loc
        assert b_d.numel() == n
        lib = nat.lib()
        maxiter = int(self.maxiter())
        wbytes = int(lib.psb_dist_pcg_workspace_bytes(n, D.n_halo))
        work = torch.empty(wbytes, dtype=torch.uint8, device=b_d.device)
        x_d = torch.empty(max(n, 1), dtype=torch.float64, device=b_d.device)
        hist_d = torch.empty(max(maxiter, 1), dtype=torch.float64, device=b_d.device)
        res = nat.SolveResult()
        nat.check(lib.psb_dist_pcg_solve(
            D.handle, ptr(b_d), ptr(x_d), ptr(work), wbytes, maxiter, float(self.tau()),
            1 if self.failOnMaxiter() else 0, ptr(hist_d), C.byref(res), current_stream_ptr()),
            'psb_dist_pcg_solve')
        hist = hist_d[:res.n_hist].cpu().numpy()
        self.last_history = hist
        for k in range(res.n_hist):
            self.reportIter(k, hist[k], res.norm_b)
        x = x_d[:n] if keep_on_device else to_host(x_d[:n])
        if res.status == nat.TRIVIAL:
            return self.handleConvergence(0, x * 0, 0, 0)
        if res.status == nat.BREAKDOWN_PAP:
            return self.handleBreakdown(res.k, 'breakdown dot(p, Ap)==0')
        if res.status == nat.CONVERGED:
            return self.handleConvergence(res.k, x, res.norm_r, res.norm_b)
        return self.handleMaxiter(res.k, x, res.norm_r, res.norm_b)


# ---------------------------------------------------------------------------------
# device-side assembly of the benchmark stencils (input synthesis, torch ops)
# ---------------------------------------------------------------------------------
def laplacian_block_device(dim, a, b, m, lo, hi, device, negate_2d=True):
    """Rows [lo, hi) of the 2-D 5-point (SPD: -FDLaplacian2D) or 3-D 7-point
    Laplacian assembled on the device with the same stored column order and
    the same values as pysolvers_b200.problems.fd_laplacian_2d / _3d."""
    h = np.abs(b - a) / np.double(m + 1)
    k = torch.arange(lo, hi, dtype=torch.int64, device=device)
    ones = torch.ones(k.numel(), dtype=torch.bool, device=device)
    if dim == 2:
        ix, iy = k % m, k // m
        diag = -4.0 / h / h
        offv = 1.0 / h / h
        if negate_2d:
            diag, offv = -diag, -offv
        cols = [k, k - m, k + m, k - 1, k + 1]
        valid = [ones, iy > 0, iy < m - 1, ix > 0, ix < m - 1]
    else:
        ix, iy, iz = k % m, (k // m) % m, k // (m * m)
        diag = 6.0 / h / h
        offv = -1.0 / h / h
        mm = m * m
        cols = [k, k - mm, k + mm, k - m, k + m, k - 1, k + 1]
        valid = [ones, iz > 0, iz < m - 1, iy > 0, iy < m - 1, ix > 0, ix < m - 1]
    V = torch.stack(valid, dim=1)
    counts = V.sum(dim=1)
    indptr = torch.zeros(k.numel() + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, dim=0, out=indptr[1:])
    Cc = torch.stack(cols, dim=1)[V]
    vals_row = torch.full((len(cols),), offv, dtype=torch.float64, device=device)
    vals_row[0] = diag
    data = vals_row.repeat(k.numel(), 1)[V]
    return indptr.to(torch.int32), Cc, data


# ---------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): strong scaling of the metric workload
# ---------------------------------------------------------------------------------
def bench_multi_gpu(args, bench):
    import torch.distributed as dist
    from .csrc.build import build_native
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
    local_rank = int(os.environ.get('LOCAL_RANK', str(rank)))
    torch.cuda.set_device(local_rank)
    build_native()
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    lib = nat.lib()
    comm = Comm()
    dev = torch.device('cuda', local_rank)

    which = os.environ.get('PSB_BENCH_WORKLOAD', 'c3')
    if which == 'c4':
        dim, m = 3, int(os.environ.get('PSB_BENCH_M3', '512'))
        n = m ** 3
        name = ('3-D 7-point Laplacian m=%d (n=%d), un-preconditioned PCG, b=1, %d iterations per step'
                % (m, n, bench.ITERS_PER_STEP))
    else:
        dim, m = 2, bench.M_GRID
        n = m * m
        name = bench.workload_name(m)
    starts = row_starts(n, world)
    lo, hi = int(starts[rank]), int(starts[rank + 1])
    indptr, cols, data = laplacian_block_device(dim, 0.0, 1.0, m, lo, hi, dev)
    nnz_loc = int(data.numel())
    D = DistCSR(comm, indptr, cols, data, lo, hi, n)
    del cols
    torch.cuda.empty_cache()
    n_loc = hi - lo
    b_d = torch.ones(n_loc, dtype=torch.float64, device=dev)
    x_d = torch.empty(n_loc, dtype=torch.float64, device=dev)
    iters = bench.ITERS_PER_STEP
    wbytes = int(lib.psb_dist_pcg_workspace_bytes(n_loc, D.n_halo))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    hist_d = torch.empty(iters, dtype=torch.float64, device=dev)
    res = nat.SolveResult()
    stream = current_stream_ptr()

    def step():
        nat.check(lib.psb_dist_pcg_solve(D.handle, ptr(b_d), ptr(x_d), ptr(work), wbytes, iters, 0.0, 0,
                                         ptr(hist_d), C.byref(res), stream), 'psb_dist_pcg_solve')
        assert res.n_hist == iters

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = bench.ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = nat.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dev_ms = float(ms.item())
    clocks = sampler.finish() if sampler else None
    value = args.steps * iters / (dev_ms * 1e-3)

    # e2e: public API with host operands on every rank (upload block + b, download x)
    skip_e2e = os.environ.get('PSB_BENCH_SKIP_E2E', '0') == '1'
    e2e_value, h2d_total, final_resid = None, 0, float(hist_d[-1].item())
    nnz_t = torch.tensor([nnz_loc], dtype=torch.float64, device=dev)
    dist.all_reduce(nnz_t)
    nnz = int(nnz_t.item())
    mode = 'nvlink-p2p' if D.p2p else 'nccl'
    if not skip_e2e:
        # host operands in page-locked memory, as in the single-GPU bench (the e2e leg copies from
        # pinned memory; pageable arrays cost 2 - 4 x in the upload)
        pin = lambda t: t.cpu().pin_memory().numpy()
        ip_h, dt_h = pin(indptr), pin(data)
        cols_h = pin(laplacian_block_device(dim, 0.0, 1.0, m, lo, hi, dev)[1])
        b_h = torch.ones(n_loc, dtype=torch.float64).pin_memory().numpy()
        solver = DistributedPCG(CommonSolverArgs(maxiter=iters, tau=0.0, failOnMaxiter=False,
                                                 showIters=False, showFinal=False))
        del D
        torch.cuda.empty_cache()

        def api_step():
            Dm = DistCSR(comm, ip_h, cols_h, dt_h, lo, hi, n)
            with contextlib.redirect_stdout(io.StringIO()):
                st = solver.solve(Dm, b_h)
            assert st.success() and st.iters() == iters
            return st
        for _ in range(2):                  # as in the single-GPU leg: the first two solves touch
            api_step()                      # freshly allocated device memory (measured 76 vs 47 ms)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st = api_step()
        torch.cuda.synchronize()
        dist.barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_value = args.steps * iters / float(e2e_s.item())
        h2d = torch.tensor([ip_h.nbytes + cols_h.nbytes + dt_h.nbytes + b_h.nbytes], dtype=torch.float64, device=dev)
        dist.all_reduce(h2d)
        h2d_total = int(h2d.item())
        final_resid = float(st.resid())

    if rank == 0:
        peak, peak_src = bench.peaks()
        iter_bytes = 12 * nnz + 4 * (n + world) + 88 * n
        iter_ms = dev_ms / (args.steps * iters)
        gbs = iter_bytes / (iter_ms * 1e-3) / 1e9
        line = {
            'metric': bench.METRIC, 'value': value, 'unit': bench.UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': dev_ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': name, 'n': n, 'nnz': nnz, 'iters_per_step': iters,
                       'parallelism': 'row partition over %d GPUs, collectives: %s' % (world, mode),
                       'l2': 'inputs larger than L2: %.2f GB touched per iteration per GPU'
                             % (iter_bytes / world / 1e9)},
            'e2e': {'value': e2e_value, 'unit': bench.UNIT, 'h2d_bytes_per_step': h2d_total,
                    'd2h_bytes_per_step': int(8 * n + 8 * iters * world)},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': 'whole PCG iteration (aggregate over GPUs)',
                         'achieved': gbs, 'peak': peak * world, 'unit': 'GB/s',
                         'frac': gbs / (peak * world), 'traffic': None,
                         'bytes_per_iteration': iter_bytes, 'ms_per_iteration': iter_ms,
                         'peak_source': peak_src + ' x %d GPUs' % world},
            'cpu_baseline': None,
            'final_residual': final_resid,
        }
        print(json.dumps(line))
    comm.close()
    dist.destroy_process_group()
    return 0
