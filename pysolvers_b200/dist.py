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
import ctypes as C
import os

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
    """Renumber the columns of a row block and list its halo.

    ``indptr`` / ``indices`` are torch tensors (CPU or CUDA) of the block with
    GLOBAL column ids; ``[lo, hi)`` is the range of column ids (entries of the
    operator's INPUT vector) this rank owns and ``starts`` the partition of that
    vector -- for a square operator the block's own row range and the row
    partition.  Returns dict(local_indices int32, recv int64 (sorted
    global ids), recv_owner int64, r0, r1) with [r0, r1) the largest
    TILE-aligned row range around the middle whose rows touch no halo column.
    """
    dev = indices.device
    n_loc = int(hi - lo)
    n_rows = int(indptr.numel()) - 1
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
    mid = n_rows // 2
    low = touched[touched < mid]
    high = touched[touched >= mid]
    r0 = int(low.max().item()) + 1 if low.numel() else 0
    r1 = int(high.min().item()) if high.numel() else n_rows
    r0 = -(-r0 // TILE) * TILE
    if r1 != n_rows:
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
        # partition plans (halo lists, peer-memory mappings, device structure) of the matrices
        # solved on this communicator, most recently used last; see DistCSR
        self._plans = []
        self.max_plans = int(os.environ.get('PSB_DIST_PLAN_CACHE', '4'))

    @property
    def handle(self):
        return self._h

    def _evict(self, keep):
        """Collective (called at the same point on every rank): destroy the oldest idle plans."""
        idle = [p for p in self._plans if not p.in_use]
        while len(self._plans) > keep and idle:
            victim = idle.pop(0)
            self._plans.remove(victim)
            victim.destroy()

    def close(self):
        """Collective: releases every cached plan (peer mappings are closed only after all
        ranks have stopped using them) and the communicator."""
        for p in list(self._plans):
            p.destroy()
        self._plans = []
        if self._h:
            nat.lib().psb_comm_destroy(self._h)
            self._h = C.c_void_p()


class _DistPlan:
    """Everything about a row block that depends only on its sparsity STRUCTURE: device copies
    of indptr / global and local column ids, halo lists, send lists, the psb_dist handle with
    its NCCL plan and NVLink peer-memory mappings.  Built collectively once per structure and
    kept on the communicator; a later DistCSR with the same structure only uploads values."""

    def __init__(self, comm, ip, ix, data, lo, hi, n, mark, n_cols=None, p2p=True):
        import torch.distributed as dist
        self.comm = comm
        self.in_use = False
        self.lo, self.hi, self.n = int(lo), int(hi), int(n)
        self.n_loc = self.hi - self.lo
        self.starts = row_starts(n, comm.world)
        assert self.starts[comm.rank] == self.lo and self.starts[comm.rank + 1] == self.hi
        # partition of the INPUT vector: the row partition for a square operator
        self.n_cols = int(n if n_cols is None else n_cols)
        self.col_starts = self.starts if self.n_cols == self.n else row_starts(self.n_cols, comm.world)
        self.clo, self.chi = int(self.col_starts[comm.rank]), int(self.col_starts[comm.rank + 1])
        self.n_own = self.chi - self.clo
        self.want_p2p = bool(p2p) and self.n_cols == self.n
        dev = ix.device
        self.ip_global, self.ix_global = ip, ix
        loc = localize(ip, ix, self.clo, self.chi, self.col_starts)
        mark('localize (halo lists)')
        self.recv = loc['recv'].cpu().numpy()
        self.recv_owner = loc['recv_owner'].cpu().numpy()
        self.n_halo = int(self.recv.size)
        self.r0, self.r1 = loc['r0'], loc['r1']
        gathered = [None] * comm.world
        dist.all_gather_object(gathered, (self.recv, self.recv_owner))
        self.send = send_lists(comm.rank, self.clo, gathered)
        mark('all_gather halo lists')
        self.A = DeviceCSR(indptr=ip.to(torch.int32), indices=loc['local_indices'], data=data,
                           shape=(self.n_loc, self.n_own + self.n_halo))
        mark('csr_create')
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
        mark('psb_dist_create')
        self.p2p = False
        if self.want_p2p and os.environ.get('PSB_DIST_MODE', 'p2p') != 'nccl':
            self._enable_p2p(gathered)
        mark('peer-memory mapping')

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

    def destroy(self):
        """Collective.  A peer's last halo / flag store into this rank's exported region may
        still be in flight when this rank is done: all ranks meet before anything is unmapped."""
        import torch.distributed as dist
        if not self._h:
            return
        torch.cuda.synchronize()
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
        nat.lib().psb_dist_destroy(self._h)
        self._h = C.c_void_p()
        self.A = None


class DistCSR:
    """This rank's row block of a row-partitioned matrix plus its halo plan.

    ``indptr, indices, data``: the block's CSR arrays with GLOBAL column ids
    (numpy arrays or torch tensors, host or device); ``lo, hi``: its row range;
    ``n``: global size.  Construction is collective.  The structure-dependent part (halo
    lists, send lists, peer-memory mappings) is looked up on the communicator: when a block
    with the same indptr / indices was partitioned before, only the arrays are uploaded (and
    compared with the cached structure on the device) -- a Newton iteration or a sequence of
    solves does not rebuild its plan.
    """

    def __init__(self, comm, indptr, indices, data, lo, hi, n, n_cols=None, p2p=True):
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
        self.comm = comm
        dev = torch.device('cuda', torch.cuda.current_device())
        ip = torch.as_tensor(indptr).to(dev, non_blocking=True)
        ix = torch.as_tensor(indices).to(dev, non_blocking=True)
        dt = torch.as_tensor(data).to(dev, non_blocking=True)
        _mark('upload indptr, indices, data')
        plan = None
        for cand in reversed(comm._plans):
            if (not cand.in_use and cand.lo == int(lo) and cand.hi == int(hi) and cand.n == int(n)
                    and cand.n_cols == int(n if n_cols is None else n_cols)
                    and cand.want_p2p == (bool(p2p) and cand.n_cols == cand.n)
                    and cand.ix_global.shape == ix.shape and cand.ix_global.dtype == ix.dtype
                    and cand.ip_global.dtype == ip.dtype):
                plan = cand
                break
        hit = torch.zeros(1, dtype=torch.int32, device=dev)
        if plan is not None:
            same = torch.equal(plan.ix_global, ix) and torch.equal(plan.ip_global, ip)
            hit.fill_(1 if same else 0)
        if comm.world > 1:
            dist.all_reduce(hit, op=dist.ReduceOp.MIN)           # every rank must take the same branch
        if int(hit.item()) == 1:
            plan.A.data.copy_(dt)                                # values only; structure stays
            comm._plans.remove(plan)
            comm._plans.append(plan)
            _mark('plan cache hit: values copied')
        else:
            comm._evict(max(comm.max_plans - 1, 0))
            plan = _DistPlan(comm, ip, ix, dt, lo, hi, n, _mark, n_cols=n_cols, p2p=p2p)
            comm._plans.append(plan)
        plan.in_use = True
        self._plan = plan

    # the plan's fields, under the names the rest of the package uses
    lo = property(lambda self: self._plan.lo)
    hi = property(lambda self: self._plan.hi)
    n = property(lambda self: self._plan.n)
    n_loc = property(lambda self: self._plan.n_loc)
    n_halo = property(lambda self: self._plan.n_halo)
    n_own = property(lambda self: self._plan.n_own)
    n_cols = property(lambda self: self._plan.n_cols)
    starts = property(lambda self: self._plan.starts)
    recv = property(lambda self: self._plan.recv)
    recv_owner = property(lambda self: self._plan.recv_owner)
    send = property(lambda self: self._plan.send)
    r0 = property(lambda self: self._plan.r0)
    r1 = property(lambda self: self._plan.r1)
    A = property(lambda self: self._plan.A)
    p2p = property(lambda self: self._plan.p2p)

    @property
    def handle(self):
        return self._plan._h

    def matvec(self, x_loc):
        """y_loc = (A x)_loc for the distributed vector whose local slice is x_loc."""
        ext = torch.zeros(self.n_own + self.n_halo, dtype=torch.float64, device=x_loc.device)
        ext[:self.n_own] = x_loc
        y = torch.empty(self.n_loc, dtype=torch.float64, device=x_loc.device)
        nat.check(nat.lib().psb_dist_spmv(self.handle, ptr(ext), ptr(y), current_stream_ptr()),
                  'psb_dist_spmv')
        torch.cuda.synchronize()
        return y

    def close(self):
        """Hand the plan back to the communicator's cache (not collective)."""
        if getattr(self, '_plan', None) is not None:
            self._plan.in_use = False
            self._plan = None

    def __del__(self):
        try:
            self.close()
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
