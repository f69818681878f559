"""Host-side logic of the multi-GPU path, on CPU: the partition map, halo lists
and local renumbering must match the numpy oracle (oracle/partition.py)
bit-exactly; the exchange of halo lists is exercised with a world_size-2 gloo
group."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import partition as opart
from pysolvers_b200 import dist as pdist
from pysolvers_b200.problems import fd_laplacian_2d, fd_laplacian_3d, load_dh_matrix


def _check_rank(A, nranks, r, all_recv=None):
    A = sp.csr_matrix(A)
    n = A.shape[0]
    ref = opart.partition(A, nranks)
    starts = pdist.row_starts(n, nranks)
    assert np.array_equal(starts, opart.row_starts(n, nranks))
    lo, hi = int(starts[r]), int(starts[r + 1])
    blk = A[lo:hi, :]
    loc = pdist.localize(torch.from_numpy(blk.indptr.astype(np.int64)),
                         torch.from_numpy(blk.indices.astype(np.int64)), lo, hi, starts)
    assert np.array_equal(loc['local_indices'].numpy(), ref[r]['indices'])
    assert np.array_equal(loc['recv'].numpy(), ref[r]['recv'])
    assert np.array_equal(loc['recv_owner'].numpy(), ref[r]['recv_owner'])
    interior, boundary = opart.interior_boundary_rows(ref[r]['indptr'], ref[r]['indices'], hi - lo)
    r0, r1 = loc['r0'], loc['r1']
    assert r0 % 4 == 0 and (r1 % 4 == 0 or r1 == hi - lo) and 0 <= r0 <= r1 <= hi - lo
    assert not np.any((boundary >= r0) & (boundary < r1))      # the window holds interior rows only
    return loc


@pytest.mark.parametrize('nranks', [1, 2, 3, 4, 8])
def test_partition_matches_oracle(nranks):
    mats = [-fd_laplacian_2d(0.0, 1.0, 40), fd_laplacian_3d(0.0, 1.0, 12), load_dh_matrix(9)]
    for A in mats:
        recvs = []
        for r in range(nranks):
            loc = _check_rank(A, nranks, r)
            recvs.append((loc['recv'].numpy(), loc['recv_owner'].numpy()))
        ref = opart.partition(A, nranks)
        starts = pdist.row_starts(A.shape[0], nranks)
        for r in range(nranks):
            mine = pdist.send_lists(r, int(starts[r]), recvs)
            assert sorted(mine) == sorted(ref[r]['send'])
            for q in mine:
                assert np.array_equal(mine[q], ref[r]['send'][q])


@pytest.mark.parametrize('nranks', [2, 3, 8])
def test_rectangular_partition_matches_oracle(nranks):
    """Restriction / prolongation row blocks of the row-partitioned V-cycle: halo lists, local
    renumbering and send lists bit-identical to oracle.partition.partition_rect."""
    from pysolvers_b200.Linear import amg_setup
    A = -fd_laplacian_2d(0.0, 1.0, 24)
    ops, ups, downs = amg_setup.build_hierarchy(A, num_levels=2)
    for M in (downs[0], ups[0]):
        M = sp.csr_matrix(M)
        ref = opart.partition_rect(M, nranks)
        rs, cs = pdist.row_starts(M.shape[0], nranks), pdist.row_starts(M.shape[1], nranks)
        recvs = []
        for r in range(nranks):
            blk = M[int(rs[r]):int(rs[r + 1]), :]
            loc = pdist.localize(torch.from_numpy(blk.indptr.astype(np.int64)),
                                 torch.from_numpy(blk.indices.astype(np.int64)), int(cs[r]), int(cs[r + 1]), cs)
            assert np.array_equal(loc['local_indices'].numpy(), ref[r]['indices'])
            assert np.array_equal(loc['recv'].numpy(), ref[r]['recv'])
            assert np.array_equal(loc['recv_owner'].numpy(), ref[r]['recv_owner'])
            nrows = blk.shape[0]
            assert 0 <= loc['r0'] <= loc['r1'] <= nrows
            _, halo_rows = opart.interior_boundary_rows(ref[r]['indptr'], ref[r]['indices'], int(cs[r + 1] - cs[r]))
            assert not np.any((halo_rows >= loc['r0']) & (halo_rows < loc['r1']))
            recvs.append((loc['recv'].numpy(), loc['recv_owner'].numpy()))
        for r in range(nranks):
            mine = pdist.send_lists(r, int(cs[r]), recvs)
            assert sorted(mine) == sorted(ref[r]['send'])
            for q in mine:
                assert np.array_equal(mine[q], ref[r]['send'][q])


def test_interior_window_is_found_for_slabs():
    A = fd_laplacian_3d(0.0, 1.0, 16)            # 4096 rows, 256-row planes
    starts = pdist.row_starts(A.shape[0], 2)
    blk = sp.csr_matrix(A)[0:int(starts[1]), :]
    loc = pdist.localize(torch.from_numpy(blk.indptr.astype(np.int64)),
                         torch.from_numpy(blk.indices.astype(np.int64)), 0, int(starts[1]), starts)
    # rank 0 only talks to rank 1: rows of its last plane touch the halo
    assert loc['r0'] == 0 and loc['r1'] == int(starts[1]) - 256


def test_device_assembly_matches_host_generators():
    for dim, gen, m in ((2, lambda m: -fd_laplacian_2d(0.0, 1.0, m), 9), (3, lambda m: fd_laplacian_3d(0.0, 1.0, m), 5)):
        A = gen(m)
        n = A.shape[0]
        lo, hi = n // 3, n - 2
        ip, cols, data = pdist.laplacian_block_device(dim, 0.0, 1.0, m, lo, hi, torch.device('cpu'))
        blk = sp.csr_matrix(A)[lo:hi, :]
        assert np.array_equal(ip.numpy(), blk.indptr)
        assert np.array_equal(cols.numpy(), blk.indices)
        assert np.array_equal(data.numpy(), blk.data)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        A = sp.csr_matrix(fd_laplacian_3d(0.0, 1.0, 10))
        n = A.shape[0]
        starts = pdist.row_starts(n, world)
        lo, hi = int(starts[rank]), int(starts[rank + 1])
        blk = A[lo:hi, :]
        loc = pdist.localize(torch.from_numpy(blk.indptr.astype(np.int64)),
                             torch.from_numpy(blk.indices.astype(np.int64)), lo, hi, starts)
        gathered = [None] * world
        dist.all_gather_object(gathered, (loc['recv'].numpy(), loc['recv_owner'].numpy()))
        mine = pdist.send_lists(rank, lo, gathered)
        ref = opart.partition(A, world)[rank]
        ok = sorted(mine) == sorted(ref['send']) and all(np.array_equal(mine[k], ref['send'][k]) for k in mine)
        # a distributed mat-vec with numpy + gloo as the collective: halo exchange semantics
        x = np.random.default_rng(0).random(n)
        ext = np.zeros(hi - lo + loc['recv'].numel())
        ext[:hi - lo] = x[lo:hi]
        for peer in range(world):
            if peer == rank:
                continue
            if peer in mine:
                dist.send(torch.from_numpy(x[lo:hi][mine[peer]].copy()), dst=peer) if rank < peer else None
            sel = np.flatnonzero(loc['recv_owner'].numpy() == peer)
            if sel.size:
                buf = torch.zeros(sel.size, dtype=torch.float64)
                if rank > peer:
                    dist.recv(buf, src=peer)
                    ext[hi - lo + sel] = buf.numpy()
            if peer in mine and rank > peer:
                dist.send(torch.from_numpy(x[lo:hi][mine[peer]].copy()), dst=peer)
            if sel.size and rank < peer:
                buf = torch.zeros(sel.size, dtype=torch.float64)
                dist.recv(buf, src=peer)
                ext[hi - lo + sel] = buf.numpy()
        Aloc = sp.csr_matrix((blk.data, loc['local_indices'].numpy(), blk.indptr),
                             shape=(hi - lo, ext.size))
        ok = ok and np.array_equal(Aloc @ ext, (A @ x)[lo:hi])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
