"""Device-side objects of the host layer: torch tensors own the HBM, the C ABI
holds raw pointers into them.

``DeviceCSR`` is the uploaded form of the scipy CSR operand the reference hands
to ``mvmult`` (PySolvers/Linear/IterativeLinearSolver.py:94-106): int32
``indptr``/``indices`` and fp64 ``data`` exactly as scipy stores them -- the
stored (possibly unsorted) column order is kept, because the STREAM SpMV sums
each row in that order to stay bit-identical to scipy's csr_matvec.
"""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp
import torch

from . import _native as nat


def require_cuda():
    if not torch.cuda.is_available():
        raise nat.NativeError(
            'pysolvers_b200 needs a CUDA device: the solve path has no CPU '
            'fallback.')


def current_stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def to_device(a, dtype=torch.float64):
    """Upload a numpy array (or pass through a CUDA tensor) as a contiguous
    1-D tensor of ``dtype``."""
    if isinstance(a, torch.Tensor):
        t = a
        if not t.is_cuda:
            t = t.cuda(non_blocking=True)
        return t.to(dtype).contiguous()
    arr = np.ascontiguousarray(a)
    want = {torch.float64: np.float64, torch.int32: np.int32}[dtype]
    if arr.dtype != want:
        arr = arr.astype(want)
    if not arr.flags.writeable:
        arr = arr.copy()
    return torch.from_numpy(arr).cuda(non_blocking=True)


def to_host(t):
    """Device tensor -> fresh numpy array.  Large vectors land in page-locked memory (torch's
    caching host allocator): a pageable destination costs ~60 ms per 134 MB in first-touch
    page faults, the pinned one 2.5 ms (measured, tools/e2e_probe.py).  The array keeps the
    block alive and hands it back to the cache when it is garbage-collected."""
    if t.is_cuda and t.numel() >= (1 << 17):
        stage = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        stage.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return stage.numpy()
    return t.cpu().numpy()


def as_csr(A):
    """Accept what the reference accepts as a matrix operand: a scipy sparse
    matrix (converted to CSR if needed) or a dense 2-D ndarray.  A dense matrix
    is converted to CSR -- the GPU path is sparse-only."""
    if isinstance(A, DeviceCSR):
        return A
    if sp.issparse(A):
        return A if sp.isspmatrix_csr(A) or isinstance(A, sp.csr_array) else A.tocsr()
    if isinstance(A, np.ndarray) and A.ndim == 2:
        return sp.csr_matrix(A)
    raise TypeError('matrix operand must be a scipy sparse matrix or a 2-D '
                    'numpy array, got %r' % type(A))


class DeviceCSR:
    """CSR matrix resident in HBM plus its psb_csr_t handle."""

    def __init__(self, A=None, indptr=None, indices=None, data=None, shape=None):
        require_cuda()
        if A is not None:
            A = as_csr(A)
            if isinstance(A, DeviceCSR):
                raise TypeError('already a DeviceCSR')
            shape = A.shape
            indptr, indices, data = A.indptr, A.indices, A.data
        if len(data) >= 2**31 - 1 or max(shape) >= 2**31 - 1:
            raise nat.NativeError('matrices that need int64 indices (nnz or n '
                                  '>= 2^31) are not supported by the GPU path')
        self.shape = (int(shape[0]), int(shape[1]))
        self.indptr = to_device(indptr, torch.int32)
        self.indices = to_device(indices, torch.int32)
        self.data = to_device(data, torch.float64)
        self.nnz = int(self.data.numel())
        if self.indptr.numel() != self.shape[0] + 1:
            raise ValueError('indptr has the wrong length')
        self._h = C.c_void_p()
        nat.check(nat.lib().psb_csr_create(
            self.shape[0], self.shape[1], self.nnz, ptr(self.indptr),
            ptr(self.indices), ptr(self.data), current_stream_ptr(),
            C.byref(self._h)), 'psb_csr_create')

    @property
    def handle(self):
        return self._h

    def info(self):
        buf = (C.c_int64 * 8)()
        nat.check(nat.lib().psb_csr_info(self._h, buf), 'psb_csr_info')
        return dict(kind=int(buf[0]), max_row=int(buf[1]), max_tile_nnz=int(buf[2]),
                    rows_per_tile=int(buf[3]), vec_width=int(buf[4]),
                    max_grid=int(buf[5]), vec_loads=bool(buf[6] & 1), cols16=bool(buf[6] & 2),
                    max_tile_nnz_512=int(buf[7]))

    def set_kind(self, kind):
        nat.check(nat.lib().psb_csr_set_kind(self._h, kind), 'psb_csr_set_kind')

    def matvec(self, x, out=None):
        """y = A x on the device; x, y are CUDA fp64 tensors."""
        if out is None:
            out = torch.empty(self.shape[0], dtype=torch.float64, device=x.device)
        nat.check(nat.lib().psb_spmv(self._h, ptr(x), ptr(out),
                                     current_stream_ptr()), 'psb_spmv')
        return out

    def to_scipy(self):
        """Download as a scipy csr_matrix (same stored order)."""
        return sp.csr_matrix((self.data.cpu().numpy(), self.indices.cpu().numpy(),
                              self.indptr.cpu().numpy()), shape=self.shape)

    def algorithmic_bytes(self):
        """HBM bytes one SpMV must move."""
        n, m = self.shape
        return 12 * self.nnz + 4 * (n + 1) + 8 * m + 8 * n

    def __del__(self):
        try:
            if self._h:
                nat.lib().psb_csr_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def host_csr(A):
    """scipy CSR (int32 indices, fp64 data, C-contiguous) of a host matrix."""
    M = sp.csr_matrix(A)
    return (np.ascontiguousarray(M.indptr, dtype=np.int32),
            np.ascontiguousarray(M.indices, dtype=np.int32),
            np.ascontiguousarray(M.data, dtype=np.float64), M.shape)


class DeviceTrsv:
    """A triangular factor analysed into level sets and resident in HBM
    (psb_trsv_t).  ``T`` is a host scipy matrix in CSR stored order."""

    def __init__(self, T, lower, unit_diag=False):
        require_cuda()
        indptr, indices, data, shape = host_csr(T)
        assert shape[0] == shape[1]
        self.n = int(shape[0])
        self.lower, self.unit_diag = bool(lower), bool(unit_diag)
        self._h = C.c_void_p()
        nat.check(nat.lib().psb_trsv_create(
            self.n, indptr.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p),
            data.ctypes.data_as(C.c_void_p), 1 if lower else 0, 1 if unit_diag else 0,
            current_stream_ptr(), C.byref(self._h)), 'psb_trsv_create')

    @property
    def handle(self):
        return self._h

    def info(self):
        buf = (C.c_int64 * 8)()
        nat.check(nat.lib().psb_trsv_info(self._h, buf), 'psb_trsv_info')
        return dict(n=int(buf[0]), levels=int(buf[1]), nnz_off=int(buf[2]),
                    nnz_packed=int(buf[3]), lower=bool(buf[4]), unit_diag=bool(buf[5]),
                    groups=int(buf[6]))

    def info2(self):
        buf = (C.c_int64 * 8)()
        nat.check(nat.lib().psb_trsv_info2(self._h, buf), 'psb_trsv_info2')
        return dict(kernel={0: 'grid', 1: 'cta', 2: 'cluster'}[int(buf[0])], wslots=int(buf[1]), n_far=int(buf[2]),
                    max_dist=int(buf[3]), forced=int(buf[4]), stage_len=int(buf[5]), cluster_ok=bool(buf[6]))

    def set_kernel(self, kernel):
        """'grid' (hand-over through L2), 'cta' (one CTA, shared-memory window), 'cluster' (4 CTAs,
        window replicated through distributed shared memory) or None (analysis)."""
        k = {'grid': 0, 'cta': 1, 'cluster': 2, None: -1}[kernel]
        nat.check(nat.lib().psb_trsv_set_kernel(self._h, k), 'psb_trsv_set_kernel')

    def levels(self):
        """(level_ptr, level_rows) as int32 numpy arrays."""
        nlev = self.info()['levels']
        lptr = np.zeros(nlev + 1, dtype=np.int32)
        rows = np.zeros(self.n, dtype=np.int32)
        nat.check(nat.lib().psb_trsv_get_levels(
            self._h, lptr.ctypes.data_as(C.c_void_p), rows.ctypes.data_as(C.c_void_p)),
            'psb_trsv_get_levels')
        return lptr, rows

    def solve(self, b, out=None):
        """x = T^-1 b; b, x CUDA fp64 tensors."""
        if out is None:
            out = torch.empty_like(b)
        nat.check(nat.lib().psb_trsv_solve(self._h, ptr(b), ptr(out), current_stream_ptr()),
                  'psb_trsv_solve')
        return out

    def check(self):
        flag = C.c_int32(0)
        nat.check(nat.lib().psb_trsv_error(self._h, C.byref(flag)), 'psb_trsv_error')
        if flag.value:
            raise nat.NativeError('triangular solve: a dependency never became ready')

    def __del__(self):
        try:
            if self._h:
                nat.lib().psb_trsv_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


class DevicePrec:
    """Owner of a psb_prec_t plus whatever it references (kept alive here)."""

    def __init__(self, handle, n, keep=()):
        self._h = handle
        self.n = n
        self._keep = tuple(keep)

    @property
    def handle(self):
        return self._h

    def apply(self, r, out=None):
        if out is None:
            out = torch.empty_like(r)
        nat.check(nat.lib().psb_prec_apply(self._h, ptr(r), ptr(out), current_stream_ptr()),
                  'psb_prec_apply')
        return out

    def apply_host(self, vec):
        """numpy in, numpy out (upload, apply on the device, download)."""
        v = np.asarray(vec, dtype=np.float64)
        out = self.apply(to_device(v))
        return to_host(out)

    def __del__(self):
        try:
            if self._h:
                nat.lib().psb_prec_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


# rows of an LU factor pair handled as explicitly inverted dense blocks: doubling the block roughly
# halves the levels left for the sparse triangular solves (~1 us each, twice per solve) and
# quadruples the GEMV bytes (8 192 rows: 0.27 GB per factor, 60 us).  With the supernodes of the
# sparse part collapsed (below) 8 192 rows are the better choice at every size measured (Bratu 2048^2
# coarse LU: 16 384 rows -> 65 levels, 1.30 ms per solve; 8 192 rows -> 82 levels, 1.16 ms).
DENSE_TAIL = 8192
COLLAPSE_MIN_LEVELS = 150     # collapse the supernodes of a sparse leading block only if it has this many levels
COLLAPSE_MIN_ROWS = 4         # ... and only supernodes of at least this many rows (Bratu 2048^2 coarse LU: 8 -> 82
                              # levels, 0.785 ms per solve; 4 -> 58 levels, 0.767 ms; 2 -> 56 levels, 0.767 ms; same host time)
DENSE_TAIL_LARGE = int(os.environ.get('PSB_DENSE_TAIL_LARGE', '8192'))     # n >= DENSE_TAIL_LARGE_N
DENSE_TAIL_LARGE_N = 300000
# Guard of the explicitly inverted dense blocks: applying inv(T22) as a GEMV instead of a triangular
# solve loses about log10(cond_1(T22)) digits in the worst case.  The Galerkin / ILUT blocks seen so far
# are at 20 - 60; a block beyond this limit is not inverted (the split falls back to one dense row, i.e.
# the sparse triangular solve does the work) and a warning names the condition number.
DENSE_BLOCK_COND_LIMIT = float(os.environ.get('PSB_DENSE_COND_LIMIT', '1e8'))


def tri_levels(T, lower):
    """Dependency levels of a triangular scipy CSR matrix (host arithmetic in the library)."""
    T = sp.csr_matrix(T)
    n = T.shape[0]
    ip = np.ascontiguousarray(T.indptr, dtype=np.int32)
    ix = np.ascontiguousarray(T.indices, dtype=np.int32)
    lev = np.zeros(n, dtype=np.int32)
    nat.check(nat.lib().psb_tri_levels(n, ip.ctypes.data_as(C.c_void_p), ix.ctypes.data_as(C.c_void_p),
                                       1 if lower else 0, lev.ctypes.data_as(C.c_void_p)), 'psb_tri_levels')
    return lev


def tri_heights_upper(U):
    """h(i) = dependency level of row i in U^T for an upper triangular scipy CSR matrix, without
    forming the transpose (host arithmetic in the library)."""
    U = sp.csr_matrix(U)
    n = U.shape[0]
    ip = np.ascontiguousarray(U.indptr, dtype=np.int32)
    ix = np.ascontiguousarray(U.indices, dtype=np.int32)
    h = np.zeros(n, dtype=np.int32)
    nat.check(nat.lib().psb_tri_heights_upper(n, ip.ctypes.data_as(C.c_void_p), ix.ctypes.data_as(C.c_void_p),
                                              h.ctypes.data_as(C.c_void_p)), 'psb_tri_heights_upper')
    return h


def _dense_block_by_level(T, lower, tail):
    """Symmetric permutation q (new position -> old row) that moves the dense block of a
    triangular factor to the end, and the number n1 of rows left in front.  The block holds the
    rows at the thin top of the elimination tree, chosen by HEIGHT: h(i) = length of the longest
    chain of rows below i that lead to it -- the dependency level of row i for a lower factor, the
    level of row i in the transposed (lower) matrix for an upper factor (the rows that the backward
    solve needs first and on which the longest chains of dependents hang).  Block = {h >= h*} with
    the smallest h* that keeps it within ``tail`` rows.  Such a set contains everything its members
    feed (lower) / need (upper), so moving it to the end keeps the factor triangular, and the
    sparse triangular solve on the rest has h* levels -- not the levels of an arbitrary trailing
    index range (Bratu 1024^2 coarse L, 8 192 dense rows: 218 instead of 1 079)."""
    n = T.shape[0]
    h = tri_levels(T, True) if lower else tri_heights_upper(T)
    counts = np.bincount(h)
    above = np.cumsum(counts[::-1])[::-1]                     # rows with h >= level
    ok = np.flatnonzero(above <= tail)
    dense = h >= (int(ok[0]) if ok.size else len(counts))
    if not dense.any():             # the top level alone exceeds the budget: just the last row (it feeds
        dense[n - 1] = True         # nobody in a lower factor and needs nobody in an upper one)
    q = np.concatenate([np.flatnonzero(~dense), np.flatnonzero(dense)])
    return q, int((~dense).sum())


class DeviceSplitLU(DevicePrec):
    """x = Pc U^-1 L^-1 Pr v for a SuperLU factorisation, with up to ``tail`` rows of L and of U
    as dense, explicitly inverted blocks (psb_splitlu2_create).  With a fill-reducing ordering the
    top separators of the elimination tree form an almost dense triangle in which every row is
    a dependency level of its own for a sparse triangular solve.  The blocks are chosen by level
    (``_dense_block_by_level``), independently for L and U; they are well conditioned (cond
    20 - 60, measured), the result agrees with SuperLU.solve to 5e-16.  The inverses are formed
    once on the device (setup; torch.linalg.solve_triangular); the apply is our own kernels.
    ``by_level=False`` takes the trailing ``tail`` rows instead (the first version; kept for
    comparison).  ``collapse``: supernodal collapse of the sparse leading blocks
    (Linear/supernodes.py) -- None: when they still have >= COLLAPSE_MIN_LEVELS levels; True / False:
    always / never."""

    def __init__(self, lu, tail=None, by_level=True, collapse=None):
        require_cuda()
        n = int(lu.shape[0])
        if tail is None:
            tail = DENSE_TAIL_LARGE if n >= DENSE_TAIL_LARGE_N else DENSE_TAIL
        tail = max(1, min(n, int(tail)))
        # Host preparation of the two factors (format conversion, heights, permutation, blocks,
        # supernodal collapse: scipy / numpy code that releases the GIL) runs for L and U side by
        # side; everything that touches the device happens afterwards on this thread.
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=2) as pool:
            fut_U = pool.submit(self._prepare_side, lu.U, False, n, tail, by_level, collapse)
            pL = self._prepare_side(lu.L, True, n, tail, by_level, collapse)
            pU = fut_U.result()
        qL, n1L, ident_L = pL['q'], pL['n1'], pL['ident']
        qU, n1U, ident_U = pU['q'], pU['n1'], pU['ident']
        dev = torch.device('cuda', torch.cuda.current_device())

        def inverse(dense, upper, unit):
            k = dense.shape[0]
            eye = torch.eye(k, dtype=torch.float64, device=dev)
            d = torch.from_numpy(dense).to(dev)
            return torch.linalg.solve_triangular(d, eye, upper=upper, unitriangular=unit).contiguous()
        self.invL22 = inverse(pL['T22'], False, True)
        self.invU22 = inverse(pU['T22'], True, False)

        def cond1(dense, inv):
            d = torch.from_numpy(dense).to(dev)
            return float(torch.linalg.matrix_norm(d, ord=1) * torch.linalg.matrix_norm(inv, ord=1))
        self.cond_dense = (cond1(pL['T22'], self.invL22), cond1(pU['T22'], self.invU22))
        if max(self.cond_dense) > DENSE_BLOCK_COND_LIMIT and tail > 1:
            import warnings
            warnings.warn('DeviceSplitLU: dense block ill conditioned (cond_1 = %.2e / %.2e > %.1e): not '
                          'inverted, the triangular solves take over' % (self.cond_dense + (DENSE_BLOCK_COND_LIMIT,)))
            del self.invL22, self.invU22, pL, pU
            DeviceSplitLU.__init__(self, lu, tail=1, by_level=by_level, collapse=collapse)
            return
        self.L11 = self.U11 = self.L21 = self.U12 = None
        bd_L, bd_U = pL['bd'], pU['bd']
        if n1L > 0:
            self.L11 = DeviceTrsv(pL['T11'], lower=True, unit_diag=True)
            self.L21 = DeviceCSR(pL['Toff'])
        if n1U > 0:
            self.U11 = DeviceTrsv(pU['T11'], lower=False)
            self.U12 = DeviceCSR(pU['Toff'])
        self.collapsed = (0 if bd_L is None else bd_L[5], 0 if bd_U is None else bd_U[5])
        del pL, pU
        # (Pr v)[perm_r[i]] = v[i]  ->  L position p (row qL[p]) takes v[iperm_r[qL[p]]]
        # (Pc z)[i] = z[perm_c[i]]  ->  U position p (row qU[p]) goes to result[iperm_c[qU[p]]]
        ipr = np.empty(n, dtype=np.int64)
        ipr[np.asarray(lu.perm_r)] = np.arange(n)
        ipc = np.empty(n, dtype=np.int64)
        ipc[np.asarray(lu.perm_c)] = np.arange(n)
        map_in = np.ascontiguousarray(ipr[qL], dtype=np.int32)
        map_out = np.ascontiguousarray(ipc[qU], dtype=np.int32)
        map_mid = None
        if not (ident_L and ident_U) or n1L != n1U:
            posL = np.empty(n, dtype=np.int64)
            posL[qL] = np.arange(n)
            map_mid = np.ascontiguousarray(posL[qU], dtype=np.int32)
        h = C.c_void_p()
        hd = lambda o: None if o is None else o.handle
        nat.check(nat.lib().psb_splitlu2_create(
            n, n1L, n1U, hd(self.L11), hd(self.U11), hd(self.L21), hd(self.U12), ptr(self.invL22),
            ptr(self.invU22), map_in.ctypes.data_as(C.c_void_p),
            None if map_mid is None else map_mid.ctypes.data_as(C.c_void_p),
            map_out.ctypes.data_as(C.c_void_p), current_stream_ptr(), C.byref(h)), 'psb_splitlu2_create')
        super().__init__(h, n, keep=(self.L11, self.U11, self.L21, self.U12, self.invL22, self.invU22))
        self.n1, self.n2 = n1L, n - n1L
        self.n1U, self.n2U = n1U, n - n1U
        for upper, bd in ((0, bd_L), (1, bd_U)):
            if bd is not None:
                row0, c_lo, c_hi, off, vals, _ = bd
                nat.check(nat.lib().psb_splitlu_set_blockdiag(
                    h, upper, row0.shape[0], row0.ctypes.data_as(C.c_void_p), c_lo.ctypes.data_as(C.c_void_p),
                    c_hi.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                    vals.ctypes.data_as(C.c_void_p), vals.shape[0], current_stream_ptr()),
                    'psb_splitlu_set_blockdiag')

    @staticmethod
    def _prepare_side(T, lower, n, tail, by_level, collapse):
        """Host-only preparation of one factor: CSR, dense block chosen (by height or trailing),
        symmetric permutation, the four blocks, supernodal collapse of the sparse leading block."""
        T = T.tocsr()
        if by_level and tail < n:
            q, n1 = _dense_block_by_level(T, lower, tail)
        else:
            q, n1 = np.arange(n), n - tail
        ident = bool(np.array_equal(q, np.arange(n)))
        if not ident:
            T = T[q][:, q].tocsr()
        out = dict(q=q, n1=n1, ident=ident, T22=T[n1:, n1:].toarray(), T11=None, Toff=None, bd=None)
        if n1 > 0:
            T11 = T[:n1, :n1].tocsr()
            if collapse is not False:
                T11, out['bd'] = DeviceSplitLU._collapse(T11, lower, collapse)
            out['T11'] = T11
            out['Toff'] = (T[n1:, :n1] if lower else T[:n1, n1:]).tocsr()
        return out

    @staticmethod
    def _collapse(T11, lower, collapse):
        """Supernodal collapse of a sparse leading block (Linear/supernodes.py) when it still has many
        dependency levels: (transformed block, packed block-diagonal stage) or (T11, None)."""
        from .Linear import supernodes as SN
        n1 = T11.shape[0]
        if collapse is None and (n1 < 2 or int(tri_levels(T11, lower).max()) + 1 < COLLAPSE_MIN_LEVELS):
            return T11, None
        M = sp.csc_matrix(T11) if lower else sp.csr_matrix(T11)       # line-major: columns of L, rows of U
        M.sort_indices()
        line = np.repeat(np.arange(n1), np.diff(M.indptr))
        if not np.array_equal(M.indices[M.indptr[:-1]], np.arange(n1)):
            return T11, None                                            # a line without its diagonal first: leave it
        ptr2, idx2, dat2, blocks = SN.collapse(M.indptr, M.indices, M.data, n1, s_min=COLLAPSE_MIN_ROWS)
        del line
        if not blocks:
            return T11, None
        cls = sp.csc_matrix if lower else sp.csr_matrix
        T2 = cls((dat2, idx2, ptr2), shape=(n1, n1)).tocsr()
        row0, c_lo, c_hi, off, vals = SN.pack_blocks(blocks, n1, transpose=lower)
        return T2, (row0, c_lo, c_hi, np.ascontiguousarray(off, dtype=np.int64), vals, len(blocks))

    def levels(self):
        """(levels of L11, levels of U11): what is left for the sparse triangular solves."""
        return (0 if self.L11 is None else self.L11.info()['levels'],
                0 if self.U11 is None else self.U11.info()['levels'])
