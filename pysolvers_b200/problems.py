"""Vectorised benchmark-input generators (host side, numpy).

Same matrices, bit for bit, as the reference's O(n) Python/dok generators --
including the *stored column order* per row, which scipy's dok->csr conversion
leaves unsorted -- but built in O(nnz) numpy so the 16.8 M- and 134 M-row
configurations are practical, and restricted to a row slab [row_lo, row_hi)
so every rank of a row-partitioned run assembles only its own rows.

* ``fd_laplacian_2d``  <- examples/FDLaplacian2D.py:5-23 (negative definite as
  shipped; callers negate it for SPD use, examples/FDBratu2D.py:15)
* ``fd_laplacian_3d``  <- builder-defined 7-point extension
* ``FDBratu2D``        <- examples/FDBratu2D.py:10-29
* ``dh_test_problem``  <- examples/DHTestProblem.py:7-36, with a seeded RNG
"""
import os
import numpy as np
import scipy.sparse as sp


def _stencil_csr(n_total, rows, cols_list, vals_list, valid_list):
    """Assemble CSR from per-row candidate entries given in stored order."""
    nrows = rows.shape[0]
    valid = np.stack(valid_list, axis=1)
    counts = valid.sum(axis=1, dtype=np.int64)
    indptr = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    flat = valid.ravel()
    cols = np.stack(cols_list, axis=1).ravel()[flat]
    vals = np.stack(vals_list, axis=1).ravel()[flat]
    nnz = int(indptr[-1])
    idx_t = np.int32 if max(nnz, n_total) < 2**31 else np.int64
    return sp.csr_matrix((vals, cols.astype(idx_t), indptr.astype(idx_t)),
                         shape=(nrows, n_total))


def fd_laplacian_2d(a, b, m, row_lo=0, row_hi=None):
    """5-point Laplacian on an m x m grid, rows [row_lo, row_hi) of m*m.

    Row k = m*iy + ix holds, in this order, [k, k-m, k+m, k-1, k+1] with values
    [-4, 1, 1, 1, 1] / h^2, h = |b-a|/(m+1)  (examples/FDLaplacian2D.py:6-21).
    """
    n = m * m
    row_hi = n if row_hi is None else row_hi
    h = np.abs(b - a) / np.double(m + 1)
    diag = -4.0 / h / h
    off = 1.0 / h / h
    k = np.arange(row_lo, row_hi, dtype=np.int64)
    ix = k % m
    iy = k // m
    ones = np.ones(k.shape[0])
    cols = [k, k - m, k + m, k - 1, k + 1]
    vals = [diag * ones, off * ones, off * ones, off * ones, off * ones]
    valid = [np.ones(k.shape[0], dtype=bool), iy > 0, iy < m - 1,
             ix > 0, ix < m - 1]
    return _stencil_csr(n, k, cols, vals, valid)


def fd_laplacian_3d(a, b, m, row_lo=0, row_hi=None):
    """7-point (positive definite) Laplacian on an m^3 grid, rows
    [row_lo, row_hi).  Row k = m^2*iz + m*iy + ix holds
    [k, k-m^2, k+m^2, k-m, k+m, k-1, k+1] with values [6, -1, ...] / h^2."""
    n = m * m * m
    row_hi = n if row_hi is None else row_hi
    h = np.abs(b - a) / np.double(m + 1)
    diag = 6.0 / h / h
    off = -1.0 / h / h
    k = np.arange(row_lo, row_hi, dtype=np.int64)
    ix = k % m
    iy = (k // m) % m
    iz = k // (m * m)
    ones = np.ones(k.shape[0])
    mm = m * m
    cols = [k, k - mm, k + mm, k - m, k + m, k - 1, k + 1]
    vals = [diag * ones] + [off * ones] * 6
    valid = [np.ones(k.shape[0], dtype=bool), iz > 0, iz < m - 1,
             iy > 0, iy < m - 1, ix > 0, ix < m - 1]
    return _stencil_csr(n, k, cols, vals, valid)


def device_fd_laplacian(dim, a, b, m, negate=False, row_lo=0, row_hi=None, raw=False):
    """The same matrices as ``fd_laplacian_2d`` (dim = 2; ``negate`` gives the SPD
    ``-FDLaplacian2D`` of examples/FDBratu2D.py:15) and ``fd_laplacian_3d`` (dim = 3),
    bit for bit, assembled directly in HBM (psb_stencil_fill): returns a DeviceCSR that the
    solvers accept in place of the scipy matrix -- no host assembly, no upload.  ``raw``: return
    the three device arrays (indptr, indices with GLOBAL column ids, data) instead, e.g. as the
    row block handed to ``dist.DistCSR``."""
    import torch
    from . import _native as nat
    from .device import DeviceCSR, current_stream_ptr, ptr, require_cuda
    require_cuda()
    n = m ** dim
    row_hi = n if row_hi is None else row_hi
    h = np.abs(b - a) / np.double(m + 1)
    if dim == 2:
        diag, off = -4.0 / h / h, 1.0 / h / h
    else:
        diag, off = 6.0 / h / h, -1.0 / h / h
    if negate:
        diag, off = -diag, -off
    nnz = int(nat.lib().psb_stencil_nnz(dim, m, row_lo, row_hi))
    assert nnz >= 0
    indptr = torch.empty(row_hi - row_lo + 1, dtype=torch.int32, device='cuda')
    indices = torch.empty(nnz, dtype=torch.int32, device='cuda')
    data = torch.empty(nnz, dtype=torch.float64, device='cuda')
    nat.check(nat.lib().psb_stencil_fill(dim, m, row_lo, row_hi, float(diag), float(off), ptr(indptr),
                                         ptr(indices), ptr(data), current_stream_ptr()), 'psb_stencil_fill')
    if raw:
        return indptr, indices, data
    return DeviceCSR(indptr=indptr, indices=indices, data=data, shape=(row_hi - row_lo, n))


class FDBratu2D:
    """-Lap(u) - alpha*exp(-u) = 0 on (-1,1)^2 (examples/FDBratu2D.py:10-29)."""

    def __init__(self, m=4, alpha=0.5):
        self.m = m
        self.alpha = alpha
        self.A = -fd_laplacian_2d(-1.0, 1.0, m)

    def initialU(self):
        return np.ones(self.m * self.m)

    def evalF(self, u):
        return self.A * u - self.alpha * np.exp(-u)

    def evalJ(self, u):
        J = self.A.copy()
        shift = self.alpha * np.exp(-u)
        J.setdiag(J.diagonal() + shift)
        return J


def load_dh_matrix(lev, root=None):
    """DH-Matrix-<lev> as scipy CSR.  Looks for the MatrixMarket file under
    ``root`` or the reference checkout, else for the repo's fixture copy
    (tests/golden/matrices/DH-Matrix-<lev>.npz: the COO triplets exactly as
    ``scipy.io.mmread`` expands the symmetric file, so ``tocsr()`` gives the
    identical matrix)."""
    from scipy.io import mmread
    here = os.path.dirname(os.path.abspath(__file__))
    name = 'DH-Matrix-%d' % lev
    cands = []
    if root is not None:
        cands.append(os.path.join(root, name + '.mtx'))
    cands.append(os.path.join(here, '..', 'tests', 'golden', 'matrices',
                              name + '.npz'))
    cands.append(os.path.join('/root/reference/TestMatrices', name + '.mtx'))
    for c in cands:
        if not os.path.exists(c):
            continue
        if c.endswith('.npz'):
            z = np.load(c)
            shape = tuple(int(v) for v in z['shape'])
            return sp.coo_matrix((z['data'], (z['row'], z['col'])),
                                 shape=shape).tocsr()
        return sp.csr_matrix(sp.coo_matrix(mmread(c)))
    raise FileNotFoundError(name + ' not found')


def dh_test_problem(lev, seed=2024, root=None):
    """(A, b, x) with b = A*x for a seeded random x
    (examples/DHTestProblem.py:24-36; the reference's RNG is unseeded)."""
    assert 0 <= lev <= 16
    A = load_dh_matrix(lev, root)
    x = np.random.default_rng(seed).random(A.shape[0])
    return A, A * x, x


class DeviceFDBratu2D:
    """FDBratu2D with u, F and the Jacobian resident in HBM:
    evalF / evalJ take and return CUDA tensors / a DeviceCSR whose diagonal values are
    rewritten in place -- the structure is uploaded once.  Drop it into NewtonSolver exactly
    like the host class; the linear solvers accept the DeviceCSR and the device right-hand
    side and return a device solution."""

    def __init__(self, m=4, alpha=0.5):
        import torch
        from .device import DeviceCSR, to_device
        self.m, self.alpha = m, alpha
        self.A = -fd_laplacian_2d(-1.0, 1.0, m)
        n = m * m
        rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(self.A.indptr))
        dpos = np.flatnonzero(rows == self.A.indices)
        assert dpos.size == n
        self._dA = DeviceCSR(self.A)
        self._dJ = DeviceCSR(self.A)
        self._diag_pos = torch.from_numpy(dpos).cuda()
        self._a_diag = to_device(self.A.diagonal())
        self._torch = torch

    def initialU(self):
        return self._torch.ones(self.m * self.m, dtype=self._torch.float64, device='cuda')

    def evalF(self, u):
        import ctypes as C
        from . import _native as nat
        from .device import current_stream_ptr, ptr
        Au = self._dA.matvec(u)
        F = self._torch.empty_like(u)
        nat.check(nat.lib().psb_bratu_residual(u.numel(), ptr(Au), ptr(u), float(self.alpha), ptr(F),
                                               current_stream_ptr()), 'psb_bratu_residual')
        return F

    def evalJ(self, u):
        from . import _native as nat
        from .device import current_stream_ptr, ptr
        nat.check(nat.lib().psb_bratu_jacobian(u.numel(), ptr(self._diag_pos), ptr(self._a_diag), ptr(u),
                                               float(self.alpha), ptr(self._dJ.data), current_stream_ptr()),
                  'psb_bratu_jacobian')
        return self._dJ
