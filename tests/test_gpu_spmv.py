"""SpMV parity on the B200: CUDA kernels (through the C ABI) vs scipy's
csr_matvec, the routine the reference executes behind ``A*x``
(PySolvers/Linear/IterativeLinearSolver.py:104)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _mats():
    from pysolvers_b200.problems import fd_laplacian_2d, fd_laplacian_3d, load_dh_matrix
    rng = np.random.default_rng(5)
    out = {
        'lap2d_m7': -fd_laplacian_2d(0.0, 1.0, 7),
        'lap2d_m64': -fd_laplacian_2d(0.0, 1.0, 64),
        'lap2d_m300': -fd_laplacian_2d(0.0, 1.0, 300),
        'lap3d_m20': fd_laplacian_3d(0.0, 1.0, 20),
        'dh12': load_dh_matrix(12),
        'rect': sp.random(1000, 377, density=0.02, random_state=rng, format='csr'),
        'empty_rows': sp.csr_matrix(sp.random(513, 513, density=0.002, random_state=rng)),
        'one_row': sp.csr_matrix(np.arange(1.0, 6.0).reshape(1, 5)),
        'dense_rows': sp.random(300, 300, density=0.5, random_state=rng, format='csr'),
    }
    # a few very long rows among short ones (skewed histogram)
    skew = sp.lil_matrix((2000, 2000))
    skew.setdiag(2.0)
    skew[7, :] = rng.random(2000)
    skew[1500, ::2] = 1.5
    out['skewed'] = skew.tocsr()
    return out


@pytest.mark.parametrize('name', ['lap2d_m7', 'lap2d_m64', 'lap2d_m300', 'lap3d_m20', 'dh12',
                                  'rect', 'empty_rows', 'one_row', 'dense_rows', 'skewed'])
def test_spmv_matches_scipy(cuda, name):
    import torch
    from pysolvers_b200 import _native as nat
    from pysolvers_b200.device import DeviceCSR, to_device
    A = _mats()[name]
    x = np.random.default_rng(1).standard_normal(A.shape[1])
    want = A @ x
    dA = DeviceCSR(A)
    xd = to_device(x)
    kinds = [dA.info()['kind']]
    if name == 'skewed':
        assert kinds[0] == nat.SPMV_MERGE          # two 1 000 - 2 000-entry rows among 1-entry rows
    for kind in (nat.SPMV_STREAM, nat.SPMV_STREAM | nat.SPMV_TILE512, nat.SPMV_STREAM_LSU,
                 nat.SPMV_STREAM_LSU | nat.SPMV_TILE512, nat.SPMV_VECTOR, nat.SPMV_MERGE):
        try:
            dA.set_kind(kind)
        except nat.NativeError:
            continue
        got = dA.matvec(xd).cpu().numpy()
        if (kind & 15) not in (nat.SPMV_VECTOR, nat.SPMV_MERGE):
            # STREAM sums each row in stored order from +0, no FMA: bit-identical
            assert np.array_equal(got, want), (name, kind)
        else:
            scale = np.abs(A.copy()) @ np.abs(x) + 1e-300
            assert np.max(np.abs(got - want) / scale) < 1e-14, (name, kind)


def test_merge_path_long_rows_and_epilogues(cuda):
    """Merge-path kind (north_star item 1: 'merge-path kernels selected by row-length histogram'):
    rows far longer than one CTA's share (the cross-CTA carry fix-up), empty rows, a dense last
    row; every epilogue against numpy.  Deterministic: two runs give the same bits."""
    import torch
    from pysolvers_b200 import _native as nat
    from pysolvers_b200.device import DeviceCSR, to_device, ptr, current_stream_ptr
    rng = np.random.default_rng(11)
    n = 30000
    M = sp.lil_matrix((n, n))
    M.setdiag(rng.random(n) + 1.0)
    M[5, :] = rng.standard_normal(n)               # 30 000 entries: ~17 CTA tiles
    M[6, ::3] = 1.0
    M[12345, 100:20100] = rng.standard_normal(20000)
    M[n - 1, :] = rng.standard_normal(n)           # long last row
    A = sp.csr_matrix(M)
    A[100:140, :] = 0                              # a stretch of (explicitly) empty rows
    A.eliminate_zeros()
    x, f = rng.standard_normal(n), rng.standard_normal(n)
    dA = DeviceCSR(A)
    assert dA.info()['kind'] == nat.SPMV_MERGE
    xd, fd = to_device(x), to_device(f)
    scale = np.abs(A) @ np.abs(x) + 1e-300
    y1 = dA.matvec(xd).cpu().numpy()
    y2 = dA.matvec(xd).cpu().numpy()
    assert np.array_equal(y1, y2)
    assert np.max(np.abs(y1 - A @ x) / scale) < 1e-14
    lib, st = nat.lib(), current_stream_ptr()
    out = torch.empty(n, dtype=torch.float64, device='cuda')
    nat.check(lib.psb_spmv_residual(dA.handle, ptr(xd), ptr(fd), ptr(out), st))
    assert np.max(np.abs(out.cpu().numpy() - (f - A @ x)) / (scale + np.abs(f))) < 1e-14
    dot = torch.zeros(1, dtype=torch.float64, device='cuda')
    nat.check(lib.psb_spmv_dot(dA.handle, ptr(xd), ptr(out), ptr(dot), st))
    want = float(x @ (A @ x))
    assert abs(float(dot.item()) - want) <= 1e-12 * float(np.abs(x) @ scale)
    acc = to_device(f.copy())
    nat.check(lib.psb_spmv_add(dA.handle, ptr(xd), ptr(acc), st))
    assert np.max(np.abs(acc.cpu().numpy() - (f + A @ x)) / (scale + np.abs(f))) < 1e-14
    diag = A.diagonal()
    diag[diag == 0.0] = 1.0                        # the emptied rows
    dinv = to_device(1.0 / diag)
    nat.check(lib.psb_jacobi_sweep(dA.handle, ptr(dinv), 2.0 / 3.0, ptr(fd), ptr(xd), ptr(out), st))
    wantj = x + (2.0 / 3.0) * ((f - A @ x) * (1.0 / diag))
    assert np.max(np.abs(out.cpu().numpy() - wantj) / (np.abs(wantj) + scale)) < 1e-13


def test_kernel_choice_long_rows_in_a_large_matrix(cuda):
    """The merge-path kind is for matrices in which the longest row is a visible share of the work
    (>= nnz / 512).  A large block with rows of up to ~600 entries among short and empty ones (the U12
    block of a coarse LU factor) stays on the sub-warp kernel, which is twice as fast there; both kinds
    agree with scipy to rounding."""
    from pysolvers_b200 import _native as nat
    from pysolvers_b200.device import DeviceCSR, to_device
    rng = np.random.default_rng(23)
    n_rows, n_cols = 60000, 2048
    lens = np.zeros(n_rows, dtype=np.int64)
    busy = rng.random(n_rows) < 0.2                               # 80 % of the rows are empty
    lens[busy] = rng.choice([4, 25, 120, 600], size=int(busy.sum()), p=[.4, .35, .2, .05])
    indptr = np.concatenate([[0], np.cumsum(lens)])
    indices = np.concatenate([np.sort(rng.choice(n_cols, size=k, replace=False)) for k in lens if k])
    A = sp.csr_matrix((rng.standard_normal(indices.size), indices, indptr), shape=(n_rows, n_cols))
    assert lens.max() >= 512 and lens.max() >= 16 * A.nnz / n_rows and lens.max() * 512 < A.nnz
    dA = DeviceCSR(A)
    assert dA.info()['kind'] != nat.SPMV_MERGE
    x = rng.standard_normal(n_cols)
    scale = np.abs(A) @ np.abs(x) + 1e-300
    xd = to_device(x)
    y_auto = dA.matvec(xd).cpu().numpy()
    dA.set_kind(nat.SPMV_VECTOR)
    y_vec = dA.matvec(xd).cpu().numpy()
    dA.set_kind(nat.SPMV_MERGE)
    y_merge = dA.matvec(xd).cpu().numpy()
    for y in (y_auto, y_vec, y_merge):
        assert np.max(np.abs(y - A @ x) / scale) < 1e-14


def test_spmv_epilogues(cuda):
    import torch
    from pysolvers_b200 import _native as nat
    from pysolvers_b200.device import DeviceCSR, to_device, ptr, current_stream_ptr
    from pysolvers_b200.problems import fd_laplacian_2d
    A = -fd_laplacian_2d(0.0, 1.0, 50)
    n = A.shape[0]
    rng = np.random.default_rng(2)
    x, f, y0 = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    dA = DeviceCSR(A)
    lib = nat.lib()
    xd, fd = to_device(x), to_device(f)
    st = current_stream_ptr()
    # y = A x and dot = x.y
    yd = torch.empty(n, dtype=torch.float64, device='cuda')
    dd = torch.zeros(1, dtype=torch.float64, device='cuda')
    nat.check(lib.psb_spmv_dot(dA.handle, ptr(xd), ptr(yd), ptr(dd), st))
    Ax = A @ x
    assert np.array_equal(yd.cpu().numpy(), Ax)
    assert abs(dd.item() - np.dot(x, Ax)) <= 1e-13 * np.dot(np.abs(x), np.abs(Ax))
    # deterministic: same bits on a second run
    dd2 = torch.zeros(1, dtype=torch.float64, device='cuda')
    nat.check(lib.psb_spmv_dot(dA.handle, ptr(xd), ptr(yd), ptr(dd2), st))
    assert dd.item() == dd2.item()
    # residual
    nat.check(lib.psb_spmv_residual(dA.handle, ptr(xd), ptr(fd), ptr(yd), st))
    assert np.array_equal(yd.cpu().numpy(), f - Ax)
    # y += A x
    yd = to_device(y0)
    nat.check(lib.psb_spmv_add(dA.handle, ptr(xd), ptr(yd), st))
    assert np.array_equal(yd.cpu().numpy(), y0 + Ax)
    # Jacobi sweeps (ClassicSmoothers.py:12-14), omega = 1 and 2/3
    dinv = np.reciprocal(A.diagonal())
    dinv_d = to_device(dinv)
    out = torch.empty(n, dtype=torch.float64, device='cuda')
    nat.check(lib.psb_jacobi_sweep(dA.handle, ptr(dinv_d), 1.0, ptr(fd), ptr(xd), ptr(out), st))
    assert np.array_equal(out.cpu().numpy(), x + np.multiply(dinv, f - Ax))
    om = 2.0 / 3.0
    nat.check(lib.psb_jacobi_sweep(dA.handle, ptr(dinv_d), om, ptr(fd), ptr(xd), ptr(out), st))
    assert np.array_equal(out.cpu().numpy(), x + om * np.multiply(dinv, f - Ax))


def test_dot_matches_numpy(cuda):
    import torch
    from pysolvers_b200 import _native as nat
    from pysolvers_b200.device import to_device, ptr, current_stream_ptr
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 255, 256, 257, 100001, 1 << 20):
        a, b = rng.standard_normal(n), rng.standard_normal(n)
        out = torch.zeros(1, dtype=torch.float64, device='cuda')
        ad, bd = to_device(a), to_device(b)       # keep alive until the kernel has run
        nat.check(nat.lib().psb_dot(n, ptr(ad), ptr(bd), ptr(out), current_stream_ptr()))
        ref = np.dot(a, b)
        assert abs(out.item() - ref) <= 1e-13 * np.dot(np.abs(a), np.abs(b)), n


def test_mvmult_drop_in(cuda):
    from pysolvers_b200.Linear import mvmult
    from pysolvers_b200.problems import load_dh_matrix
    A = load_dh_matrix(9)
    x = np.random.default_rng(4).random(A.shape[0])
    assert np.array_equal(mvmult(A, x), A * x)


def test_spmv_3d_slab_full_width(cuda):
    """The 3-D 7-point Laplacian at the per-GPU share of C4 on 8 GPUs (256^3 = 16.8 M rows, 117 M
    nonzeros, stored order [k, k-m^2, k+m^2, k-m, k+m, k-1, k+1]): bit-identical to scipy."""
    from pysolvers_b200.device import DeviceCSR, to_device
    from pysolvers_b200.problems import fd_laplacian_3d
    A = fd_laplacian_3d(0.0, 1.0, 256)
    n = A.shape[0]
    assert n == 256 ** 3 and A.nnz == 7 * n - 6 * 256 ** 2
    x = np.random.default_rng(11).standard_normal(n)
    dA = DeviceCSR(A)
    y = dA.matvec(to_device(x)).cpu().numpy()
    assert np.array_equal(y, A @ x)


def test_spmv_16bit_column_deltas(cuda):
    """Banded matrices get a second copy of the column indices as 16-bit distances from the
    diagonal (10 instead of 12 bytes per entry in the STREAM kernels); the result stays
    bit-identical to scipy.  Distances of -32768 ... 32767 fit, 32768 does not."""
    from pysolvers_b200 import _native as nat
    from pysolvers_b200.device import DeviceCSR, to_device
    rng = np.random.default_rng(21)
    n = 120000
    nat.check(nat.lib().psb_csr_set_cols16(1), 'psb_csr_set_cols16')      # optional path, off by default
    try:
        _cols16_cases(rng, n, DeviceCSR, to_device)
    finally:
        nat.check(nat.lib().psb_csr_set_cols16(0), 'psb_csr_set_cols16')
    A = sp.diags([np.ones(n - 1), np.ones(n)], [-1, 0], shape=(n, n), format='csr')
    assert not DeviceCSR(A).info()['cols16']


def _cols16_cases(rng, n, DeviceCSR, to_device):

    def banded(offsets):
        diags = [rng.standard_normal(n - abs(o)) for o in offsets]
        return sp.diags(diags, offsets, shape=(n, n), format='csr')

    for offsets, want in (((0, -1, 1, -300, 300), True),
                          ((0, -32768, 32767, 5), True),
                          ((0, 32768, -7), False),
                          ((0, -32769, 2), False)):
        A = banded(offsets)
        dA = DeviceCSR(A)
        assert dA.info()['kind'] == 1
        assert dA.info()['cols16'] == want, offsets
        x = rng.standard_normal(n)
        y = dA.matvec(to_device(x)).cpu().numpy()
        assert np.array_equal(y, A @ x), offsets
    # odd sizes / unaligned tails of the 16-bit staging: n and nnz not multiples of 8
    for n2 in (65537, 70001):
        A = sp.diags([rng.standard_normal(n2 - 3), rng.standard_normal(n2), rng.standard_normal(n2 - 1)],
                     [-3, 0, 1], shape=(n2, n2), format='csr')
        dA = DeviceCSR(A)
        assert dA.info()['cols16']
        x = rng.standard_normal(n2)
        assert np.array_equal(dA.matvec(to_device(x)).cpu().numpy(), A @ x)


def test_device_generators_bit_identical(cuda):
    """The Laplacians assembled in HBM (psb_stencil_fill) are the host generators' matrices bit for
    bit -- row pointers, stored column order, values -- for whole grids and row slabs, and a solver
    takes the DeviceCSR in place of the scipy matrix."""
    import contextlib
    import io
    from pysolvers_b200 import CommonSolverArgs
    from pysolvers_b200.Linear import PCG
    from pysolvers_b200.problems import device_fd_laplacian, fd_laplacian_2d, fd_laplacian_3d

    def same(dA, A):
        B = dA.to_scipy()
        assert B.shape == A.shape
        assert np.array_equal(B.indptr, A.indptr) and np.array_equal(B.indices, A.indices)
        assert np.array_equal(B.data, A.data)

    for m in (1, 2, 7, 300):
        same(device_fd_laplacian(2, 0.0, 1.0, m), fd_laplacian_2d(0.0, 1.0, m))
    same(device_fd_laplacian(2, -1.0, 1.0, 33, negate=True), -fd_laplacian_2d(-1.0, 1.0, 33))
    same(device_fd_laplacian(2, 0.0, 1.0, 50, row_lo=1203, row_hi=2077), fd_laplacian_2d(0.0, 1.0, 50, 1203, 2077))
    for m in (1, 2, 20):
        same(device_fd_laplacian(3, 0.0, 1.0, m), fd_laplacian_3d(0.0, 1.0, m))
    same(device_fd_laplacian(3, 0.0, 1.0, 16, row_lo=777, row_hi=3001), fd_laplacian_3d(0.0, 1.0, 16, 777, 3001))

    A = -fd_laplacian_2d(0.0, 1.0, 48)
    dA = device_fd_laplacian(2, 0.0, 1.0, 48, negate=True)
    b = np.ones(48 * 48)
    out = []
    for M in (A, dA):
        s = PCG(CommonSolverArgs(maxiter=500, tau=1e-8)).makeSolver()
        with contextlib.redirect_stdout(io.StringIO()):
            out.append(s.solve(M, b))
    assert out[0].iters() == out[1].iters() and np.array_equal(out[0].soln(), out[1].soln())
