"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a,
loads, and exports every symbol include/pysolv_b200.h declares; the ctypes
table in pysolvers_b200/_native.py covers the header one to one."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, 'include', 'pysolv_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(psb_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def libpath():
    from pysolvers_b200.csrc.build import build_native
    return build_native()


def test_header_declares_functions():
    fns = _header_functions()
    assert 'psb_pcg_solve' in fns and 'psb_spmv' in fns and len(fns) >= 10


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    missing = [f for f in _header_functions() if not hasattr(lib, f)]
    assert not missing, 'declared in the header but not exported: %s' % missing


def test_ctypes_table_matches_header():
    from pysolvers_b200 import _native
    assert sorted(_native.SIGNATURES) == _header_functions()


def test_version_and_error_text(libpath):
    from pysolvers_b200 import _native
    lib = _native.lib()
    assert lib.psb_version() >= 100
    # argument validation happens before any CUDA call: NULL handle -> PSB_ERR_ARG
    rc = lib.psb_spmv(None, None, None, None)
    assert rc == -2
    assert b'NULL' in lib.psb_last_error()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure; nothing under pysolvers_b200/ may
    import it."""
    pkg = os.path.join(ROOT, 'pysolvers_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_stencil_prefix_matches_host_generators():
    """psb_stencil_nnz (host arithmetic of the device generator) against the row pointers of the
    numpy generators, which are pinned bit-exactly to the reference's (test_oracle_golden)."""
    from pysolvers_b200 import _native as nat
    from pysolvers_b200.problems import fd_laplacian_2d, fd_laplacian_3d
    lib = nat.lib()
    for dim, m, gen in ((2, 1, fd_laplacian_2d), (2, 2, fd_laplacian_2d), (2, 9, fd_laplacian_2d),
                        (2, 40, fd_laplacian_2d), (3, 1, fd_laplacian_3d), (3, 2, fd_laplacian_3d),
                        (3, 6, fd_laplacian_3d), (3, 11, fd_laplacian_3d)):
        A = gen(0.0, 1.0, m)
        n = m ** dim
        for k in range(n + 1):
            assert lib.psb_stencil_nnz(dim, m, 0, k) == A.indptr[k]
        assert lib.psb_stencil_nnz(dim, m, n // 3, n) == A.indptr[n] - A.indptr[n // 3]
    assert lib.psb_stencil_nnz(4, 3, 0, 1) == -1 and lib.psb_stencil_nnz(2, 3, 5, 2) == -1


def test_argument_validation_without_a_device():
    """Entry points reject bad arguments with an error code and a message BEFORE any device work."""
    import ctypes as C
    from pysolvers_b200 import _native as nat
    lib = nat.lib()
    out = C.c_void_p()
    z32 = (C.c_int32 * 4)(0, 1, 2, 3)
    one = (C.c_double * 1)(1.0)
    # split LU: n1 must be < n, NULL maps, out-of-range map entries
    assert lib.psb_splitlu2_create(4, 4, 0, None, None, None, None, one, one, z32, None, z32, None, C.byref(out)) < 0
    assert b'n1' in lib.psb_last_error()
    assert lib.psb_splitlu2_create(4, 0, 0, None, None, None, None, one, one, None, None, z32, None, C.byref(out)) < 0
    bad = (C.c_int32 * 4)(0, 1, 2, 9)
    assert lib.psb_splitlu2_create(4, 0, 0, None, None, None, None, one, one, bad, None, z32, None, C.byref(out)) < 0
    assert b'map entry' in lib.psb_last_error()
    # generators
    assert lib.psb_stencil_fill(5, 3, 0, 1, 1.0, 1.0, None, None, None, None) < 0
    assert lib.psb_stencil_fill(2, 3, 0, 10, 1.0, 1.0, z32, z32, one, None) < 0        # row_hi > n
    # levels / heights helpers: NULL pointers
    assert lib.psb_tri_levels(3, None, None, 1, None) < 0
    assert lib.psb_tri_heights_upper(3, None, None, None) < 0
    # kernel selection on a NULL factor
    assert lib.psb_trsv_set_kernel(None, 1) < 0
    assert lib.psb_trsv_info2(None, None) < 0
    # aggregation sweep: count larger than n
    cnt = C.c_int64(7)
    i64 = (C.c_int64 * 4)(0, 0, 0, 0)
    assert lib.psb_sa_phase1(3, i64, i64, i64, i64, C.byref(cnt)) < 0
