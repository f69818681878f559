import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'reference_golden.npz')
GOLDEN_LARGE = os.path.join(ROOT, 'tests', 'golden', 'reference_golden_large.npz')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden():
    """Fixtures produced by running the reference itself
    (tests/golden/make_golden.py)."""
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def golden_large():
    """At-size fixtures of the reference (tests/golden/make_golden_large.py): SHA-256 digests of
    the m = 128 / 256 smoothed-aggregation hierarchies and IC factors, IC-PCG histories."""
    z = np.load(GOLDEN_LARGE)
    return {k: z[k] for k in z.files}


def csr_digest(M):
    """(sha indptr, sha indices, sha data) with make_golden_large.py's convention."""
    import hashlib
    import scipy.sparse as sp
    M = sp.csr_matrix(M)
    h = lambda a, t: hashlib.sha256(np.ascontiguousarray(a, dtype=t).tobytes()).hexdigest()
    return h(M.indptr, np.int32), h(M.indices, np.int32), h(M.data, np.float64)


def assert_csr_digest(M, g, prefix):
    import scipy.sparse as sp
    M = sp.csr_matrix(M)
    assert tuple(int(v) for v in g[prefix + '/shape']) == M.shape, prefix
    assert int(g[prefix + '/nnz']) == M.nnz, prefix
    got = csr_digest(M)
    want = tuple(str(g[prefix + k]) for k in ('/sha_indptr', '/sha_indices', '/sha_data'))
    assert got == want, prefix


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pysolvers_b200.csrc.build import build_native
    build_native()
    return torch.device('cuda:0')


@pytest.fixture(scope='session', autouse=True)
def _single_thread_blas():
    """The fixtures were generated with one OpenBLAS thread (ddot's summation
    order depends on the thread count); pin the same here."""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1, user_api='blas'):
            yield
    except ImportError:
        yield


def assert_same_history(got, want, what=''):
    """Bit-equal when the same BLAS kernels run (the build container); on a
    host whose OpenBLAS picks other ddot kernels the last bits may move, so
    fall back to 1e-9 relative."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, what
    if not np.array_equal(got, want):
        assert rel_err(got, want) < 1e-9, what


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), np.finfo(np.float64).tiny)
    return np.max(np.abs(a - b) / den) if a.size else 0.0


def golden_csr(g, prefix):
    import scipy.sparse as sp
    shape = tuple(int(v) for v in g[prefix + '/shape'])
    return sp.csr_matrix((g[prefix + '/data'], g[prefix + '/indices'],
                          g[prefix + '/indptr']), shape=shape)


def assert_history_close(hist, want, A, x, rtol=1e-10, what=''):
    """Residual norms agree to ``rtol`` relative down to the cancellation floor
    of evaluating b - A x in fp64: an absolute 8 eps ||A||_inf ||x|| (below it
    the reference's own digits are rounding noise)."""
    import scipy.sparse as sp
    hist, want = np.asarray(hist), np.asarray(want)
    assert hist.shape == want.shape, (what, hist.shape, want.shape)
    # NB: abs() of a scipy matrix sorts ITS indices in place (sum_duplicates) -- work on a
    # copy: the stored column order is part of the reference's arithmetic
    norm_a = float(abs(sp.csr_matrix(A, copy=True)).sum(axis=1).max())
    floor = 8.0 * np.finfo(np.float64).eps * norm_a * float(np.linalg.norm(x))
    bad = np.abs(hist - want) > rtol * np.abs(want) + floor
    assert not bad.any(), (what, hist[bad], want[bad], floor)
